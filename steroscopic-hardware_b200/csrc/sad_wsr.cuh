// sad_wsr.cuh — warp-specialised kernel for the large windows, block_size 18..31 (h = 9..15): the scheduling of sad_ws.cuh (fixed
// warp roles, one __syncthreads per batch, TMA tile loads three batches ahead) with the rows of the window in a SHARED-MEMORY ring.
//
// Why not the register ring of sad_ws.cuh: a consumer would hold 2 x (window + 1) = up to 64 registers per disparity group, which
// leaves one group per warp and 17 bits of window sum (31 * 31 * 255) need 32-bit arithmetic on top.  Here the consumer loads the
// row that enters AND the row that leaves the window from the ring (one more LDS.64 per four evaluations), so it carries four
// running sums per group and nothing else.
//
// A CTA owns one 32-column strip of a row band and one chunk of 9 disparity groups (36 disparity slots; described below) or of 13
// groups (52 slots, batches of 7 rows: the cheaper cover for 10..13, 37..39 and 64..65 groups), 20 warps (warp w lives on
// sub-partition w % 4):
//   warps 0..2    walkers: a batch is 10 rows x 9 groups = 90 (row, group) walks = the lanes of three warps.  A walk is 32 + 2h steps
//                 (up to 62), keeps the 2h+1 old terms of both packed sums as SSA registers (setmaxnreg gives this warpgroup 120),
//                 issues its shared-memory loads four steps early and does its running-sum adds on the FMA pipe;
//   warp 3        finisher (key -> d*255/D LUT -> store, or the global key map when the disparity range is chunked);
//   warp 7        TMA loader (cp.async.bulk.tensor.3d + mbarrier, zero fill outside the image) + replication of the left pixels;
//   warps 8..19   consumers: five of them own two groups (one) for 32 columns (lanes = columns; three groups on sub-partition 3, two
//                 on each of the sub-partitions that carry a walker), so that a consumer's chain per batch is about as long as a
//                 walker's: the loads of five rows first (a shared-memory load cannot be hoisted over the atomicMin of the row
//                 before), then 32-bit vertical running sums (raw packed sum + high-lane sum, keys sum*512+d, as the h = 8
//                 instance of sad_ws.cuh); the consumers of a pixel meet in one key through a shared-memory atomicMin.
// Ring: NB = ceil(window / 10) + 2 batches of 10 rows; while the walkers write batch i the consumers read batch i-1 and the rows
// window (= 10 q + m) rows older, which lie at compile-time offsets in batches i-1-q and i-2-q.  The ring starts zeroed: rows above
// the band are the zero padding of the box filter.
//
// Measured (B200, 1080p, 8 frames per launch, profiles/r02_wsr_*): block 31: 39 / 75 / 252 us per frame at max_disparity 32 / 64 / 256
// against 101 / 102 / 379 us for the mbarrier-pipelined ring kernel (sad_ring.cuh) it replaces as the planner's choice.  What was
// tried and measured: batches of 14 rows on four walker warps (40 / 90 / 307 us: one sub-partition then carries a walker AND three
// consumers), the finishing dealt to the three idle consumer warps (42 us: they sit on the walker sub-partitions), shared-memory
// stores instead of the atomicMin (no change), the consumer's adds as shared or separate three-input adds (no change), walker sums
// on the ALU pipe (+2.5 %), the hand-over between the roles through named barriers (bar.arrive / bar.sync per batch and direction
// instead of one __syncthreads: bit-exact, 2..8 % slower — the ring leaves the walkers one batch of slack either way), the walk's
// running sum as difference-then-add (one dependent add per step: no change).  Role-isolation runs (WSR_SKIP): walkers alone 29 us, consumers alone 26 us, together 39 us — issue
// slots 56 %, ALU pipe 64 %, shared-memory pipe 63 % of peak: no single resource binds, the kernel is short of independent warps
// (one CTA per SM, 201 KB of shared memory).
#pragma once
#include <type_traits>
#include "sad_common.cuh"

// developer switches behind the measurements above
#ifndef WSR_RB
#define WSR_RB 10       // rows per batch
#endif
#ifndef WSR_FMAS
#define WSR_FMAS 2
#endif
#ifndef WSR_NG2
#define WSR_NG2 1       // 1 = two groups per consumer warp where possible (measured: D=64 78.2 -> 74.8 us, D=256 265.7 -> 252.2 us at block 31)
#endif
#ifndef WSR_FIN3
#define WSR_FIN3 0      // 1 = the finishing is dealt to the three consumer warps without a group (measured slower)
#endif
#ifndef WSR_SKIP
#define WSR_SKIP 0      // developer timing experiments (results wrong): 1 = walkers idle, 2 = consumers idle, 8 = finisher idle, 128 = no tile loads
#endif

namespace sadgpu {

// Consumer warp k (sub-partition k % 4) -> first disparity group of the chunk it owns (or -1) and how many.  One group per warp:
// sub-partition 3 takes three groups and the others (which carry a walker) two.  WSR_NG2: two groups per warp where possible, so that
// a consumer's chain is about as long as a walker's.
__host__ __device__ constexpr int wsr_group(int ngc, int k)
{
    if (ngc == 13) {                                    // 3 3 3 on the walker sub-partitions, 2 + 2 on sub-partition 3
        const int t13[12] = {4, 7, 10, 0, -1, -1, -1, 2, -1, -1, -1, -1};
        return t13[k];
    }
#if WSR_NG2
    const int t[12] = {3, 5, 7, 0, -1, -1, -1, 2, -1, -1, -1, -1};
#else
    const int t[12] = {3, 5, 7, 0, 4, 6, 8, 1, -1, -1, -1, 2};
#endif
    return t[k];
}
__host__ __device__ constexpr int wsr_ngroups(int ngc, int k)
{
    if (ngc == 13) return k <= 2 ? 3 : 2;
#if WSR_NG2
    return k <= 3 ? 2 : 1;
#else
    (void)k; return 1;
#endif
}

template <int HALF, int NGC_ = 9> struct WsrCfg {
    static_assert(HALF >= 5 && HALF <= 15, "shared-memory-ring warp-specialised kernel: block_size 10..31");
    static constexpr int WIN = 2 * HALF + 1;
    static_assert(NGC_ == 9 || NGC_ == 13, "chunks of 9 or 13 disparity groups");
    static constexpr int NGC = NGC_;                    // groups per chunk
    static constexpr int RB = NGC == 9 ? WSR_RB : 7;    // rows per batch: RB * NGC = 90 / 91 walker lanes = three warps
    static constexpr int TW = 32, TWP = 33, CW = 32;
    static constexpr int NSTEP = TW + 2 * HALF;
    static constexpr int NWW = (RB * NGC + 31) / 32;    // walker warps (warpgroup 0)
    // warp w lives on sub-partition w % 4.  Three walker warps: sub-partitions 0..2 carry a walker and two consumers each,
    // sub-partition 3 the finisher (warp 3), the loader (warp 7) and three consumers
    static constexpr int W_FIN = NWW < 4 ? 3 : 6, W_LOAD = 7, W_CONS = 8, K = 12;
    static constexpr bool FIN3 = NWW == 3 && WSR_FIN3;  // the three consumer warps without a group finish (a third of the rows each)
    static constexpr int NT = 32 * (W_CONS + K);        // 640 threads
    static constexpr int Q = WIN / RB, M = WIN % RB;    // the row leaving the window: batch - Q (rows >= M) or batch - Q - 1
    static constexpr int NB = (WIN + RB - 1) / RB + 2;  // ring length in batches
    static constexpr int NR = NB * RB;
    static constexpr int NTILE = 4;                     // tiles are requested three batches ahead, completed two ahead
    static constexpr int REGS_WALK = 120, REGS_SERVICE = 40;
    static constexpr int OFF = walk_off(HALF);
    static constexpr int NWALKW = walk_words(HALF, NSTEP);
    static constexpr int RW = NGC - 1 + NWALKW;
    // a walker warp holds lanes of four rows: row strides of 4 (left) and 12 (right) mod 32 words spread their loads over the banks
    static constexpr int LW0 = (CW + 2 * HALF + 3) & ~3;
    static constexpr int LW = LW0 + ((4 - LW0 % 32 + 32) % 32);
    static constexpr int RWT0 = ((RW * 4 + 12 + 15) / 16) * 4;
    static constexpr int RWT = RWT0 + ((12 - RWT0 % 32 + 32) % 32);
    static constexpr int LSH = (16 - HALF % 16) % 16;
    static constexpr int LBOX = ((LSH + LW + 15) / 16) * 16;
    static constexpr int HROW = NGC * TWP;                                     // uint2 per ring row
    static constexpr int H_BYTES = ((NR * HROW * 8 + 127) / 128) * 128;
    static constexpr int R_BYTES = ((RB * RWT * 4 + 127) / 128) * 128;
    static constexpr int LRAW_BYTES = ((RB * LBOX + 127) / 128) * 128;
    static constexpr int L_BYTES = RB * LW * 4;
    static constexpr int PK_BYTES = RB * TW * 4;
    static constexpr int OFF_R = H_BYTES;
    static constexpr int OFF_LRAW = OFF_R + NTILE * R_BYTES;
    static constexpr int OFF_L = OFF_LRAW + NTILE * LRAW_BYTES;
    static constexpr int OFF_PK = OFF_L + NTILE * L_BYTES;
    static constexpr int OFF_LUT = OFF_PK + 2 * PK_BYTES;
    static constexpr int OFF_MBAR = OFF_LUT + 1040;
    static constexpr int SMEM = OFF_MBAR + 64;
    static_assert(M == 0 ? NB >= Q + 2 : NB > Q + 2, "ring geometry");
    static_assert(RB * NGC <= 32 * NWW && NWW == 3, "walker lanes");
    static_assert(RWT * 4 <= 256 && LBOX <= 256, "TMA box");
    static_assert(128 * REGS_WALK + 128 * REGS_SERVICE + 384 * 96 <= NT * 96, "register budget");
    static_assert(SMEM <= 232448, "shared memory");
};

template <int HALF, int NGC_>
__global__ void __launch_bounds__(WsrCfg<HALF, NGC_>::NT, 1) sad_wsr_kernel(const __grid_constant__ FastArgs a)
{
    using C = WsrCfg<HALF, NGC_>;
    constexpr int TW = C::TW, TWP = C::TWP, RB = C::RB, NGC = C::NGC, WIN = C::WIN, HROW = C::HROW;
    extern __shared__ __align__(128) unsigned char smem[];
    uint2* Hs = reinterpret_cast<uint2*>(smem);                                   // [NR][NGC][TWP]
    uint32_t* Ral = reinterpret_cast<uint32_t*>(smem + C::OFF_R);                // [NTILE][RB][RWT]
    uint32_t* Lrep = reinterpret_cast<uint32_t*>(smem + C::OFF_L);               // [NTILE][RB][LW]
    uint32_t* pk = reinterpret_cast<uint32_t*>(smem + C::OFF_PK);                // [2][RB][TW]
    uint8_t* lut = smem + C::OFF_LUT;
    constexpr int LBUF = RB * C::LW, RBUF = C::R_BYTES / 4, PKBUF = RB * TW;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int frame = blockIdx.z / a.NC, chunk = blockIdx.z - frame * a.NC;
    const int x0 = blockIdx.x * C::CW;
    const int yb0 = a.y0 + blockIdx.y * a.BH;
    const int yb1 = min(a.y1, yb0 + a.BH);
    const int g0 = chunk * NGC;
    if (yb0 >= yb1) return;
    // a chunk none of whose disparities is a candidate anywhere in this strip (d > X-h, sad.go:64-67 + :212-218) contributes
    // nothing: chunk 0 always runs and writes every pixel
    if (g0 > 0 && min(x0 + C::CW, a.W) - 1 - HALF < 4 * g0) return;
    const int nga = min(NGC, a.NG - g0);                                      // groups of this chunk that hold candidates at all
    const int r0 = yb0 - HALF;
    const int nb = ((yb1 - yb0) + 2 * HALF + RB - 1) / RB;
    const int xr0 = x0 - HALF - 3 - 4 * (g0 + NGC - 1) - C::OFF;              // image column of right-tile word 0 (multiple of 4)
    const int xr0a = xr0 - (((xr0 % 16) + 16) % 16);                          // tile rows start 16-byte aligned (TMA rule)
    const int rext = (xr0 - xr0a) >> 2;

    for (int d = tid; d < 1040; d += C::NT) lut[d] = d <= a.D ? (uint8_t)((d * 255) / a.D) : 0;
    for (int idx = tid; idx < 2 * PKBUF; idx += C::NT) pk[idx] = 0xFFFFFFFFu;
    for (int idx = tid; idx < C::H_BYTES / 16; idx += C::NT) reinterpret_cast<uint4*>(smem)[idx] = make_uint4(0, 0, 0, 0);

    // Finishing (batch it-2): one key per pixel -> LUT -> store, and the key is re-armed.  All key loads first, then the LUT loads,
    // then the stores: the finisher is a single chain per batch.
    // Rows [RA, RE) of every batch; the caller has passed the start-up barrier.
    auto finisher = [&](auto ra_c, auto re_c) {
        constexpr int RA = decltype(ra_c)::value, RE = decltype(re_c)::value;
        uint8_t* __restrict__ Og = a.out + (long long)frame * a.frameOut;
        const bool xin = x0 + lane < a.W;
        const bool zero0 = x0 == 0 && lane < HALF;                          // X < h: both windows clamp, d = 0 wins (sad.go:212-218)
        for (int it = 0; it < nb + 2; ++it) {
            if (it >= 2 && !(WSR_SKIP & 8)) {
                const int batch = it - 2;
                uint32_t* pkb = pk + (batch & 1) * PKBUF + lane;
                const int row0 = r0 + batch * RB - HALF;                   // image row of item row 0
                const int rb_lo = 2 * HALF - batch * RB, rb_hi = yb1 - row0;   // output rows of this batch: rb in [rb_lo, rb_hi)
                uint8_t* Orow = Og + (long long)row0 * a.pitchOut + x0 + lane;
                const long long grow = ((long long)frame * a.H + row0) * a.W + x0 + lane;
                constexpr int FG = 5;
#pragma unroll
                for (int r0g = RA; r0g < RE; r0g += FG) {
                    uint32_t best[FG];
#pragma unroll
                    for (int g = 0; g < FG; ++g) if (r0g + g < RE) best[g] = pkb[(r0g + g) * TW];
#pragma unroll
                    for (int g = 0; g < FG; ++g) if (r0g + g < RE) pkb[(r0g + g) * TW] = 0xFFFFFFFFu;
                    if (a.NC == 1) {
                        uint8_t v[FG];
#pragma unroll
                        for (int g = 0; g < FG; ++g) if (r0g + g < RE) v[g] = lut[min((zero0 ? 0u : best[g]) & 511u, 1039u)];
#pragma unroll
                        for (int g = 0; g < FG; ++g)
                            if (r0g + g < RE && xin && r0g + g >= rb_lo && r0g + g < rb_hi) Orow[(r0g + g) * a.pitchOut] = v[g];
                    } else {
#pragma unroll
                        for (int g = 0; g < FG; ++g)
                            if (r0g + g < RE && xin && r0g + g >= rb_lo && r0g + g < rb_hi)
                                atomicMin(a.gkey + grow + (long long)(r0g + g) * a.W, zero0 ? 0u : best[g]);
                    }
                }
            }
            __syncthreads();
        }
    };
    using IC0 = std::integral_constant<int, 0>; using ICN = std::integral_constant<int, RB>;
    using IC1 = std::integral_constant<int, (RB + 2) / 3>; using IC2 = std::integral_constant<int, (2 * RB + 2) / 3>;

    if (warp < 4) {
        // ======================= warpgroup 0: walkers (and the finisher when a warp is free) =======================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(C::REGS_WALK));
        if (warp >= C::NWW) {
            __syncthreads();
            if (!C::FIN3 && warp == C::W_FIN) finisher(IC0{}, ICN{});
            else for (int it = 0; it < nb + 2; ++it) __syncthreads();
            return;
        }
        // the last chunk of a range may hold fewer than NGC groups (max_disparity 256: 65 = 7 x 9 + 2): its walks fill fewer warps
        const int u = warp * 32 + lane;
        const bool act = u < RB * nga;
        const int rb = act ? u / nga : 0, gl = act ? u - rb * nga : 0;
        const int nvalid = a.W - (x0 - HALF);                                 // steps of the walk inside the image
        const bool edge = nvalid < C::NSTEP;                                  // uniform: the strip touches x >= W
        const uint32_t one = opaque(a.k65536 >> 16), mone = opaque(0u - (a.k65536 >> 16));   // +1 / -1 in registers: adds on the FMA pipe
        __syncthreads();
        int ringb = 0;
        for (int it = 0; it < nb + 2; ++it) {
            if (it < nb && act && !(WSR_SKIP & 1)) {
                const int tb = it % C::NTILE;
                const uint32_t* Lr = Lrep + tb * LBUF + rb * C::LW;
                const uint32_t* Rr = Ral + tb * RBUF + rb * C::RWT + rext + (NGC - 1 - gl);
                uint2* Hout = Hs + (ringb * RB + rb) * HROW + gl * TWP;
                if (!edge) sad_walk<HALF, TW, false, true, WSR_FMAS>(Lr, Rr, Hout, nvalid, true, one, mone);
                else       sad_walk<HALF, TW, true, true, WSR_FMAS>(Lr, Rr, Hout, nvalid, true, one, mone);
            }
            ringb = ringb + 1 == C::NB ? 0 : ringb + 1;
            __syncthreads();
        }
    } else if (warp < C::W_CONS) {
        // ======================= service warpgroup: loader (and the finisher when warpgroup 0 is full) =======================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(C::REGS_SERVICE));
        if (!C::FIN3 && warp == C::W_FIN) {
            __syncthreads();
            finisher(IC0{}, ICN{});
        } else if (warp == C::W_LOAD) {
            uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + C::OFF_MBAR);
            const uint32_t mbar0 = (uint32_t)__cvta_generic_to_shared(mbar);
            auto request = [&](int batch) {                 // two bulk tensor copies, completion counted on the buffer's mbarrier
                const int tb = batch % C::NTILE;
                const uint32_t bar = mbar0 + 8 * tb;
                const uint32_t dstR = (uint32_t)__cvta_generic_to_shared(smem + C::OFF_R + tb * C::R_BYTES);
                const uint32_t dstL = (uint32_t)__cvta_generic_to_shared(smem + C::OFF_LRAW + tb * C::LRAW_BYTES);
                const int y = r0 + batch * RB;
                if (lane == 0) {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(RB * C::RWT * 4 + RB * C::LBOX) : "memory");
                    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                                 :: "r"(dstR), "l"(&a.tmapR), "r"(xr0a), "r"(y), "r"(frame), "r"(bar) : "memory");
                    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                                 :: "r"(dstL), "l"(&a.tmapL), "r"(x0 - HALF - C::LSH), "r"(y), "r"(frame), "r"(bar) : "memory");
                }
            };
            auto complete = [&](int batch) {                // wait for the copies of `batch`, then replicate its left pixels
                const int tb = batch % C::NTILE;
                const uint32_t bar = mbar0 + 8 * tb;
                const uint32_t parity = (uint32_t)(batch / C::NTILE) & 1u;
                uint32_t done = 0;
                for (int spin = 0; spin < (1 << 24) && !done; ++spin)          // a stuck copy traps instead of hanging
                    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
                if (!done) __trap();
                const uint32_t* raw = reinterpret_cast<const uint32_t*>(smem + C::OFF_LRAW + tb * C::LRAW_BYTES);
                uint4* Ld = reinterpret_cast<uint4*>(Lrep + tb * LBUF);
                constexpr int LQ = C::LW / 4, NIT = (RB * LQ + 31) / 32;
#pragma unroll
                for (int k0 = 0; k0 < NIT; k0 += 4) {                // four steps at a time, all loads first
                    uint32_t v[4], v2[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int idx = lane + 32 * (k0 + k);
                        v[k] = 0; v2[k] = 0;
                        if (k0 + k < NIT && idx < RB * LQ) {
                            const int rb = idx / LQ, q = idx - rb * LQ;
                            const uint32_t* p = raw + rb * (C::LBOX / 4) + (C::LSH >> 2) + q;
                            v[k] = p[0];
                            if (C::LSH & 3) v2[k] = p[1];
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int idx = lane + 32 * (k0 + k);
                        if (k0 + k < NIT && idx < RB * LQ) {
                            const uint32_t w = (C::LSH & 3) ? __funnelshift_r(v[k], v2[k], 8 * (C::LSH & 3)) : v[k];
                            Ld[idx] = make_uint4(__byte_perm(w, 0u, 0x0000), __byte_perm(w, 0u, 0x1111), __byte_perm(w, 0u, 0x2222), __byte_perm(w, 0u, 0x3333));
                        }
                    }
                }
            };
            if (lane == 0) {
#pragma unroll
                for (int t = 0; t < C::NTILE; ++t) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mbar0 + 8 * t));
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            __syncwarp();
            request(0);
            if (nb > 1) request(1);
            if (nb > 2) request(2);
            complete(0);
            if (nb > 1) complete(1);
            __syncthreads();
            for (int it = 0; it < nb + 2; ++it) {
                if (!(WSR_SKIP & 128)) {
                if (it + 3 < nb) request(it + 3);              // buffer (it+3)%4 was last read in iteration it-1
                if (it + 2 < nb) complete(it + 2);             // requested one iteration ago: already landed
                }
                __syncthreads();
            }
        } else {
            __syncthreads();
            for (int it = 0; it < nb + 2; ++it) __syncthreads();
        }
    } else {
        // ======================= consumers: warp k owns group k for 32 columns =======================
        const int kB = warp - C::W_CONS;
        int grp = wsr_group(NGC, 0), ngw = wsr_ngroups(NGC, 0);
#pragma unroll
        for (int k = 1; k < C::K; ++k)
            if (k == kB) { grp = wsr_group(NGC, k); ngw = wsr_ngroups(NGC, k); }
        __syncthreads();
        if (grp >= nga) grp = -1;
        else if (grp + ngw > nga) ngw = nga - grp;
        if (grp < 0) {
            // the consumer warps without a group (one on each of the sub-partitions 0..2) finish a third of the rows each
            if (C::FIN3 && kB == 8) finisher(IC0{}, IC1{});
            else if (C::FIN3 && kB == 9) finisher(IC1{}, IC2{});
            else if (C::FIN3 && kB == 10) finisher(IC2{}, ICN{});
            else for (int it = 0; it < nb + 2; ++it) __syncthreads();
            return;
        }
        const int xB = x0 + lane;
        const int dmax = min(a.D, xB - HALF);                 // largest evaluated disparity of this column (sad.go:64-67, :212-218)
        const uint32_t k512 = opaque(a.k65536 >> 7), m25 = opaque(0u - (a.k65536 << 9));    // keys sum*512 + d: 512 and -2^25
        const uint2* Hbase = Hs + grp * TWP + lane;
        uint32_t* pkbase = pk + lane;
        // NG groups (grp, grp + 1, ...) of this warp
        auto consume = [&](auto ng_c) {
            constexpr int NG = decltype(ng_c)::value;
            // rows whose loads are issued together (register budget: 4 registers per row and group)
            constexpr int RH = NG == 1 ? (RB < 12 ? RB : 9) : NG == 2 ? ((RB + 1) / 2 < 6 ? (RB + 1) / 2 : 6) : 4;
            // 32-bit sums: VE / VO are the RAW packed sums (low lane + 65536 * high lane, mod 2^32), V3 / V2 the high-lane sums
            // alone; a never-evaluated candidate starts 2^22 above every real sum (245 055), so its key carries 2^31
            uint32_t VE[NG], V3[NG], VO[NG], V2[NG], c3[NG], c1[NG], c2[NG], c0[NG];
#pragma unroll
            for (int j = 0; j < NG; ++j) {
                const int dbase = 4 * (g0 + grp + j);
                VE[j] = dbase + 3 > dmax ? 1u << 22 : 0u; V3[j] = dbase + 1 > dmax ? 1u << 22 : 0u;
                VO[j] = dbase + 2 > dmax ? 1u << 22 : 0u; V2[j] = dbase + 0 > dmax ? 1u << 22 : 0u;
                // the absolute disparities ride in the addends of the key multiply-adds (registers)
                c3[j] = opaque((uint32_t)dbase + 3u); c1[j] = opaque((uint32_t)dbase + 1u);
                c2[j] = opaque((uint32_t)dbase + 2u); c0[j] = opaque((uint32_t)dbase);
            }
            int sn = 0, so1 = C::NB - C::Q, so2 = C::NB - C::Q - 1;             // ring batch of the rows entering / leaving the window
            for (int it = 0; it < nb + 2; ++it) {
                if (it >= 1 && it <= nb && !(WSR_SKIP & 2)) {
                    const int batch = it - 1;
                    const uint2* Hn = Hbase + sn * RB * HROW;
                    const uint2* Ho1 = Hbase + so1 * RB * HROW;
                    const uint2* Ho2 = Hbase + so2 * RB * HROW;
                    uint32_t* pkb = pkbase + (batch & 1) * PKBUF;
#pragma unroll
                    for (int rh = 0; rh < RB; rh += RH) {
                        // all loads first: a shared-memory load cannot be hoisted over the atomicMin of the row before it
                        uint2 n[RH][NG], o[RH][NG];
#pragma unroll
                        for (int r = 0; r < RH; ++r) {
                            const int rb = rh + r;
                            if (rb >= RB) continue;
#pragma unroll
                            for (int j = 0; j < NG; ++j) {
                                n[r][j] = Hn[rb * HROW + j * TWP];
                                o[r][j] = rb >= C::M ? Ho1[(rb - C::M) * HROW + j * TWP] : Ho2[(rb - C::M + RB) * HROW + j * TWP];
                            }
                        }
#pragma unroll
                        for (int r = 0; r < RH; ++r) {
                            if (rh + r >= RB) continue;
                            uint32_t best = 0xFFFFFFFFu;
#pragma unroll
                            for (int j = 0; j < NG; ++j) {
                                // packed difference with the low lane biased by 0x8000: it never borrows from the high lane, so the
                                // arithmetic shift yields the signed high-lane difference.  The copies are opaque so that the raw
                                // sum and the biased difference stay ONE three-input add each (no shared n - o).
                                const uint32_t nx = opaque(n[r][j].x), ny = opaque(n[r][j].y);
                                const uint32_t dE = nx - o[r][j].x + 0x8000u, dO = ny - o[r][j].y + 0x8000u;
                                VE[j] = VE[j] + n[r][j].x - o[r][j].x; V3[j] += (uint32_t)((int)dE >> 16);
                                VO[j] = VO[j] + n[r][j].y - o[r][j].y; V2[j] += (uint32_t)((int)dO >> 16);
                                const uint32_t kEl = VE[j] * k512 + (V3[j] * m25 + c3[j]), kEh = V3[j] * k512 + c1[j];
                                const uint32_t kOl = VO[j] * k512 + (V2[j] * m25 + c2[j]), kOh = V2[j] * k512 + c0[j];
                                best = min(best, min(min(kEl, kEh), min(kOl, kOh)));
                            }
                            atomicMin(pkb + (rh + r) * TW, best);  // rows that are not output rows are filtered by the finisher
                        }
                    }
                    sn = sn + 1 == C::NB ? 0 : sn + 1;
                    so1 = so1 + 1 == C::NB ? 0 : so1 + 1;
                    so2 = so2 + 1 == C::NB ? 0 : so2 + 1;
                }
                __syncthreads();
            }
        };
        if (NGC == 13 && ngw == 3) consume(std::integral_constant<int, 3>{});
        else if (ngw == 2) consume(std::integral_constant<int, 2>{});
        else          consume(std::integral_constant<int, 1>{});
    }
}

}  // namespace sadgpu