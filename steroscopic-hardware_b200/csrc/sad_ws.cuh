// sad_ws.cuh — warp-specialised, double-buffered variant of the fast path (block_size <= 9, chunks of
// 33 disparity groups = 132 disparity slots, i.e. max_disparity 65..128 in one chunk, up to 256 in two).
//
// Same arithmetic as sad_fast.cuh; the difference is scheduling.  A CTA owns a 32-column strip and has
// 24 warps with FIXED roles (no phase alternation, one __syncthreads per 9-row batch), registers rebalanced
// between the roles with setmaxnreg:
//   warps 0..8   walkers: warp w walks row w of the batch for 32 disparity groups (lanes = groups),
//                horizontal running window sums -> H[buf][row][group][column] in shared memory;
//   warp  9      walks the 33rd group (the single candidate d = 128 when D = 128), one lane per row;
//                each walker also prefetches its row of the batch after next (three tile buffers);
//   warp  10     finisher: min over the 11 partial keys of a pixel, d*255/D LUT, store;
//   warps 12..22 consumers: warp 12+k owns groups 3k..3k+2 for 32 columns (lanes = columns): vertical
//                running sums with the previous 2h+1 rows in a register ring, key-min argmin, partial
//                best -> pk; then the cross-warp min + LUT + store of the batch before.
// Producers work on batch i while consumers work on batch i-1 (H and tiles are double-buffered).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "sad_fast.cuh"

namespace sadgpu {

template <int HALF> struct WsCfg {
    static constexpr int WIN = 2 * HALF + 1;
    static constexpr int TW = 32, TWP = 33;
    static constexpr int NSTEP = TW + 2 * HALF;
    static constexpr int LW = (NSTEP + 3) & ~3;
    static constexpr int RB = 9;                       // rows per batch = row-walker warps = ring length
    static constexpr int NGC = 33, GT = 3, K = 11;     // groups per chunk, groups per consumer thread, consumer warps
    // warp roles (6 warpgroups of 4 warps): producers = warps 0..11, consumers = warps 12..23
    static constexpr int W_TAIL = RB;                  // warp 9: 33rd group, one lane per row
    static constexpr int W_FIN = 10;                   // warp 10: cross-warp min + LUT + store (warp 11 idles)
    static constexpr int NTILE = 4;                    // tile buffers: tiles are requested three batches ahead, completed two ahead
    static constexpr int W_CONS = 12;                  // warps 12..22: consumer k = warp-12 (warp 23 idles)
    static constexpr int NT = 768;
    static constexpr int REGS_LAUNCH = 80, REGS_PROD = 56, REGS_CONS = 104;
    static constexpr int REGS_PROD_TMA = 40, REGS_CONS_TMA = 120;           // with TMA the walkers carry no prefetch state   // setmaxnreg moves registers inside the CTA's launch allocation
    static constexpr int OFF = ((-(HALF + 3)) % 4 + 4) % 4;
    static constexpr int NWALKW = ((NSTEP - 1 + OFF) >> 2) + 2;
    static constexpr int RW = NGC - 1 + NWALKW;
    static constexpr int H_BYTES = ((RB * NGC * TWP * 8 + 15) / 16) * 16;      // one buffer
    static constexpr int L_BYTES = RB * LW * 4;
    // TMA needs the innermost start coordinate on a 16-byte boundary: the right tile starts up to 12 bytes early (RWT words
    // per row), the raw left tile LSH bytes early (x0 is a multiple of 32, so LSH only depends on h).
    static constexpr int RWT = ((RW * 4 + 12 + 15) / 16) * 4;
    static constexpr int R_BYTES = ((RB * RWT * 4 + 127) / 128) * 128;        // 128-byte multiple: each buffer is a TMA destination
    static constexpr int LSH = (16 - HALF % 16) % 16;
    static constexpr int LBOX = ((LSH + NSTEP + 15) / 16) * 16;               // TMA box width of the raw left tile (bytes)
    static constexpr int LRAW_BYTES = ((RB * LBOX + 127) / 128) * 128;
    static constexpr int PK_BYTES = RB * K * TW * 4;
    static constexpr int OFF_R = ((2 * H_BYTES + 127) / 128) * 128;            // TMA destinations first (128-byte aligned)
    static constexpr int OFF_LRAW = OFF_R + NTILE * R_BYTES;
    static constexpr int OFF_L = OFF_LRAW + NTILE * LRAW_BYTES;
    static constexpr int OFF_PK = OFF_L + NTILE * L_BYTES;
    static constexpr int OFF_LUT = OFF_PK + 2 * PK_BYTES;
    static constexpr int OFF_MBAR = OFF_LUT + 1040;
    static constexpr int SMEM = OFF_MBAR + 64;
    static_assert(WIN <= RB, "register ring shorter than the window");
    static_assert(NT * REGS_LAUNCH <= 65536 && 384 * REGS_PROD + 384 * REGS_CONS <= NT * REGS_LAUNCH, "register budget");
    static_assert(GT * K == NGC && W_CONS + K <= 24, "warp roles");
};

// One (row, group) walk: TW outputs, NSTEP steps, fully unrolled (see fast_walk in sad_fast.cuh).
template <int HALF, bool EDGE>
__device__ __forceinline__ void ws_walk(const uint32_t* __restrict__ Lr, const uint32_t* __restrict__ Rr,
                                        uint2* __restrict__ Hout, int nvalid)
{
    using T = WsCfg<HALF>;
    uint32_t e[T::NSTEP], o[T::NSTEP];
    uint32_t hE = 0, hO = 0, w0 = 0, w1 = 0;
    uint4 lv = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int i = 0; i < T::NSTEP; ++i) {
        if ((i & 3) == 0) lv = *reinterpret_cast<const uint4*>(Lr + i);
        const int bi = i + T::OFF;
        if (i == 0) { w0 = Rr[bi >> 2]; w1 = Rr[(bi >> 2) + 1]; }
        else if ((bi & 3) == 0) { w0 = w1; w1 = Rr[(bi >> 2) + 1]; }
        const uint32_t lw = (i & 3) == 0 ? lv.x : (i & 3) == 1 ? lv.y : (i & 3) == 2 ? lv.z : lv.w;
        const uint32_t rw = (bi & 3) == 0 ? w0 : __funnelshift_r(w0, w1, 8 * (bi & 3));
        uint32_t ad = __vabsdiffu4(lw, rw);
        if (EDGE) ad = (i < nvalid) ? ad : 0u;
        e[i] = __byte_perm(ad, 0u, 0x4240);
        o[i] = __byte_perm(ad, 0u, 0x4341);
        if (i >= T::WIN) { hE = hE + e[i] - e[i - T::WIN]; hO = hO + o[i] - o[i - T::WIN]; }
        else             { hE += e[i]; hO += o[i]; }
        if (i >= 2 * HALF) Hout[i - 2 * HALF] = make_uint2(hE, hO);
    }
}

// Consumer warp: NGB groups (20 or 12 disparities) x 32 columns; one barrier per batch.
template <int HALF, int NGB>
__device__ __forceinline__ void ws_consume(const FastArgs& a, const uint2* __restrict__ Hs, uint32_t* __restrict__ pk,
                                           int kB, int lane, int x0, int g0, int r0, int nb)
{
    using C = WsCfg<HALF>;
    constexpr int WIN = C::WIN, TW = C::TW, TWP = C::TWP, RB = C::RB, GT = C::GT, K = C::K, NGC = C::NGC;
    constexpr int HBUF = C::H_BYTES / 8, PKBUF = RB * K * TW;
    const int xB = x0 + lane;
    uint32_t VE[NGB], VO[NGB], ringE[RB][NGB], ringO[RB][NGB];
#pragma unroll
    for (int j = 0; j < NGB; ++j) {
        const int dbase = 4 * (g0 + kB * GT + j);
        const int dmax = min(a.D, xB - HALF);
        const uint32_t iE = (dbase + 3 > dmax ? 0x0000FFFFu : 0u) | (dbase + 1 > dmax ? 0xFFFF0000u : 0u);
        const uint32_t iO = (dbase + 2 > dmax ? 0x0000FFFFu : 0u) | (dbase + 0 > dmax ? 0xFFFF0000u : 0u);
        VE[j] = iE & 0x80008000u;                             // bias: never-evaluated candidates lose
        VO[j] = iO & 0x80008000u;
#pragma unroll
        for (int r = 0; r < RB; ++r) { ringE[r][j] = 0; ringO[r][j] = 0; }
    }
    const uint32_t keybase = 4u * (uint32_t)(g0 + kB * GT);
    const uint32_t k16 = opaque(a.k65536), mhi = opaque(a.k65536 * 0xFFFFu);
    long long tw = 0, tb = 0, cprev = 0;
    // unrolled by two so that the ring registers can alternate roles across iterations (the value loaded for row rb is
    // the "old" value of the next batch): without it every ring update costs a register move
#pragma unroll 2
    for (int it = 0; it < nb + 2; ++it) {
        if (a.debug_skip & 4) { const volatile uint32_t* vq = pk; tw += (long long)(vq[0] & 0u); }       // forces the deferred barrier wait to complete
        const long long c0 = clock64();
        if ((a.debug_skip & 4) && it > 0) tb += c0 - cprev;
        if (it >= 1 && it <= nb && (a.debug_skip & 3) != 2) {
            const int batch = it - 1;
            const uint2* Hp = Hs + (batch & 1) * HBUF + (kB * GT) * TWP + lane;
            uint32_t* pkb = pk + (batch & 1) * PKBUF + kB * TW + lane;
#pragma unroll
            for (int rb = 0; rb < RB; ++rb) {
                uint32_t best = 0xFFFFFFFFu;
#pragma unroll
                for (int j = 0; j < NGB; ++j) {
                    const uint2 n = Hp[(rb * NGC + j) * TWP];
                    VE[j] = VE[j] + n.x - ringE[(rb + RB - WIN) % RB][j];
                    VO[j] = VO[j] + n.y - ringO[(rb + RB - WIN) % RB][j];
                    ringE[rb][j] = n.x; ringO[rb][j] = n.y;
                    const uint32_t kEl = key_lo(VE[j], k16, 4u * j + 3u);
                    const uint32_t kEh = key_hi(VE[j], mhi, 4u * j + 1u);
                    const uint32_t kOl = key_lo(VO[j], k16, 4u * j + 2u);
                    const uint32_t kOh = key_hi(VO[j], mhi, 4u * j + 0u);
                    best = min(best, min(kEl, kEh));
                    best = min(best, min(kOl, kOh));
                }
                pkb[rb * K * TW] = best + keybase;            // rows that are not output rows are filtered by the finisher
            }
        }
        const long long c1 = clock64();
        __syncthreads();
        if (a.debug_skip & 4) { tw += c1 - c0; cprev = c1; }
    }
    if ((a.debug_skip & 4) && lane == 0 && blockIdx.x == 7 && blockIdx.y == 0 && blockIdx.z == 0) {
        const int warp = kB + C::W_CONS;
        a.gkey[warp * 4 + 0] = (uint32_t)tw; a.gkey[warp * 4 + 1] = 0; a.gkey[warp * 4 + 2] = (uint32_t)tb; a.gkey[warp * 4 + 3] = nb;
    }
}

template <int HALF, bool TMA>
__global__ void __launch_bounds__(WsCfg<HALF>::NT, 1) sad_ws_kernel(const __grid_constant__ FastArgs a)
{
    using C = WsCfg<HALF>;
    constexpr int WIN = C::WIN, TW = C::TW, TWP = C::TWP, RB = C::RB, GT = C::GT, K = C::K, NGC = C::NGC;
    extern __shared__ __align__(128) unsigned char smem[];
    uint2* Hs = reinterpret_cast<uint2*>(smem);                                   // [2][RB][NGC][TWP]
    uint32_t* Lrep = reinterpret_cast<uint32_t*>(smem + C::OFF_L);               // [2][RB][LW]
    uint32_t* Ral = reinterpret_cast<uint32_t*>(smem + C::OFF_R);                // [2][RB][RW]
    uint32_t* pk = reinterpret_cast<uint32_t*>(smem + C::OFF_PK);                // [2][RB][K][TW]
    uint8_t* lut = smem + C::OFF_LUT;
    constexpr int HBUF = C::H_BYTES / 8, LBUF = RB * C::LW, RBUF = C::R_BYTES / 4, PKBUF = RB * K * TW;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int frame = blockIdx.z / a.NC, chunk = blockIdx.z - frame * a.NC;
    const int x0 = blockIdx.x * TW;
    const int yb0 = a.y0 + blockIdx.y * a.BH;
    const int yb1 = min(a.y1, yb0 + a.BH);
    const int g0 = chunk * NGC;
    if (yb0 >= yb1) return;
    // a chunk none of whose disparities is a candidate anywhere in this strip (d > X-h for every column, sad.go:64-67 + :212-218)
    // has nothing to contribute: chunk 0 always runs and writes every pixel
    if (g0 > 0 && min(x0 + TW, a.W) - 1 - HALF < 4 * g0) return;
    const int r0 = yb0 - HALF;
    const int nb = ((yb1 - yb0) + 2 * HALF + RB - 1) / RB;
    const int nvalid = a.W - (x0 - HALF);

    for (int d = tid; d < 1040; d += C::NT) lut[d] = d <= a.D ? (uint8_t)((d * 255) / a.D) : 0;

    if (warp < C::W_CONS) {
        // ======================= producer warpgroups (warps 0..11) =======================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(TMA ? C::REGS_PROD_TMA : C::REGS_PROD));
        if (warp == C::W_FIN) {
            // ---- finisher: min over the K partial keys of a pixel, LUT, store (batch it-2) ----
            uint8_t* __restrict__ Og = a.out + (long long)frame * a.frameOut;
            __syncthreads();
            for (int it = 0; it < nb + 2; ++it) {
                if (it >= 2) {
                    const int batch = it - 2;
                    const uint32_t* pkb = pk + (batch & 1) * PKBUF + lane;
                    const int x = x0 + lane;
#pragma unroll
                    for (int rb = 0; rb < RB; ++rb) {
                        const int rel = batch * RB + rb, y = r0 + rel - HALF;
                        if (rel < 2 * HALF || y >= yb1 || x >= a.W) continue;
                        uint32_t best = 0xFFFFFFFFu;
#pragma unroll
                        for (int k = 0; k < K; ++k) best = min(best, pkb[(rb * K + k) * TW]);
                        if (x < HALF) best = 0;                  // sad.go:212-218: both windows clamp, d = 0 wins
                        if (a.NC == 1) Og[(size_t)y * a.pitchOut + x] = lut[best & 0xFFFFu];
                        else atomicMin(a.gkey + ((size_t)frame * a.H + y) * a.W + x, ((best >> 16) << 9) | (best & 511u));
                    }
                }
                __syncthreads();
            }
        } else if (warp > C::W_FIN) {
            // ---- warp 11: TMA tile loader (a.use_tma) — two cp.async.bulk.tensor per batch (raw left rows, aligned right
            //      rows; hardware zero-fill outside the image), completion on an mbarrier, then the left pixels are
            //      replicated into Lrep.  Without TMA this warp idles and the walkers prefetch their own rows. ----
            if (TMA) {
                uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + C::OFF_MBAR);
                const uint32_t mbar0 = (uint32_t)__cvta_generic_to_shared(mbar);
                const int xr0 = x0 - HALF - 3 - 4 * (g0 + NGC - 1) - C::OFF;
                const int xr0a = xr0 - (((xr0 % 16) + 16) % 16);                // 16-byte aligned start of the right tile
                if (lane == 0) {
#pragma unroll
                    for (int t = 0; t < C::NTILE; ++t) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mbar0 + 8 * t));
                    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                }
                __syncwarp();
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
                    // bounded wait on the phase of this buffer's (batch / NTILE)-th use; a stuck copy traps instead of hanging
                    const uint32_t parity = (uint32_t)(batch / C::NTILE) & 1u;
                    uint32_t done = 0;
                    for (int spin = 0; spin < (1 << 24) && !done; ++spin)
                        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
                    if (!done) __trap();
                    const uint8_t* raw = smem + C::OFF_LRAW + tb * C::LRAW_BYTES;
                    uint32_t* Ld = Lrep + tb * LBUF;
                    for (int idx = lane; idx < RB * C::LW; idx += 32) {
                        const int rb = idx / C::LW, i = idx - rb * C::LW;
                        Ld[idx] = (uint32_t)raw[rb * C::LBOX + C::LSH + i] * 0x01010101u;
                    }
                };
                request(0);
                if (nb > 1) request(1);
                if (nb > 2) request(2);
                complete(0);
                if (nb > 1) complete(1);
                __syncthreads();
                for (int it = 0; it < nb + 2; ++it) {
                    if (it + 3 < nb) request(it + 3);              // buffer (it+3)%4 was last read in iteration it-1
                    if (it + 2 < nb) complete(it + 2);             // requested one iteration ago: already landed
                    __syncthreads();
                }
            } else {
                __syncthreads();
                for (int it = 0; it < nb + 2; ++it) __syncthreads();
            }
        } else {
            // ---- walkers: warp w < 9 walks row w for groups 0..31 and prefetches row w of the batch after next
            //      (L pixels replicated, R as aligned words) into the third tile buffer; warp 9 walks group 32
            //      of every row (one lane per row). ----
            const bool tail = warp == C::W_TAIL;
            const int rb = tail ? lane : warp, gl = tail ? NGC - 1 : lane;
            const bool act = !tail || lane < RB;
            const uint8_t* __restrict__ Lg = a.L + (long long)frame * a.frameL;
            const uint8_t* __restrict__ Rg = a.R + (long long)frame * a.frameR;
            const int xr0 = x0 - HALF - 3 - 4 * (g0 + NGC - 1) - C::OFF;
            const int xr0a = xr0 - (((xr0 % 16) + 16) % 16);                    // tile rows start 16-byte aligned (TMA rule), same layout without TMA
            const int rext = (xr0 - xr0a) >> 2;                                 // words to skip at the start of a tile row
            constexpr int NLQ = (C::LW + 31) / 32, NRQ = (C::RWT + 31) / 32;
            int lx[NLQ], rx[NRQ], rmode[NRQ];            // column of each slot of this lane; -1 / mode 0 = zero
#pragma unroll
            for (int q = 0; q < NLQ; ++q) {
                const int i = lane + 32 * q, x = x0 - HALF + i;
                lx[q] = (i < C::LW && (unsigned)x < (unsigned)a.W) ? x : -1;
            }
#pragma unroll
            for (int q = 0; q < NRQ; ++q) {
                const int j = lane + 32 * q, x = xr0a + 4 * j;
                const bool in = j < C::RWT && x + 3 >= 0 && x < a.W;
                rx[q] = x;
                rmode[q] = !in ? 0 : (a.aligned && x >= 0 && x + 3 < a.W) ? 1 : 2;
            }
            uint32_t vl[NLQ], vr[NRQ];
            auto issue = [&](int batch) {               // global loads of row rb of `batch` (warp-uniform row test)
                const int y = r0 + batch * RB + rb;
                const bool yin = (unsigned)y < (unsigned)a.H;
                const uint8_t* pl = Lg + (size_t)(yin ? y : 0) * a.pitchL;
                const uint8_t* pr = Rg + (size_t)(yin ? y : 0) * a.pitchR;
#pragma unroll
                for (int q = 0; q < NLQ; ++q) { vl[q] = 0; if (yin && lx[q] >= 0) vl[q] = pl[lx[q]]; }
#pragma unroll
                for (int q = 0; q < NRQ; ++q) {
                    uint32_t v = 0;
                    if (yin && rmode[q] == 1) v = *reinterpret_cast<const uint32_t*>(pr + rx[q]);
                    else if (yin && rmode[q] == 2) {
#pragma unroll
                        for (int b = 0; b < 4; ++b)
                            if ((unsigned)(rx[q] + b) < (unsigned)a.W) v |= (uint32_t)pr[rx[q] + b] << (8 * b);
                    }
                    vr[q] = v;
                }
            };
            auto commit = [&](int batch) {
                uint32_t* Ld = Lrep + (batch % C::NTILE) * LBUF + rb * C::LW;
                uint32_t* Rd = Ral + (batch % C::NTILE) * RBUF + rb * C::RWT;
#pragma unroll
                for (int q = 0; q < NLQ; ++q) { const int i = lane + 32 * q; if (i < C::LW) Ld[i] = vl[q] * 0x01010101u; }
#pragma unroll
                for (int q = 0; q < NRQ; ++q) { const int j = lane + 32 * q; if (j < C::RWT) Rd[j] = vr[q]; }
            };
            const bool self_load = !tail && !TMA;
            if (self_load) {
                issue(0); commit(0);
                if (nb > 1) { issue(1); commit(1); }
            }
            __syncthreads();
            long long tw = 0, tc = 0, tb = 0, cprev = 0;
            for (int it = 0; it < nb + 2; ++it) {
                if (a.debug_skip & 4) { volatile uint8_t* vq = lut; tw += (long long)(vq[0] & 0) ; }   // forces the deferred barrier wait to complete
                const long long c0 = clock64();
                if ((a.debug_skip & 4) && it > 0) tb += c0 - cprev;
                const bool pre = self_load && it + 2 < nb;
                if (pre) issue(it + 2);
                if (it < nb && act && (a.debug_skip & 3) != 1) {
                    const int buf = it & 1, tb = it % C::NTILE;
                    const uint32_t* Lr = Lrep + tb * LBUF + rb * C::LW;
                    const uint32_t* Rr = Ral + tb * RBUF + rb * C::RWT + rext + (NGC - 1 - gl);
                    uint2* Hout = Hs + buf * HBUF + (rb * NGC + gl) * TWP;
                    if (nvalid >= C::NSTEP) ws_walk<HALF, false>(Lr, Rr, Hout, nvalid);
                    else                    ws_walk<HALF, true>(Lr, Rr, Hout, nvalid);
                }
                const long long c1 = clock64();
                if (pre) commit(it + 2);               // tile buffer (it+2)%3 was last read in iteration it-1
                const long long c2 = clock64();
                __syncthreads();
                if (a.debug_skip & 4) { tw += c1 - c0; tc += c2 - c1; cprev = c2; }
            }
            if ((a.debug_skip & 4) && lane == 0 && blockIdx.x == 7 && blockIdx.y == 0 && blockIdx.z == 0) {
                a.gkey[warp * 4 + 0] = (uint32_t)tw; a.gkey[warp * 4 + 1] = (uint32_t)tc; a.gkey[warp * 4 + 2] = (uint32_t)tb; a.gkey[warp * 4 + 3] = nb;
            }
        }
    } else {
        // ======================= consumer warpgroups (warps 12..23) =======================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(TMA ? C::REGS_CONS_TMA : C::REGS_CONS));
        __syncthreads();
        const int kB = warp - C::W_CONS;
        if (kB < K) {
            // ---- consumers: vertical running sums (register ring) + argmin keys for GT groups x 32 columns ----
            ws_consume<HALF, GT>(a, Hs, pk, kB, lane, x0, g0, r0, nb);
        } else {
            for (int it = 0; it < nb + 2; ++it) __syncthreads();       // spare warp of the consumer register class
        }
    }
}

}  // namespace sadgpu
