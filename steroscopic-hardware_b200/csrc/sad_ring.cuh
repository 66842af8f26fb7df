// sad_ring.cuh — warp-specialised kernel for the mid-size windows (block_size 11..15, h = 5..7) where the 2h+1
// previous rows of horizontal sums neither fit a register ring (sad_ws.cuh) nor need 32-bit arithmetic.
//
// Same arithmetic and the same two passes as sad_ws.cuh (pkg/despair/sad.go:55-95, :205-244 through the separable box
// filter): row walkers produce the horizontal window sums H(row, group, column), column consumers keep the vertical
// running sums V += H(y+h) - H(y-h-1) and the key-min argmin.  The difference is where the leaving row comes from: all
// H rows of the window live in a SHARED-MEMORY RING of 2h+2+NWK rows (24 rows x 8.5 KB at h = 7) and the consumers
// read the entering and the leaving row (two LDS.64 per 4 candidates).  A ring that deep leaves no room for double
// buffering by batches, so every hand-over is row-granular and goes through an mbarrier:
//   loaders --tile_full--> walkers / tail walker --h_full--> consumers --pk_full--> finisher
//           <--tile_empty--                      <--h_empty--          <--pk_empty--
// 24 warps with fixed roles: 8 walkers (one row each, lanes = 32 disparity groups), 1 tail walker (33rd group of four
// rows, lanes = row x column segment), 11 consumers (3 groups x 32 columns, four rows per step), 1 finisher, 2 loaders.
// Candidates the reference never evaluates (d > min(D, X-h)) lose through per-thread key constants (multiplier 0 /
// all-ones addend): no instruction is spent on them.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "sad_fast.cuh"

#ifndef RING_WIDE_GT
#define RING_WIDE_GT 2          // groups per consumer thread of the 17-group instances (3 -> 6 consumer warps, 2 -> 9)
#endif

namespace sadgpu {

#ifndef RING_SPLIT8_MIN_HALF
#define RING_SPLIT8_MIN_HALF 15   // from this h on, bursts of 8 rows are walked by eight warps (half-warps = column halves) instead of four
                                  // (half-warps = rows): +48 % walker instructions but half the walk latency.  Measured: h = 15 5 % faster
                                  // (216 -> 205 us at B=31 D=128), h = 8..12 unchanged or 1 % slower
#endif
#ifndef RING_WIDE_SLACK
#define RING_WIDE_SLACK 0         // extra bursts of ring slack in the 17-group instances where shared memory allows
#endif
#ifndef RING_PKATOM
#define RING_PKATOM 1
#endif
#ifndef RING_BURST6
#define RING_BURST6 0             // experiment, off: h >= 12 with bursts of 6 rows, unsplit walks by three warps and a row-granular
#endif                            // ring release.  Bit-exact but slower (B=31 D=128: 346 vs 253 us): 42 rows leave no slack at h = 15
                                  // (walkers and consumers wait for each other) and three walker warps are latency-bound.
#ifndef RING_UNSPLIT4
#define RING_UNSPLIT4 0           // experiment, off: bursts of 4 rows walked unsplit by two warps (lanes = 16 groups x 2 rows):
#endif                            // two walker warps are latency-bound (B=31 D=128: 336 vs 253 us)
#ifndef RING_BURST8_MAX_HALF
#define RING_BURST8_MAX_HALF 15   // largest h whose ring (window + two bursts of 8 rows of 17 groups) fits shared memory (11 without RING_PKATOM)
#endif
constexpr bool SMEM_OK(int bytes) { return bytes + 1024 <= 232448; }

template <int HALF> struct RingCfg {
    static_assert(HALF >= 1 && HALF <= 15, "ring kernel: block_size <= 31");
    static constexpr int WIN = 2 * HALF + 1;
    static constexpr bool WIDE = WIN * WIN * 255 >= 65536;          // h >= 8: window sums need 18 bits -> 32-bit consumer sums
    static constexpr int TW = 32, TWP = 33;
    static constexpr int NSTEP = TW + 2 * HALF;
    static constexpr int LW = (NSTEP + 3) & ~3;
#ifndef RING_PAD
#define RING_PAD 1
#endif
    // h <= 7: chunks of 33 groups (walker lanes = 32 groups); h >= 8: the window needs up to 40 ring rows, so a chunk has 17
    // groups (walker lanes = 16 groups x 2 column halves) and a ring row is half as large
    static constexpr int NGC = WIDE ? 17 : 33, NGL = NGC - 1, GT = WIDE ? RING_WIDE_GT : 3, K = WIDE ? 18 / RING_WIDE_GT : 11;
    static constexpr int NGS = GT * K;                              // group slots of an H row (>= NGC; the surplus slot is never valid)
    // rows per burst.  17-group rows are small enough for bursts of 8 up to h = 11: the walkers then take whole 32-column rows
    // (lanes = 16 groups x 2 rows) instead of re-warming 16-column halves
    // h >= 12: bursts of 6 with the ring released row by row (42 rows hold the 2h+1 rows of the window plus two bursts only if a
    // row is handed back as soon as its last reader is done)
    static constexpr int NWK = !WIDE ? 4 : HALF <= RING_BURST8_MAX_HALF ? 8 : RING_BURST6 ? 6 : 4;
    // 17-group instances: the consumers fold their partial keys into ONE word per pixel with a shared-memory atomicMin instead
    // of K words that the finishers reduce, and an H row holds exactly 17 group slots (the ninth consumer's second slot reads
    // into the next row; its keys are invalid by construction).  That frees 25 KB: room for a ring of 8-row bursts up to h = 15.
    static constexpr bool PKATOM = WIDE && RING_PKATOM;
    static constexpr int PKK = PKATOM ? 1 : K;                      // words per pixel and row in pk
    static constexpr bool ROWREL = NWK == 6;                         // h_empty per row instead of per burst
    static constexpr int WMODE = !WIDE ? 0 : ((NWK == 4 && !RING_UNSPLIT4) || (NWK == 8 && HALF >= RING_SPLIT8_MIN_HALF)) ? 1 : 2;       // 0: warp = (row, column half); 1: warp = row, half-warps = column halves; 2: warp = row pair, half-warps = rows
    // Ring sizes are multiples of the number of warps that take turns on them (8 walkers; 2 loaders x 4 rows), so that a
    // slot is always produced by the same warp: a parity wait is only sound for a waiter that has seen every phase.
    static constexpr int OB = (WIN + NWK - 1) / NWK;                // a burst has left every window OB bursts later
    static constexpr int NB = ROWREL ? 7 : WIDE ? OB + 2 + (HALF <= 11 ? RING_WIDE_SLACK : 0) : 6;   // bursts in the H ring: window + the burst being consumed + the one being written (+ slack)
    static constexpr int NRH = NB * NWK;                            // H ring rows
    static constexpr int TR = NWK == 6 ? 24 : 16;                   // pixel-tile ring rows (a multiple of the burst and of 2 loaders x 4 rows)
    static constexpr int RT = NWK == 8 ? 8 : 4, NSEG = 32 / RT, SEGW = TW / NSEG;  // tail walker: lane = (row of a pass, segment of SEGW columns)
    static constexpr int TPASS = (NWK + RT - 1) / RT;               // tail passes per burst
    static constexpr int OFF = ((-(HALF + 3)) % 4 + 4) % 4;
    static constexpr int NWALKW = ((NSTEP - 1 + OFF) >> 2) + 2;
    static constexpr int RW = NGC - 1 + NWALKW;                     // words per aligned-R tile row
    // tile row strides: with two rows per walker warp (17-group instances) the strides are padded to 4 / 12 mod 32 words so that the
    // loads of different rows fall into different banks (the same padding took the warp-specialised kernel from 78 to 58 us at
    // block 16, max disparity 64)
    static constexpr int LWS = (WIDE && RING_PAD) ? LW + ((4 - LW % 32 + 32) % 32) : LW;
    static constexpr int RWS = (WIDE && RING_PAD) ? RW + ((12 - RW % 32 + 32) % 32) : RW;
    static constexpr int HROW = (PKATOM ? NGC : NGS) * TWP;         // uint2 per H row
    static constexpr int NT = 768;
    // warp roles; SM sub-partition = warp % 4
    static constexpr int W_CONS = 8, W_TAIL = 19, W_FIN = 20, W_LD0 = 21, W_LD1 = 22, W_FIN2 = 23;
    static constexpr int H_BYTES = (NRH * HROW + (NGS - NGC) * TWP) * 8;    // + the surplus slot of the last row
    static constexpr int OFF_L = ((H_BYTES + 15) / 16) * 16;
    static constexpr int OFF_R = OFF_L + TR * LWS * 4;
    static constexpr int OFF_PK = OFF_R + TR * RWS * 4;             // [2][NWK][PKK][TW]
    static constexpr int OFF_LUT = OFF_PK + 2 * NWK * PKK * TW * 4;
    static constexpr int OFF_BAR = ((OFF_LUT + 1040 + 7) / 8) * 8;
    static constexpr int WSPLIT = WMODE == 2 ? 1 : 2;               // column segments of a row walk
    static constexpr int NWW = WMODE == 0 ? NWK * 2 : WMODE == 1 ? NWK : NWK / 2;    // walker warps
    static constexpr int TEC = (WMODE == 0 ? 2 : 1) + 1;            // arrivals that free a pixel-tile row: its walker warps + the tail walker
    static constexpr int OB0 = (2 * HALF) / NWK;                    // first burst that holds an output row
    static constexpr int NHE = ROWREL ? NRH : NB;                   // h_empty barriers
    static constexpr int NBAR = NB + NHE + 2 * TR + 4;
    static constexpr int SMEM = OFF_BAR + NBAR * 8;
    static_assert(NGS >= NGC && RT * NSEG == 32 && SEGW % 4 == 0 && W_CONS + K <= W_TAIL && NWW <= W_CONS, "roles");
    static_assert((NWK == 4 || NWK == 6 || NWK == 8) && NRH % NWK == 0 && TR % 8 == 0 && TR % NWK == 0, "every ring slot belongs to one producer warp");
    static_assert(SMEM_OK(OFF_BAR), "shared memory");
    static_assert(ROWREL ? NRH + 1 >= WIN + 2 * NWK : NB >= OB + 2, "ring: window + the burst being consumed + the burst being written");
};

__device__ __forceinline__ void ring_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
// Bounded wait (try_wait suspends the warp in hardware): a protocol error traps instead of hanging the GPU.
#ifndef RING_SLEEP_NS
#define RING_SLEEP_NS 0          // back-off between two failed try_wait probes (0 = probe again at once)
#endif
#define RING_STR2(x) #x
#define RING_STR(x) RING_STR2(x)
__device__ __forceinline__ void ring_wait_(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .u32 n;\n"
        "mov.u32 n, 0;\n"
        "RING_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "@p bra RING_WAIT_DONE;\n"
#if RING_SLEEP_NS > 0
        "nanosleep.u32 " RING_STR(RING_SLEEP_NS) ";\n"
#endif
        "add.u32 n, n, 1;\n"
        "setp.lt.u32 p, n, 4000000;\n"
        "@p bra RING_WAIT_LOOP;\n"
        "RING_WAIT_DONE:\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok) __trap();
}

#ifdef RING_PROFILE       // developer build: per-warp cycles spent waiting on each barrier family, written for one CTA
#define RING_WAIT(bar, parity, slot) do { const long long c0_ = clock64(); ring_wait_(bar, parity); wt[slot] += clock64() - c0_; } while (0)
#define RING_PROF_BEGIN long long wt[8] = {0, 0, 0, 0, 0, 0, 0, 0}; const long long c00_ = clock64()
#define RING_PROF_END(dbg) do { if (dbg && (threadIdx.x & 31) == 0) { wt[7] = clock64() - c00_; \
    for (int i_ = 0; i_ < 8; ++i_) dbg[(threadIdx.x >> 5) * 8 + i_] = (uint32_t)(wt[i_] >> 4); } } while (0)
#else
#define RING_WAIT(bar, parity, slot) ring_wait_(bar, parity)
#define RING_PROF_BEGIN do {} while (0)
#define RING_PROF_END(dbg) do {} while (0)
#endif

template <int HALF>
__global__ void __launch_bounds__(RingCfg<HALF>::NT, 1) sad_ring_kernel(const __grid_constant__ FastArgs a)
{
    using C = RingCfg<HALF>;
    constexpr int WIN = C::WIN, TW = C::TW, TWP = C::TWP, NGC = C::NGC, GT = C::GT, K = C::K, NRH = C::NRH, TR = C::TR;
    constexpr int HROW = C::HROW, LW = C::LW, RW = C::RW, NWK = C::NWK, NB = C::NB, OB = C::OB, OB0 = C::OB0;
    extern __shared__ __align__(128) unsigned char smem[];
    uint2* Hs = reinterpret_cast<uint2*>(smem);                                   // [NRH][NGC][TWP]
    uint32_t* Lrep = reinterpret_cast<uint32_t*>(smem + C::OFF_L);               // [TR][LW]
    uint32_t* Ral = reinterpret_cast<uint32_t*>(smem + C::OFF_R);                // [TR][RW]
    uint32_t* pk = reinterpret_cast<uint32_t*>(smem + C::OFF_PK);                // [2][4][K][TW]
    uint8_t* lut = smem + C::OFF_LUT;
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(smem + C::OFF_BAR);
    const uint32_t hfull = bar0, hempty = bar0 + 8 * NB, tfull = hempty + 8 * C::NHE, tempty = tfull + 8 * TR;
    const uint32_t pkfull = tempty + 8 * TR, pkempty = pkfull + 16;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int frame = blockIdx.z / a.NC, chunk = blockIdx.z - frame * a.NC;
    const int x0 = blockIdx.x * TW;
    const int yb0 = a.y0 + blockIdx.y * a.BH;
    const int yb1 = min(a.y1, yb0 + a.BH);
    if (yb0 >= yb1) return;
    const int g0 = chunk * NGC;
    // a chunk none of whose disparities is a candidate anywhere in this strip (d > X-h for every column, sad.go:64-67 + :212-218)
    // has nothing to contribute: chunk 0 always runs and writes every pixel
    if (g0 > 0 && min(x0 + TW, a.W) - 1 - HALF < 4 * g0) return;
    const int bhc = yb1 - yb0, nin = bhc + 2 * HALF;          // output rows / H rows of this band (H row r = image row yb0-h+r)
    const int nvalid = a.W - (x0 - HALF);
    const int nbur = (nin + NWK - 1) / NWK;                   // bursts of NWK rows

    for (int d = tid; d < 1040; d += C::NT) lut[d] = d <= a.D ? (uint8_t)((d * 255) / a.D) : 0;
    if (C::PKATOM) for (int i = tid; i < 2 * NWK * TW; i += C::NT) pk[i] = 0xFFFFFFFFu;
    if (tid == 0) {
        for (int i = 0; i < NB; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(hfull + 8 * i), "n"(C::NWW + 1));  // walkers + tail walker
        }
        for (int i = 0; i < C::NHE; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(hempty + 8 * i), "n"(K));
        for (int i = 0; i < TR; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(tfull + 8 * i));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(tempty + 8 * i), "n"(C::TEC));   // walkers + tail walker
        }
        for (int i = 0; i < 2; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(pkfull + 8 * i), "n"(K));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 2;" :: "r"(pkempty + 8 * i));             // two finishers
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    uint32_t* dbg = ((a.debug_skip & 4) && blockIdx.x == 7 && blockIdx.y == 0 && blockIdx.z == 0) ? a.gkey : nullptr;
    (void)dbg;
    RING_PROF_BEGIN;
    if (warp < C::NWW) {
        // ---- walkers (lanes = groups).  33-group chunks: warps w and w+NWK walk the two 16-column halves of row w of every
        //      burst.  17-group chunks: the two half-warps of warp w walk the halves of row w (bursts of 4), or whole rows
        //      2w and 2w+1 (bursts of 8).  A half re-warms its window over 2h columns. ----
        const int hw = lane >> 4;
        const int wr = C::WMODE == 0 ? warp % NWK : C::WMODE == 1 ? warp : 2 * warp + hw;      // row of the burst
        const int wh = C::WMODE == 0 ? warp / NWK : C::WMODE == 1 ? hw : 0;                    // column segment
        const int gl = lane & (C::NGL - 1);
        constexpr int HW = TW / C::WSPLIT;
        for (int bi = 0; bi < nbur; ++bi) {
            const int r = NWK * bi + wr, bs = bi % NB;
            const int ts = r % TR;
            const uint32_t* Lr = Lrep + ts * C::LWS + HW * wh;
            const uint32_t* Rr = Ral + ts * C::RWS + (NGC - 1 - gl) + (HW / 4) * wh;
            uint2* Hout = Hs + (bs * NWK + wr) * HROW + gl * TWP + HW * wh;
            if (C::WMODE != 2) {                               // one row per warp: everything below is warp-uniform
                if (r < nin) {
                    RING_WAIT(tfull + 8 * ts, (uint32_t)(r / TR) & 1u, 0);
                    if (bi >= NB) RING_WAIT(hempty + 8 * bs, (uint32_t)(bi / NB - 1) & 1u, 1);
                    if (nvalid >= C::NSTEP) sad_walk<HALF, HW, false>(Lr, Rr, Hout, nvalid - HW * wh, true);
                    else                    sad_walk<HALF, HW, true>(Lr, Rr, Hout, nvalid - HW * wh, true);
                    __syncwarp();
                    if (lane == 0) ring_arrive(tempty + 8 * ts);
                }
            } else {                                           // one row per half-warp
                const bool act = r < nin;
                if (act) RING_WAIT(tfull + 8 * ts, (uint32_t)(r / TR) & 1u, 0);
                if (C::ROWREL) { if (act && r >= NRH) RING_WAIT(hempty + 8 * (r % NRH), (uint32_t)(r / NRH - 1) & 1u, 1); }
                else if (bi >= NB) RING_WAIT(hempty + 8 * bs, (uint32_t)(bi / NB - 1) & 1u, 1);
                __syncwarp();
                if (nvalid >= C::NSTEP) sad_walk<HALF, HW, false>(Lr, Rr, Hout, nvalid - HW * wh, act);
                else                    sad_walk<HALF, HW, true>(Lr, Rr, Hout, nvalid - HW * wh, act);
                __syncwarp();
                if (act && (lane & 15) == 0) ring_arrive(tempty + 8 * ts);
            }
            if (lane == 0) ring_arrive(hfull + 8 * bs);
        }
    } else if (warp == C::W_TAIL) {
        // ---- tail walker: group NGC-1 of the rows of a burst, RT rows per pass; lane = (row j, segment s of SEGW columns) ----
        const int j0 = lane / C::NSEG, s = lane % C::NSEG;
        for (int bi = 0; bi < nbur; ++bi) {
            const int bs = bi % NB;
#pragma unroll
            for (int pass = 0; pass < C::TPASS; ++pass) {
                const int j = j0 + C::RT * pass;
                const int r = NWK * bi + j;
                const bool act = (C::TPASS == 1 || j < NWK) && r < nin;
                const int ts = r % TR;
                if (act) {
                    RING_WAIT(tfull + 8 * ts, (uint32_t)(r / TR) & 1u, 0);
                    if (C::ROWREL && r >= NRH) RING_WAIT(hempty + 8 * (r % NRH), (uint32_t)(r / NRH - 1) & 1u, 1);
                }
                if (!C::ROWREL && pass == 0 && bi >= NB) RING_WAIT(hempty + 8 * bs, (uint32_t)(bi / NB - 1) & 1u, 1);
                __syncwarp();
                const uint32_t* Lr = Lrep + ts * C::LWS + C::SEGW * s;
                const uint32_t* Rr = Ral + ts * C::RWS + (C::SEGW / 4) * s;
                uint2* Hout = Hs + (bs * NWK + ((C::TPASS == 1 || j < NWK) ? j : 0)) * HROW + (NGC - 1) * TWP + C::SEGW * s;
                if (nvalid >= C::NSTEP) sad_walk<HALF, C::SEGW, false>(Lr, Rr, Hout, nvalid - C::SEGW * s, act);
                else                    sad_walk<HALF, C::SEGW, true>(Lr, Rr, Hout, nvalid - C::SEGW * s, act);
                __syncwarp();
                if (act && s == 0) ring_arrive(tempty + 8 * ts);
            }
            if (lane == 0) ring_arrive(hfull + 8 * bs);
        }
    } else if (warp == C::W_LD0 || warp == C::W_LD1) {
        // ---- loaders: rows in groups of four, alternating between the two warps; replicated left pixels and aligned
        //      right words of one row per tile slot ----
        const int li = warp == C::W_LD0 ? 0 : 1;
        const uint8_t* __restrict__ Lg = a.L + (long long)frame * a.frameL;
        const uint8_t* __restrict__ Rg = a.R + (long long)frame * a.frameR;
        const int xr0 = x0 - HALF - 3 - 4 * (g0 + NGC - 1) - C::OFF;     // image column of R tile word 0 (multiple of 4)
        constexpr int NLQ = (LW + 31) / 32, NRQ = (RW + 31) / 32;
        int lx[NLQ], rx[NRQ], rmode[NRQ];
#pragma unroll
        for (int q = 0; q < NLQ; ++q) {
            const int i = lane + 32 * q, x = x0 - HALF + i;
            lx[q] = (i < LW && (unsigned)x < (unsigned)a.W) ? x : -1;
        }
#pragma unroll
        for (int q = 0; q < NRQ; ++q) {
            const int jw = lane + 32 * q, x = xr0 + 4 * jw;
            const bool in = jw < RW && x + 3 >= 0 && x < a.W;
            rx[q] = x;
            rmode[q] = !in ? 0 : (a.aligned && x >= 0 && x + 3 < a.W) ? 1 : 2;
        }
        constexpr int GRP = 4;
        // interior strips (every tile column inside the image, 4-byte aligned rows): no per-lane tests in the load loop
        const bool interior = a.aligned && x0 - HALF >= 0 && x0 - HALF + LW <= a.W && xr0 >= 0 && xr0 + 4 * RW <= a.W;
        int lic[NLQ], rjc[NRQ];                                                   // clamped: the surplus lanes reload a valid element
#pragma unroll
        for (int q = 0; q < NLQ; ++q) lic[q] = x0 - HALF + min(lane + 32 * q, LW - 1);
#pragma unroll
        for (int q = 0; q < NRQ; ++q) rjc[q] = xr0 + 4 * min(lane + 32 * q, RW - 1);
        auto issue = [&](int t, uint32_t (&vl)[GRP][NLQ], uint32_t (&vr)[GRP][NRQ]) {      // global loads of the burst starting at row t
#pragma unroll
            for (int u = 0; u < GRP; ++u) {
                const int y = yb0 - HALF + t + u;
                const bool yin = t + u < nin && (unsigned)y < (unsigned)a.H;
                const uint8_t* pl = Lg + (size_t)(yin ? y : 0) * a.pitchL;
                const uint8_t* pr = Rg + (size_t)(yin ? y : 0) * a.pitchR;
                if (interior) {
#pragma unroll
                    for (int q = 0; q < NLQ; ++q) vl[u][q] = yin ? (uint32_t)pl[lic[q]] : 0u;
#pragma unroll
                    for (int q = 0; q < NRQ; ++q) vr[u][q] = yin ? *reinterpret_cast<const uint32_t*>(pr + rjc[q]) : 0u;
                    continue;
                }
#pragma unroll
                for (int q = 0; q < NLQ; ++q) { vl[u][q] = 0; if (yin && lx[q] >= 0) vl[u][q] = pl[lx[q]]; }
#pragma unroll
                for (int q = 0; q < NRQ; ++q) {
                    uint32_t v = 0;
                    if (yin && rmode[q] == 1) v = *reinterpret_cast<const uint32_t*>(pr + rx[q]);
                    else if (yin && rmode[q] == 2) {
#pragma unroll
                        for (int b = 0; b < 4; ++b)
                            if ((unsigned)(rx[q] + b) < (unsigned)a.W) v |= (uint32_t)pr[rx[q] + b] << (8 * b);
                    }
                    vr[u][q] = v;
                }
            }
        };
        auto commit = [&](int t, uint32_t (&vl)[GRP][NLQ], uint32_t (&vr)[GRP][NRQ]) {     // registers -> tile slots, one arrive per row
#pragma unroll
            for (int u = 0; u < GRP; ++u) {
                const int r = t + u;
                if (r >= nin) break;
                const int ts = r % TR;
                if (r >= TR) RING_WAIT(tempty + 8 * ts, (uint32_t)(r / TR - 1) & 1u, 2);
                uint32_t* Ld = Lrep + ts * C::LWS;
                uint32_t* Rd = Ral + ts * C::RWS;
#pragma unroll
                for (int q = 0; q < NLQ; ++q) { const int i = lane + 32 * q; if (i < LW) Ld[i] = vl[u][q] * 0x01010101u; }
#pragma unroll
                for (int q = 0; q < NRQ; ++q) { const int jw = lane + 32 * q; if (jw < RW) Rd[jw] = vr[u][q]; }
                __syncwarp();
                if (lane == 0) ring_arrive(tfull + 8 * ts);
            }
        };
        // two register sets: the loads of this warp's next burst are in flight while the current one waits for its slots
        uint32_t vlA[GRP][NLQ], vrA[GRP][NRQ], vlB[GRP][NLQ], vrB[GRP][NRQ];
        int t = GRP * li;
        if (t < nin) issue(t, vlA, vrA);
        for (; t < nin; t += 4 * GRP) {
            const bool more = t + 2 * GRP < nin;
            if (more) issue(t + 2 * GRP, vlB, vrB);
            commit(t, vlA, vrA);
            if (more) {
                if (t + 4 * GRP < nin) issue(t + 4 * GRP, vlA, vrA);
                commit(t + 2 * GRP, vlB, vrB);
            }
        }
    } else if (warp == C::W_FIN || warp == C::W_FIN2) {
        // ---- finishers: min over the K partial keys of a pixel, LUT, store; each warp takes two rows of every burst ----
        const int fh = warp == C::W_FIN ? 0 : NWK / 2;
        uint8_t* __restrict__ Og = a.out + (long long)frame * a.frameOut;
        const int x = x0 + lane;
        for (int bi = OB0, ob = 0; bi < nbur; ++bi, ++ob) {
            RING_WAIT(pkfull + 8 * (ob & 1), (uint32_t)(ob >> 1) & 1u, 4);
            uint32_t* pkb = pk + (ob & 1) * NWK * C::PKK * TW + lane;
#pragma unroll
            for (int uu = 0; uu < NWK / 2; ++uu) {
                const int u = fh + uu;
                const int k = NWK * bi + u - 2 * HALF, y = yb0 + k;
                if (k < 0 || k >= bhc || x >= a.W) continue;
                uint32_t best = 0xFFFFFFFFu;
                if (C::PKATOM) { best = pkb[u * TW]; pkb[u * TW] = 0xFFFFFFFFu; }     // read and re-arm for the burst after next
                else {
#pragma unroll
                    for (int kk = 0; kk < K; ++kk) best = min(best, pkb[(u * K + kk) * TW]);
                }
                if (x < HALF) best = 0;                      // sad.go:212-218: both windows clamp, d = 0 wins
                if (a.NC == 1) Og[(size_t)y * a.pitchOut + x] = lut[best & (C::WIDE ? 511u : 0xFFFFu)];
                else atomicMin(a.gkey + ((size_t)frame * a.H + y) * a.W + x,
                               C::WIDE ? best : (((best >> 16) << 9) | (best & 511u)));
            }
            __syncwarp();
            if (lane == 0) ring_arrive(pkempty + 8 * (ob & 1));
        }
    } else if (warp >= C::W_CONS && warp < C::W_CONS + K) {
        // ---- consumers: warp k owns groups 3k..3k+2 of 32 columns (lanes = columns); vertical running sums from the
        //      entering and the leaving H row, keys, partial best -> pk; one barrier wait per burst of four rows ----
        const int kB = warp - C::W_CONS;
        const int xB = x0 + lane;
        const int dmax = min(a.D, xB - HALF);               // largest evaluated disparity of this column (sad.go:64-67, :212-218)
        // key constants: (sum << 16 | d) from 16x2-packed sums, or sum * 512 + d from 32-bit sums (h >= 8); a candidate that is
        // never evaluated gets multiplier 0 / mask 0 and an all-ones addend
        const uint32_t kmul = C::WIDE ? (opaque(a.k65536) >> 7) : opaque(a.k65536);
        uint32_t mE[GT], aE[GT], nE[GT], oE[GT], mO[GT], aO[GT], nO[GT], oO[GT], m3E[GT], m3O[GT];
        uint32_t VE[GT], VO[GT], V3[GT], V2[GT];          // narrow: VE/VO packed; wide: VE/VO raw low-lane sums, V3/V2 high-lane sums
#pragma unroll
        for (int j = 0; j < GT; ++j) {
            const int gs = kB * GT + j;
            const int dG = 4 * (g0 + gs);
            const int dm = gs < NGC ? dmax : -1;             // surplus group slot: nothing valid
            const bool v3 = dG + 3 <= dm, v1 = dG + 1 <= dm, v2 = dG + 2 <= dm, v0 = dG <= dm;
            mE[j] = v3 ? kmul : 0u; aE[j] = v3 ? (uint32_t)(dG + 3) : 0xFFFFFFFFu;
            mO[j] = v2 ? kmul : 0u; aO[j] = v2 ? (uint32_t)(dG + 2) : 0xFFFFFFFFu;
            nE[j] = v1 ? (C::WIDE ? kmul : 0xFFFF0000u) : 0u; oE[j] = v1 ? (uint32_t)(dG + 1) : 0xFFFFFFFFu;
            nO[j] = v0 ? (C::WIDE ? kmul : 0xFFFF0000u) : 0u; oO[j] = v0 ? (uint32_t)dG : 0xFFFFFFFFu;
            m3E[j] = v3 ? 0u - (kmul << 16) : 0u;            // wide: -(2^25) removes the high lane from the raw low-lane sum
            m3O[j] = v2 ? 0u - (kmul << 16) : 0u;
            VE[j] = 0; VO[j] = 0; V3[j] = 0; V2[j] = 0;
        }
        const uint2* Hk = Hs + (kB * GT) * TWP + lane;
        auto keys = [&]() {
            uint32_t best = 0xFFFFFFFFu;
#pragma unroll
            for (int j = 0; j < GT; ++j) {
                uint32_t kEl, kEh, kOl, kOh;
                if (C::WIDE) {
                    kEl = VE[j] * mE[j] + (V3[j] * m3E[j] + aE[j]); kEh = V3[j] * nE[j] + oE[j];
                    kOl = VO[j] * mO[j] + (V2[j] * m3O[j] + aO[j]); kOh = V2[j] * nO[j] + oO[j];
                } else {
                    kEl = VE[j] * mE[j] + aE[j]; kEh = (VE[j] & nE[j]) | oE[j];
                    kOl = VO[j] * mO[j] + aO[j]; kOh = (VO[j] & nO[j]) | oO[j];
                }
                best = min(best, min(kEl, kEh));
                best = min(best, min(kOl, kOh));
            }
            return best;
        };
        // V += entering - leaving.  Wide: the packed difference is formed with a bias of 0x8000 in the low lane only, so
        // that the low lane never borrows from the high lane; the arithmetic shift then yields the signed high-lane
        // difference and the raw sum (bias removed in the same IADD3) carries low + 65536 * high.
        auto update = [&](int j, uint2 n, uint2 o) {
            if (C::WIDE) {
                const uint32_t dE = n.x - o.x + 0x8000u, dO = n.y - o.y + 0x8000u;
                VE[j] = VE[j] + dE - 0x8000u; V3[j] += (uint32_t)((int)dE >> 16);
                VO[j] = VO[j] + dO - 0x8000u; V2[j] += (uint32_t)((int)dO >> 16);
            } else {
                VE[j] = VE[j] + n.x - o.x; VO[j] = VO[j] + n.y - o.y;
            }
        };
        for (int bi = 0; bi < nbur; ++bi) {
            const int bs = bi % NB, ob = bi - OB0;
            RING_WAIT(hfull + 8 * bs, (uint32_t)(bi / NB) & 1u, 3);
            if (ob >= 2) RING_WAIT(pkempty + 8 * (ob & 1), (uint32_t)((ob >> 1) - 1) & 1u, 5);
            uint32_t* pkb = pk + ((ob & 1) * NWK * C::PKK + (C::PKATOM ? 0 : kB)) * TW + lane;
            auto emit = [&](int u, uint32_t key) {
                if (C::PKATOM) atomicMin(pkb + u * TW, key);
                else pkb[u * K * TW] = key;
            };
            const int sn0 = bs * NWK;
            if (NWK * bi >= WIN && NWK * bi + NWK - 1 < nin) {
                // steady state: every row of the burst has a leaving row and an output row
#pragma unroll
                for (int u = 0; u < NWK; ++u) {
                    int so = sn0 + u - WIN; if (so < 0) so += NRH;
                    const uint2* Hn = Hk + (sn0 + u) * HROW;
                    const uint2* Ho = Hk + so * HROW;
#pragma unroll
                    for (int j = 0; j < GT; ++j) update(j, Hn[j * TWP], Ho[j * TWP]);
                    if (C::ROWREL) { __syncwarp(); if (lane == 0) ring_arrive(hempty + 8 * so); }     // the leaving row is free at once
                    emit(u, keys());
                }
            } else {
                for (int u = 0; u < NWK; ++u) {
                    const int r = NWK * bi + u;
                    if (r >= nin) break;
                    const uint2* Hn = Hk + (sn0 + u) * HROW;
                    int so = sn0 + u - WIN; if (so < 0) so += NRH;
                    const uint2* Ho = Hk + so * HROW;
#pragma unroll
                    for (int j = 0; j < GT; ++j) update(j, Hn[j * TWP], r >= WIN ? Ho[j * TWP] : make_uint2(0u, 0u));
                    if (C::ROWREL && r >= WIN) { __syncwarp(); if (lane == 0) ring_arrive(hempty + 8 * so); }
                    if (r >= 2 * HALF) emit(u, keys());
                }
            }
            __syncwarp();
            if (lane == 0) {
                if (!C::ROWREL && bi >= OB) ring_arrive(hempty + 8 * ((bi - OB) % NB));      // burst bi-OB has left every window
                if (ob >= 0) ring_arrive(pkfull + 8 * (ob & 1));
            }
        }
    }
    RING_PROF_END(dbg);
}

}  // namespace sadgpu
