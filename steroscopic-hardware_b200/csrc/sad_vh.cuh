// sad_vh.cuh — vertical-first, warp-specialised kernel for the large windows (block_size 11..31, h = 5..15).
//
// Same arithmetic as the other kernels (pkg/despair/sad.go:55-95, :205-244 through the separable box filter), but the
// two passes are swapped so that NO history of partial sums has to be kept on chip, whatever the window size:
//   V stage  thread = (disparity group, 8 columns), marches DOWN the rows: running column sums
//            C_d(x, Y) = sum_{|dy|<=h} AD_d(x, Y+dy).  The row that leaves the window is not remembered: its absolute
//            differences are RECOMPUTED from the pixel tiles (one VABSDIFF4 + two PRMT per 4 candidates), so the only
//            history is 2h+2 rows of pixels.  Column sums are < 2^13, 16x2-packed for every block size;
//   H stage  thread = (row, disparity group), walks ALONG the row: S_d(X, Y) = sum_{|dx|<=h} C_d(X+dx, Y) with the 2h+1
//            previous C values as SSA registers of the fully unrolled walk; four keys (sum, d) per step, minimum over
//            the warp's 32 groups with one REDUX.MIN, lane 0 stores the pixel's best key.
//            h <= 7: 16x2-packed sums, keys sum<<16|d (sad_fast.cuh); h >= 8: 32-bit sums unpacked from one biased
//            packed subtraction, the common per-step bias cancels in the argmin, keys sum*512+d.
// C rows travel through an 8-row ring in shared memory; every hand-over is an mbarrier (row-granular, no
// __syncthreads in the steady state):
//   loader warp --tile_full--> V warps --c_full--> H warps / tail-H warp --tail_full--> H warps (final min, LUT, store)
//          <--tile_empty--            <--c_empty--                      <--tail_empty--
// Lanes are disparity groups; the 33rd group of a chunk (d = 128 at D = 128) is handled by one tail warp per stage.
// Candidates the reference never evaluates (d > D) lose through per-lane key constants (multiplier 0 / all-ones
// addend); d > X-h only occurs in the first strips, which run an EDGE instance of the walk.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "sad_fast.cuh"

#ifndef VH_NLOAD
#define VH_NLOAD 1        // loader warps (a second one measured slower: profiles/README.md)
#endif
#ifndef VH_DEFER
#define VH_DEFER 1        // H warps finish a row one C-ring turn after walking it
#endif

namespace sadgpu {

template <int HALF> struct VhCfg {
    static_assert(HALF >= 1 && HALF <= 15, "vh kernel: block_size 3..31");
    static constexpr int WIN = 2 * HALF + 1;
    static constexpr bool WIDE = WIN * WIN * 255 >= 65536;          // h >= 8: window sums need 18 bits
    static constexpr int NQ = 20, NCOL = 4 * NQ;                    // column quads / columns of C per strip
    static constexpr int OFF = ((-(HALF + 3)) % 4 + 4) % 4;         // first column of the strip that the H walk uses
    static constexpr int TW = (NCOL - 2 * HALF - OFF) & ~7;         // output columns per CTA (a multiple of the 8 tail segments)
    static constexpr int NSTEP = TW + 2 * HALF;                     // H walk length
    static constexpr int NCH = (TW + 31) / 32;
    static constexpr int NGC = 33;                                  // groups per chunk: 32 lanes + tail
    static constexpr int CP = 82;                                   // uint2 per (row, group): 164 words = 4 mod 32 -> 128-bit accesses conflict-free
    static constexpr int CROWS = 8;                                 // C ring rows = H warps
    static constexpr int PF = 8;                                    // rows the loader may run ahead
    static constexpr int NR = ((WIN + 1 + PF + 7) / 8) * 8;         // pixel-tile ring rows (a multiple of 2 loaders x 4 rows: one loader per slot)
    static constexpr int RWS = 56;                                  // words per R tile row (NQ + NGC = 53 used)
    static constexpr int PKW = 73;                                  // words per best-key row
    static constexpr int NV = 10;                                   // V warps (2 quads each)
    static constexpr int NT = 768;
    // warp roles: warps 0..11 = V class (10 V, tail-V, loader), warps 12..23 = H class (8 H, tail-H, 3 idle); SM sub-partition = warp % 4
    static constexpr int W_VT = 3, W_LD = 7, W_HT = 18, W_LD2 = 20;
    static constexpr int C_BYTES = CROWS * NGC * CP * 8;
    static constexpr int OFF_L = C_BYTES;
    static constexpr int OFF_R = OFF_L + NR * NCOL * 4;
    static constexpr int SG = TW / 8;                               // tail-H warp: lanes = 4 rows x 8 segments of SG outputs
    static constexpr int PKR = 2 * CROWS;                           // best-key rows: a row is finished one C-ring turn after its walk
    static constexpr int OFF_PK = OFF_R + NR * RWS * 4;             // [PKR][PKW] best keys of groups 0..31
    static constexpr int OFF_PKT = OFF_PK + PKR * PKW * 4;          // [PKR][PKW] keys of the tail group
    static constexpr int OFF_LUT = ((OFF_PKT + PKR * PKW * 4 + 15) / 16) * 16;
    static constexpr int OFF_BAR = OFF_LUT + 1040;
    static constexpr int NBAR = 2 * NR + 2 * CROWS + 2 * PKR;
    static constexpr int SMEM = OFF_BAR + NBAR * 8;
    static constexpr int REGS_V = 56, REGS_H = 104;                 // setmaxnreg targets of the WIDE instances
    static_assert(OFF + NSTEP <= NCOL && TW >= 32 && TW <= 2 * 32 + 8 && CROWS % 4 == 0, "strip geometry");
    static_assert(OFF_BAR % 8 == 0 && C_BYTES % 16 == 0, "alignment");
};

__device__ __forceinline__ void vh_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
// Bounded wait: a protocol error traps instead of hanging the GPU.  try_wait suspends the warp in hardware; the loop
// around it is kept to three instructions.
#ifdef VH_PROFILE
#define VH_WAIT(bar, parity, slot) do { const long long c0_ = clock64(); vh_wait_(bar, parity); wt[slot] += clock64() - c0_; } while (0)
#else
#define VH_WAIT(bar, parity, slot) vh_wait_(bar, parity)
#endif
#ifdef VH_PROFILE        // developer build: per-warp cycles spent in each barrier family, written for one CTA
#define VH_PROF_BEGIN long long wt[8] = {0, 0, 0, 0, 0, 0, 0, 0}; const long long c00_ = clock64()
#define VH_PROF_END(dbg) do { if (dbg && (threadIdx.x & 31) == 0) { wt[7] = clock64() - c00_; \
    for (int i_ = 0; i_ < 8; ++i_) dbg[(threadIdx.x >> 5) * 8 + i_] = (uint32_t)(wt[i_] >> 4); } } while (0)
#else
#define VH_PROF_BEGIN do {} while (0)
#define VH_PROF_END(dbg) do {} while (0)
#endif
__device__ __forceinline__ void vh_wait_(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .u32 n;\n"
        "mov.u32 n, 0;\n"
        "VH_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "@p bra VH_WAIT_DONE;\n"
        "add.u32 n, n, 1;\n"
        "setp.lt.u32 p, n, 4000000;\n"
        "@p bra VH_WAIT_LOOP;\n"
        "VH_WAIT_DONE:\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok) __trap();
}

// Per-lane key constants: candidates d > D get multiplier 0 / mask 0 and an all-ones addend, i.e. the maximal key.
struct VhKeys {
    uint32_t mE, aE, nE, oE, mO, aO, nO, oO;       // E = (d+3 | d+1 << 16), O = (d+2 | d << 16)
    uint32_t m3E, m3O;                             // wide only: -(2^25) or 0
};

template <bool WIDE>
__device__ __forceinline__ VhKeys vh_make_keys(int dG, int D, uint32_t k65536)
{
    VhKeys k;
    const uint32_t mul = WIDE ? (k65536 >> 7) : k65536;           // 512 or 65536, opaque to the compiler: keys stay IMADs
    const bool v3 = dG + 3 <= D, v1 = dG + 1 <= D, v2 = dG + 2 <= D, v0 = dG <= D;
    k.mE = v3 ? mul : 0u; k.aE = v3 ? (uint32_t)(dG + 3) : 0xFFFFFFFFu;
    k.mO = v2 ? mul : 0u; k.aO = v2 ? (uint32_t)(dG + 2) : 0xFFFFFFFFu;
    if (WIDE) {
        k.nE = v1 ? mul : 0u; k.oE = v1 ? (uint32_t)(dG + 1) : 0xFFFFFFFFu;
        k.nO = v0 ? mul : 0u; k.oO = v0 ? (uint32_t)dG : 0xFFFFFFFFu;
        k.m3E = v3 ? 0u - (k65536 << 9) : 0u;                      // -(2^25): removes the high lane from the raw low-lane sum
        k.m3O = v2 ? 0u - (k65536 << 9) : 0u;
    } else {
        k.nE = v1 ? 0xFFFF0000u : 0u; k.oE = v1 ? (uint32_t)(dG + 1) : 0xFFFFFFFFu;
        k.nO = v0 ? 0xFFFF0000u : 0u; k.oO = v0 ? (uint32_t)dG : 0xFFFFFFFFu;
        k.m3E = 0; k.m3O = 0;
    }
    return k;
}

// ---- H stage: one (row, group) walk over the strip.  Crow = &C[slot][group][0].  REDUCE: minimum over the warp's
//      lanes + lane-0 store (H warps); otherwise every lane stores its own keys (tail warp, lanes = rows). ----
template <int HALF, bool EDGE, bool REDUCE>
__device__ __forceinline__ void vh_hwalk(const uint2* __restrict__ Crow, uint32_t* __restrict__ pkrow, const VhKeys& K,
                                         int t0, bool store)
{
    using T = VhCfg<HALF>;
    constexpr int WIN = T::WIN, NSTEP = T::NSTEP, OFF = T::OFF;
    uint32_t ce[NSTEP], co[NSTEP];
    uint32_t SE = 0, SO = 0, S3 = 0, S2 = 0;           // narrow: SE/SO packed; wide: SE/SO raw low-lane sums, S3/S2 high lanes
    uint4 cur = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int i = 0; i < NSTEP; ++i) {
        const int col = OFF + i;
        if (i == 0 || (col & 1) == 0) cur = *reinterpret_cast<const uint4*>(Crow + (col & ~1));
        ce[i] = (col & 1) ? cur.z : cur.x;
        co[i] = (col & 1) ? cur.w : cur.y;
        uint32_t k3, k1, k2, k0;
        if (!T::WIDE) {
            if (i >= WIN) { SE = SE + ce[i] - ce[i - WIN]; SO = SO + co[i] - co[i - WIN]; }
            else          { SE += ce[i]; SO += co[i]; }
            if (i < 2 * HALF) continue;
            k3 = SE * K.mE + K.aE; k1 = (SE & K.nE) | K.oE;
            k2 = SO * K.mO + K.aO; k0 = (SO & K.nO) | K.oO;
        } else {
            // every lane value of dE/dO is 0x8000 + (entering - leaving) > 0: no borrow between the 16-bit lanes; the
            // bias accumulates equally in all sums of a step and cancels in the argmin (removed when the key is stored)
            uint32_t dE, dO;
            if (i >= WIN) { dE = ce[i] + 0x80008000u - ce[i - WIN]; dO = co[i] + 0x80008000u - co[i - WIN]; }
            else          { dE = ce[i] + 0x80008000u; dO = co[i] + 0x80008000u; }
            SE += dE; S3 += dE >> 16; SO += dO; S2 += dO >> 16;
            if (i < 2 * HALF) continue;
            k3 = SE * K.mE + (S3 * K.m3E + K.aE);                   // (SE - 65536*S3)*512 + d
            k1 = S3 * K.nE + K.oE;
            k2 = SO * K.mO + (S2 * K.m3O + K.aO);
            k0 = S2 * K.nO + K.oO;
        }
        const int j = i - 2 * HALF;
        if (EDGE) {                                        // sad.go:64-67 + :212-218: only d <= X-h are candidates
            k3 = (t0 + j >= 3) ? k3 : 0xFFFFFFFFu; k2 = (t0 + j >= 2) ? k2 : 0xFFFFFFFFu;
            k1 = (t0 + j >= 1) ? k1 : 0xFFFFFFFFu; k0 = (t0 + j >= 0) ? k0 : 0xFFFFFFFFu;
        }
        uint32_t m = min(min(k3, k1), min(k2, k0));
        if (T::WIDE) m -= (uint32_t)(i + 1) << 24;         // the accumulated bias, (i+1)*0x8000*512 (invalid keys stay > 2^31)
        if (REDUCE) {
            m = __reduce_min_sync(0xFFFFFFFFu, m);
            if (store) pkrow[j] = m;
        } else {
            if (store) pkrow[j] = m;
        }
    }
}

// ---- H stage of the tail group: lanes = segments of SG outputs (row-granular, short latency: the C ring row is released
//      quickly); each lane warms its window up over 2h columns.  Crow = &C[slot][NGC-1][0]. ----
template <int HALF, bool EDGE>
__device__ __forceinline__ void vh_tailwalk(const uint2* __restrict__ Crow, uint32_t* __restrict__ pkrow, const VhKeys& K,
                                            int t0, int seg, bool act)
{
    using T = VhCfg<HALF>;
    constexpr int WIN = T::WIN, SG = T::SG, OFF = T::OFF;
    const int j0 = seg * SG;
    const uint2* Cp = Crow + OFF + j0;
    uint32_t ce[SG + 2 * HALF], co[SG + 2 * HALF];
    uint32_t SE = 0, SO = 0, S3 = 0, S2 = 0;
#pragma unroll
    for (int i = 0; i < SG + 2 * HALF; ++i) {
        const uint2 c = Cp[i];
        ce[i] = c.x; co[i] = c.y;
        uint32_t k3, k1, k2, k0;
        if (!T::WIDE) {
            if (i >= WIN) { SE = SE + ce[i] - ce[i - WIN]; SO = SO + co[i] - co[i - WIN]; }
            else          { SE += ce[i]; SO += co[i]; }
            if (i < 2 * HALF) continue;
            k3 = SE * K.mE + K.aE; k1 = (SE & K.nE) | K.oE;
            k2 = SO * K.mO + K.aO; k0 = (SO & K.nO) | K.oO;
        } else {
            uint32_t dE, dO;
            if (i >= WIN) { dE = ce[i] + 0x80008000u - ce[i - WIN]; dO = co[i] + 0x80008000u - co[i - WIN]; }
            else          { dE = ce[i] + 0x80008000u; dO = co[i] + 0x80008000u; }
            SE += dE; S3 += dE >> 16; SO += dO; S2 += dO >> 16;
            if (i < 2 * HALF) continue;
            k3 = SE * K.mE + (S3 * K.m3E + K.aE);
            k1 = S3 * K.nE + K.oE;
            k2 = SO * K.mO + (S2 * K.m3O + K.aO);
            k0 = S2 * K.nO + K.oO;
        }
        const int jj = i - 2 * HALF;                          // output j0 + jj
        if (EDGE) {
            const int t = t0 + j0 + jj;
            k3 = (t >= 3) ? k3 : 0xFFFFFFFFu; k2 = (t >= 2) ? k2 : 0xFFFFFFFFu;
            k1 = (t >= 1) ? k1 : 0xFFFFFFFFu; k0 = (t >= 0) ? k0 : 0xFFFFFFFFu;
        }
        uint32_t m = min(min(k3, k1), min(k2, k0));
        if (T::WIDE) m -= (uint32_t)(i + 1) << 24;
        if (act) pkrow[j0 + jj] = m;
    }
}

// ---- V stage: NCT columns x one group, marching down the rows of the band ----
template <int HALF, int NCT, bool EDGE>
__device__ __forceinline__ void vh_vmarch(unsigned char* smem, int g, int cb, bool active, int lane, int nin, int nvalid, uint32_t* dbg)
{
    VH_PROF_BEGIN;
    using T = VhCfg<HALF>;
    constexpr int WIN = T::WIN, NR = T::NR, NCOL = T::NCOL, RWS = T::RWS, NGC = T::NGC, CP = T::CP, CROWS = T::CROWS;
    constexpr int NW = NCT / 4 + 1;
    const uint32_t* Lrep = reinterpret_cast<const uint32_t*>(smem + T::OFF_L);
    const uint32_t* Ral = reinterpret_cast<const uint32_t*>(smem + T::OFF_R);
    uint2* Cs = reinterpret_cast<uint2*>(smem);
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(smem + T::OFF_BAR);
    const uint32_t tfull = bar0, tempty = bar0 + 8 * NR, cfull = bar0 + 16 * NR, cempty = cfull + 8 * CROWS;
    const int widx0 = (cb >> 2) + (NGC - 1 - g);
    const int nv = nvalid - cb;
    uint32_t CE[NCT], CO[NCT];
#pragma unroll
    for (int c = 0; c < NCT; ++c) { CE[c] = 0; CO[c] = 0; }

    auto row_ad = [&](int slot, uint32_t (&ad)[NCT]) {
        const uint32_t* Lr = Lrep + slot * NCOL + cb;
        const uint32_t* Rr = Ral + slot * RWS + widx0;
        uint32_t lw[NCT], rw[NW];
#pragma unroll
        for (int q = 0; q < NCT / 4; ++q) {
            const uint4 v = *reinterpret_cast<const uint4*>(Lr + 4 * q);
            lw[4 * q] = v.x; lw[4 * q + 1] = v.y; lw[4 * q + 2] = v.z; lw[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int q = 0; q < NW; ++q) rw[q] = Rr[q];
#pragma unroll
        for (int c = 0; c < NCT; ++c) {
            const uint32_t r = (c & 3) == 0 ? rw[c >> 2] : __funnelshift_r(rw[c >> 2], rw[(c >> 2) + 1], 8 * (c & 3));
            ad[c] = __vabsdiffu4(lw[c], r);
            if (EDGE) ad[c] = (c < nv) ? ad[c] : 0u;        // columns x >= W contribute nothing
        }
    };

    int sn = 0, so = 0, cs = 0;
    uint32_t phn = 0, phc = 0;                               // parity of the tile_full wait / the c_empty wait
    for (int t = 0; t < nin; ++t) {
        VH_WAIT(tfull + 8 * sn, phn, 0);
        uint32_t adn[NCT];
        row_ad(sn, adn);
        if (t >= WIN) {
            uint32_t ado[NCT];
            row_ad(so, ado);
#pragma unroll
            for (int c = 0; c < NCT; ++c) {
                CE[c] = CE[c] + __byte_perm(adn[c], 0u, 0x4240) - __byte_perm(ado[c], 0u, 0x4240);
                CO[c] = CO[c] + __byte_perm(adn[c], 0u, 0x4341) - __byte_perm(ado[c], 0u, 0x4341);
            }
            __syncwarp();
            if (lane == 0) vh_arrive(tempty + 8 * so);       // the leaving row is not needed any more
            if (++so == NR) so = 0;
        } else {
#pragma unroll
            for (int c = 0; c < NCT; ++c) {
                CE[c] += __byte_perm(adn[c], 0u, 0x4240);
                CO[c] += __byte_perm(adn[c], 0u, 0x4341);
            }
        }
        if (++sn == NR) { sn = 0; phn ^= 1u; }
        if (t >= 2 * HALF) {                                 // C row Y = yb0 + t - 2h is complete
            if (t >= 2 * HALF + CROWS) VH_WAIT(cempty + 8 * cs, phc, 1);
            if (active) {
                uint4* dst = reinterpret_cast<uint4*>(Cs + (cs * NGC + g) * CP + cb);
#pragma unroll
                for (int c = 0; c < NCT; c += 2) dst[c >> 1] = make_uint4(CE[c], CO[c], CE[c + 1], CO[c + 1]);
            }
            __syncwarp();
            if (lane == 0) vh_arrive(cfull + 8 * cs);
            if (++cs == CROWS) { cs = 0; if (t >= 2 * HALF + CROWS) phc ^= 1u; }
        }
    }
    VH_PROF_END(dbg);
}

// ---- loaders: two warps take turns on groups of four rows; one row of replicated left pixels and aligned right words per
//      tile slot; the global loads of a warp's next group are in flight while the current one waits for its slots ----
template <int HALF>
__device__ __forceinline__ void vh_loader(const FastArgs& a, unsigned char* smem, int li, int lane, int frame, int g0, int xq0,
                                          int yb0, int nin, uint32_t* dbg)
{
    using T = VhCfg<HALF>;
    constexpr int NR = T::NR, NCOL = T::NCOL, RWS = T::RWS, NGC = T::NGC;
    uint32_t* Lrep = reinterpret_cast<uint32_t*>(smem + T::OFF_L);
    uint32_t* Ral = reinterpret_cast<uint32_t*>(smem + T::OFF_R);
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(smem + T::OFF_BAR);
    const uint32_t tfull = bar0, tempty = bar0 + 8 * NR;
    const uint8_t* __restrict__ Lg = a.L + (long long)frame * a.frameL;
    const uint8_t* __restrict__ Rg = a.R + (long long)frame * a.frameR;
    const int xr0 = xq0 - 3 - 4 * (g0 + NGC - 1);   // image column of R tile word 0 (multiple of 4)
    constexpr int NLQ = (NCOL + 31) / 32, NRQ = (RWS + 31) / 32;
    int lx[NLQ], rx[NRQ], rmode[NRQ];
#pragma unroll
    for (int q = 0; q < NLQ; ++q) {
        const int c = lane + 32 * q, x = xq0 + c;
        lx[q] = (c < NCOL && (unsigned)x < (unsigned)a.W) ? x : -1;
    }
#pragma unroll
    for (int q = 0; q < NRQ; ++q) {
        const int j = lane + 32 * q, x = xr0 + 4 * j;
        const bool in = j < RWS && x + 3 >= 0 && x < a.W;
        rx[q] = x;
        rmode[q] = !in ? 0 : (a.aligned && x >= 0 && x + 3 < a.W) ? 1 : 2;
    }
    VH_PROF_BEGIN;
    constexpr int GRP = 4;
    auto issue = [&](int t, uint32_t (&vl)[GRP][NLQ], uint32_t (&vr)[GRP][NRQ]) {
#pragma unroll
        for (int u = 0; u < GRP; ++u) {
            const int y = yb0 - HALF + t + u;
            const bool yin = t + u < nin && (unsigned)y < (unsigned)a.H;
            const uint8_t* pl = Lg + (size_t)(yin ? y : 0) * a.pitchL;
            const uint8_t* pr = Rg + (size_t)(yin ? y : 0) * a.pitchR;
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
    auto commit = [&](int t, uint32_t (&vl)[GRP][NLQ], uint32_t (&vr)[GRP][NRQ]) {
#pragma unroll
        for (int u = 0; u < GRP; ++u) {
            const int r = t + u;
            if (r >= nin) break;
            const int slot = r % NR;
            if (r >= NR) VH_WAIT(tempty + 8 * slot, (uint32_t)(r / NR - 1) & 1u, 2);
            uint32_t* Ld = Lrep + slot * NCOL;
            uint32_t* Rd = Ral + slot * RWS;
#pragma unroll
            for (int q = 0; q < NLQ; ++q) { const int c = lane + 32 * q; if (c < NCOL) Ld[c] = vl[u][q] * 0x01010101u; }
#pragma unroll
            for (int q = 0; q < NRQ; ++q) { const int j = lane + 32 * q; if (j < RWS) Rd[j] = vr[u][q]; }
            __syncwarp();
            if (lane == 0) vh_arrive(tfull + 8 * slot);
        }
    };
    uint32_t vlA[GRP][NLQ], vrA[GRP][NRQ], vlB[GRP][NLQ], vrB[GRP][NRQ];
    constexpr int ST = VH_NLOAD * GRP;               // rows between two groups of the same loader
    int t = GRP * li;
    if (t < nin) issue(t, vlA, vrA);
    for (; t < nin; t += 2 * ST) {
        const bool more = t + ST < nin;
        if (more) issue(t + ST, vlB, vrB);
        commit(t, vlA, vrA);
        if (more) {
            if (t + 2 * ST < nin) issue(t + 2 * ST, vlA, vrA);
            commit(t + ST, vlB, vrB);
        }
    }
    VH_PROF_END(dbg);
}

template <int HALF>
__global__ void __launch_bounds__(VhCfg<HALF>::NT, 1) sad_vh_kernel(const __grid_constant__ FastArgs a)
{
    using T = VhCfg<HALF>;
    constexpr int WIN = T::WIN, NR = T::NR, NCOL = T::NCOL, RWS = T::RWS, NGC = T::NGC, CP = T::CP, CROWS = T::CROWS;
    constexpr int TW = T::TW, PKW = T::PKW, NQ = T::NQ;
    extern __shared__ __align__(128) unsigned char smem[];
    uint32_t* Lrep = reinterpret_cast<uint32_t*>(smem + T::OFF_L);
    uint32_t* Ral = reinterpret_cast<uint32_t*>(smem + T::OFF_R);
    uint32_t* pk = reinterpret_cast<uint32_t*>(smem + T::OFF_PK);
    uint32_t* pkT = reinterpret_cast<uint32_t*>(smem + T::OFF_PKT);
    uint8_t* lut = smem + T::OFF_LUT;
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(smem + T::OFF_BAR);
    const uint32_t tfull = bar0, tempty = bar0 + 8 * NR, cfull = bar0 + 16 * NR, cempty = cfull + 8 * CROWS;
    const uint32_t tkfull = cempty + 8 * CROWS, tkempty = tkfull + 8 * T::PKR;

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
    const int bhc = yb1 - yb0, nin = bhc + 2 * HALF;
    const int xq0 = x0 - HALF - T::OFF;                      // image column of C column 0 (= 3 mod 4)
    const int nvalid = a.W - xq0;                            // C columns >= nvalid lie right of the image
    const bool edge_h = x0 - HALF < min(a.D, 4 * (g0 + NGC) - 1);   // some candidate of this chunk exceeds X-h in this strip
    uint32_t* dbg = ((a.debug_skip & 4) && blockIdx.x == 7 && blockIdx.y == 0 && blockIdx.z == 0) ? a.gkey : nullptr;
    (void)dbg;

    for (int d = tid; d < 1040; d += T::NT) lut[d] = d <= a.D ? (uint8_t)((d * 255) / a.D) : 0;
    if (tid == 0) {
        for (int i = 0; i < NR; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(tfull + 8 * i));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(tempty + 8 * i), "n"(T::NV + 1));
        }
        for (int i = 0; i < CROWS; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(cfull + 8 * i), "n"(T::NV + 1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 2;" :: "r"(cempty + 8 * i));
        }
        for (int i = 0; i < T::PKR; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(tkfull + 8 * i));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(tkempty + 8 * i));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp < 12) {
        // ============================ V class ============================
        if (T::WIDE) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(T::REGS_V));
        if (warp == T::W_LD) {
            vh_loader<HALF>(a, smem, 0, lane, frame, g0, xq0, yb0, nin, dbg);
        } else if (warp == T::W_VT) {
            const int q = lane < NQ ? lane : 0;
            if (nvalid >= NCOL) vh_vmarch<HALF, 4, false>(smem, NGC - 1, 4 * q, lane < NQ, lane, nin, nvalid, dbg);
            else                vh_vmarch<HALF, 4, true>(smem, NGC - 1, 4 * q, lane < NQ, lane, nin, nvalid, dbg);
        } else {
            const int vw = warp < T::W_VT ? warp : warp < T::W_LD ? warp - 1 : warp - 2;     // 0..9
            if (nvalid >= NCOL) vh_vmarch<HALF, 8, false>(smem, lane, 8 * vw, true, lane, nin, nvalid, dbg);
            else                vh_vmarch<HALF, 8, true>(smem, lane, 8 * vw, true, lane, nin, nvalid, dbg);
        }
    } else {
        // ============================ H class ============================
        if (T::WIDE) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(T::REGS_H));
        // H warps 12,16,13,17,14,15,19,23 own C ring rows 0..7; warp 18 is the tail-H warp; warps 20,21,22 idle
        int hr = -1;
        switch (warp) { case 12: hr = 0; break; case 16: hr = 1; break; case 13: hr = 2; break; case 17: hr = 3; break;
                        case 14: hr = 4; break; case 15: hr = 5; break; case 19: hr = 6; break; case 23: hr = 7; break; default: break; }
        const uint2* Cs = reinterpret_cast<const uint2*>(smem);
        uint8_t* __restrict__ Og = a.out + (long long)frame * a.frameOut;
        if (hr >= 0) {
            const int dG = 4 * (g0 + lane);
            const VhKeys K = vh_make_keys<T::WIDE>(dG, a.D, opaque(a.k65536));
            const int t0 = x0 - HALF - dG;
            const uint2* Crow = Cs + (hr * NGC + lane) * CP;
            VH_PROF_BEGIN;
            auto finish = [&](int k) {                       // final min with the tail group, LUT, store of C row k
                const int ps = k & (T::PKR - 1);
                VH_WAIT(tkfull + 8 * ps, (uint32_t)(k / T::PKR) & 1u, 4);
                const uint32_t* p0 = pk + ps * PKW;
                const uint32_t* p1 = pkT + ps * PKW;
                const int y = yb0 + k;
#pragma unroll
                for (int c = 0; c < T::NCH; ++c) {
                    const int j = lane + 32 * c, x = x0 + j;
                    if (j < TW && x < a.W) {
                        uint32_t best = min(p0[j], p1[j]);
                        if (x < HALF) best = 0;              // sad.go:212-218: both windows clamp, d = 0 wins
                        if (a.NC == 1) Og[(size_t)y * a.pitchOut + x] = lut[best & (T::WIDE ? 511u : 0xFFFFu)];
                        else atomicMin(a.gkey + ((size_t)frame * a.H + y) * a.W + x,
                                       T::WIDE ? best : (((best >> 16) << 9) | (best & 511u)));
                    }
                }
                __syncwarp();
                if (lane == 0) vh_arrive(tkempty + 8 * ps);
            };
            uint32_t ph = 0;
            int kprev = -1;
            for (int k = hr; k < bhc; k += CROWS) {
                VH_WAIT(cfull + 8 * hr, ph, 3);
                uint32_t* pkrow = pk + (k & (T::PKR - 1)) * PKW;
                if (edge_h) vh_hwalk<HALF, true, true>(Crow, pkrow, K, t0, lane == 0);
                else        vh_hwalk<HALF, false, true>(Crow, pkrow, K, t0, lane == 0);
                __syncwarp();
                if (lane == 0) vh_arrive(cempty + 8 * hr);
#if VH_DEFER
                if (kprev >= 0) finish(kprev);               // one ring turn late: the tail keys of that row have long arrived
                kprev = k;
#else
                finish(k);
#endif
                ph ^= 1u;
            }
            if (kprev >= 0) finish(kprev);
            VH_PROF_END(dbg);
        } else if (warp == T::W_LD2) {
            if (VH_NLOAD == 2) vh_loader<HALF>(a, smem, 1, lane, frame, g0, xq0, yb0, nin, dbg);
        } else if (warp == T::W_HT) {
            // ---- tail-H warp: group NGC-1 of four C rows at a time; lane = (row j, segment s of SG outputs).  Each lane
            //      re-warms its window over 2h columns, so the segments are long and the rows few. ----
            const int dG = 4 * (g0 + NGC - 1);
            const VhKeys K = vh_make_keys<T::WIDE>(dG, a.D, opaque(a.k65536));
            const int t0 = x0 - HALF - dG;
            VH_PROF_BEGIN;
            const int j = lane >> 3, sg = lane & 7;
            for (int k = 0; k < bhc; k += 4) {
                const int cs = (k & (CROWS - 1)) + j;
                const uint32_t ph = (uint32_t)(k / CROWS) & 1u;
                const bool act = k + j < bhc;
                const int ps = (k + j) & (T::PKR - 1);
                if (act) {
                    VH_WAIT(cfull + 8 * cs, ph, 3);
                    if (k + j >= T::PKR) VH_WAIT(tkempty + 8 * ps, (uint32_t)((k + j) / T::PKR - 1) & 1u, 5);
                }
                __syncwarp();
                const uint2* Crow = Cs + (cs * NGC + NGC - 1) * CP;
                if (edge_h) vh_tailwalk<HALF, true>(Crow, pkT + ps * PKW, K, t0, sg, act);
                else        vh_tailwalk<HALF, false>(Crow, pkT + ps * PKW, K, t0, sg, act);
                __syncwarp();
                if (act && sg == 0) { vh_arrive(cempty + 8 * cs); vh_arrive(tkfull + 8 * ps); }
            }
            VH_PROF_END(dbg);
        }
    }
}

}  // namespace sadgpu
