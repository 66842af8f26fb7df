// sad_wide.cuh — compile-time specialised kernel for large windows, block_size 16..31 (h = 8..15), where a window
// sum needs 18 bits.  Same arithmetic as the other kernels (pkg/despair/sad.go:55-95, :205-244 via the separable box
// filter); structure = sad_fast.cuh (phase-alternating CTA, fully unrolled walks, immediate-offset shared-memory
// access) with three differences forced by the window size:
//   * the 2h+1 previous rows of H (up to 31) do not fit a register ring: they live in a SHARED-MEMORY ring of
//     RB + 2h+1 rows; phase B reads the entering and the leaving row (two LDS.64 per 4 disparities);
//   * vertical sums are 32-bit: the packed difference of the two rows is formed in one IADD3 with a per-lane bias
//     (H <= 7905 < 2^13, so n + 0x20002000 - o never borrows across the 16-bit lanes) and unpacked;
//     keys are sum*512 + d (IMAD);
//   * a row walk is split into 2 or 4 segments (32 / 16 outputs) so that a 12-row x 8-group batch still has enough
//     items to occupy the CTA's 512 threads (phase B: one thread per (column, group)).
// The disparity range is processed in chunks of 8 groups (32 disparities) merged by atomicMin on the key map.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "sad_fast.cuh"

namespace sadgpu {

template <int HALF> struct WideCfg {
    static_assert(HALF >= 8 && HALF <= 15, "wide kernel: block_size 16..31");
    static constexpr int WIN = 2 * HALF + 1;
    static constexpr int TW = 64, TWP = 65;
    static constexpr int SEG = HALF >= 12 ? 4 : 2, SEGW = TW / SEG;  // walk segments per row (more items per batch for the widest windows)
    static constexpr int NSTEP = TW + 2 * HALF;                     // columns of L a row needs
    static constexpr int SSTEP = SEGW + 2 * HALF;                   // steps of one segment walk
    static constexpr int LW = (NSTEP + 3) & ~3;
    static constexpr int NGC = 8, GT = 1, K = NGC / GT;             // groups per chunk / per phase-B thread / threads per column
    static constexpr int NT = TW * K;                               // 512
    static constexpr int OFF = ((-(HALF + 3)) % 4 + 4) % 4;
    static constexpr int NWALKW = ((NSTEP - 1 + OFF) >> 2) + 2;
    static constexpr int RW = NGC - 1 + NWALKW;
    static constexpr int ROW_BYTES = NGC * TWP * 8;                 // one H row
    static constexpr int RB = HALF <= 11 ? 16 : 12;                 // rows per batch
    static constexpr int NRS = RB + WIN;                            // ring rows
    static constexpr int H_BYTES = ((NRS * ROW_BYTES + 15) / 16) * 16;
    static constexpr int L_BYTES = RB * LW * 4;
    static constexpr int R_BYTES = ((RB * RW * 4 + 15) / 16) * 16;
    static constexpr int PK_BYTES = RB * K * TW * 4;
    static constexpr int OFF_L = H_BYTES;
    static constexpr int OFF_R = OFF_L + L_BYTES;
    static constexpr int OFF_PK = OFF_R + R_BYTES;
    static constexpr int OFF_LUT = OFF_PK + PK_BYTES;
    static constexpr int SMEM = OFF_LUT + 1040;
    static constexpr uint32_t BIAS = 1u << 22;                      // > any window sum (245 055); (BIAS + sum) * 512 < 2^32
};

template <int HALF>
__global__ void __launch_bounds__(WideCfg<HALF>::NT, 1) sad_wide_kernel(const FastArgs a)
{
    using C = WideCfg<HALF>;
    constexpr int WIN = C::WIN, TW = C::TW, TWP = C::TWP, RB = C::RB, GT = C::GT, K = C::K, NGC = C::NGC, NT = C::NT, NRS = C::NRS;
    extern __shared__ __align__(16) unsigned char smem[];
    uint2* Hs = reinterpret_cast<uint2*>(smem);                                  // [NRS][NGC][TWP]
    uint32_t* Lrep = reinterpret_cast<uint32_t*>(smem + C::OFF_L);
    uint32_t* Ral = reinterpret_cast<uint32_t*>(smem + C::OFF_R);
    uint32_t* pk = reinterpret_cast<uint32_t*>(smem + C::OFF_PK);
    uint8_t* lut = smem + C::OFF_LUT;
    constexpr int HROW = NGC * TWP;                                              // uint2 per ring row

    const int tid = threadIdx.x;
    const int frame = blockIdx.z / a.NC, chunk = blockIdx.z - frame * a.NC;
    const uint8_t* __restrict__ Lg = a.L + (long long)frame * a.frameL;
    const uint8_t* __restrict__ Rg = a.R + (long long)frame * a.frameR;
    const int x0 = blockIdx.x * TW;
    const int yb0 = a.y0 + blockIdx.y * a.BH;
    const int yb1 = min(a.y1, yb0 + a.BH);
    const int g0 = chunk * NGC;
    if (yb0 >= yb1) return;
    // a chunk none of whose disparities is a candidate anywhere in this strip (d > X-h for every column, sad.go:64-67 + :212-218)
    // has nothing to contribute: chunk 0 always runs and writes every pixel
    if (g0 > 0 && min(x0 + TW, a.W) - 1 - HALF < 4 * g0) return;

    for (int d = tid; d < 1040; d += NT) lut[d] = d <= a.D ? (uint8_t)((d * 255) / a.D) : 0;

    // ---- phase-B identity and 32-bit running sums; never-evaluated candidates start at BIAS ----
    const int kB = tid / TW, xlB = tid - kB * TW, xB = x0 + xlB;
    uint32_t V[GT][4];                                           // [group][lane]: lanes = d+3, d+1 (E lo, hi), d+2, d+0 (O lo, hi)
#pragma unroll
    for (int j = 0; j < GT; ++j) {
        const int dbase = 4 * (g0 + kB * GT + j);
        const int dmax = min(a.D, xB - HALF);
        V[j][0] = dbase + 3 > dmax ? C::BIAS : 0u;
        V[j][1] = dbase + 1 > dmax ? C::BIAS : 0u;
        V[j][2] = dbase + 2 > dmax ? C::BIAS : 0u;
        V[j][3] = dbase + 0 > dmax ? C::BIAS : 0u;
    }
    const uint32_t keybase = 4u * (uint32_t)(g0 + kB * GT);
    const uint32_t k512 = opaque(a.k65536 >> 7);                 // 512 in a register: keys are IMADs with an immediate d

    const int r0 = yb0 - HALF;
    const int nrows = (yb1 - yb0) + 2 * HALF;
    const int nbatches = (nrows + RB - 1) / RB;
    const int xr0 = x0 - HALF - 3 - 4 * (g0 + NGC - 1) - C::OFF;
    const int nvalid = a.W - (x0 - HALF);

    // Tile loads are split in two: the global loads of the next batch are issued into registers at the start of phase B and
    // written to shared memory at its end, so that their latency (HBM when the frames stream) overlaps phase B.
    constexpr int NLE = (RB * C::LW + NT - 1) / NT, NRE = (RB * C::RW + NT - 1) / NT;
    uint32_t tl[NLE], tr[NRE];
    auto issue_tiles = [&](int rbase) {
#pragma unroll
        for (int q = 0; q < NLE; ++q) {
            const int idx = tid + q * NT;
            const int rb = idx / C::LW, i = idx - rb * C::LW;
            const int y = rbase + rb, x = x0 - HALF + i;
            uint32_t v = 0;
            if (idx < RB * C::LW && (unsigned)y < (unsigned)a.H && (unsigned)x < (unsigned)a.W) v = Lg[(size_t)y * a.pitchL + x];
            tl[q] = v;
        }
#pragma unroll
        for (int q = 0; q < NRE; ++q) {
            const int idx = tid + q * NT;
            const int rb = idx / C::RW, j = idx - rb * C::RW;
            const int y = rbase + rb, x = xr0 + 4 * j;
            uint32_t v = 0;
            if (idx < RB * C::RW && (unsigned)y < (unsigned)a.H && x + 3 >= 0 && x < a.W) {
                const uint8_t* p = Rg + (size_t)y * a.pitchR;
                if (a.aligned && x >= 0 && x + 3 < a.W) v = *reinterpret_cast<const uint32_t*>(p + x);
                else {
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if ((unsigned)(x + b) < (unsigned)a.W) v |= (uint32_t)p[x + b] << (8 * b);
                }
            }
            tr[q] = v;
        }
    };
    auto commit_tiles = [&]() {
#pragma unroll
        for (int q = 0; q < NLE; ++q) { const int idx = tid + q * NT; if (idx < RB * C::LW) Lrep[idx] = tl[q] * 0x01010101u; }
#pragma unroll
        for (int q = 0; q < NRE; ++q) { const int idx = tid + q * NT; if (idx < RB * C::RW) Ral[idx] = tr[q]; }
    };

    auto phaseC = [&](int batch) {
        for (int idx = tid; idx < RB * TW; idx += NT) {
            const int rb = idx / TW, xl = idx - rb * TW;
            const int rel = batch * RB + rb, y = r0 + rel - HALF, x = x0 + xl;
            if (rel < 2 * HALF || y >= yb1 || x >= a.W) continue;
            uint32_t best = 0xFFFFFFFFu;
#pragma unroll
            for (int k = 0; k < K; ++k) best = min(best, pk[(rb * K + k) * TW + xl]);
            if (x < HALF) best = 0;                              // sad.go:212-218: both windows clamp, d = 0 wins
            if (a.NC == 1) (a.out + (long long)frame * a.frameOut)[(size_t)y * a.pitchOut + x] = lut[best & 511u];
            else atomicMin(a.gkey + ((size_t)frame * a.H + y) * a.W + x, best);
        }
    };

    issue_tiles(r0); commit_tiles();
    __syncthreads();
    int slot0 = 0;                                               // ring slot of the first row of the current batch
    for (int batch = 0; batch < nbatches; ++batch) {
        const int rbase = r0 + batch * RB;
        // ---- phase A (and phase C of the previous batch): item = (row, group, segment) ----
        if (batch > 0) phaseC(batch - 1);
        for (int item = tid; item < RB * NGC * C::SEG; item += NT) {
            const int seg = item / (RB * NGC), rem = item - seg * (RB * NGC);
            const int rb = rem / NGC, gl = rem - rb * NGC;
            int slot = slot0 + rb; if (slot >= NRS) slot -= NRS;
            const uint32_t* Lr = Lrep + rb * C::LW + seg * C::SEGW;
            const uint32_t* Rr = Ral + rb * C::RW + (NGC - 1 - gl) + seg * (C::SEGW / 4);
            uint2* Hout = Hs + slot * HROW + gl * TWP + seg * C::SEGW;
            const int nv = nvalid - seg * C::SEGW;
            if (nv >= C::SSTEP) sad_walk<HALF, C::SEGW, false>(Lr, Rr, Hout, nv);
            else                sad_walk<HALF, C::SEGW, true>(Lr, Rr, Hout, nv);
        }
        __syncthreads();
        // ---- phase B (and the tile load of the next batch) ----
        if (batch + 1 < nbatches) issue_tiles(rbase + RB);
        {
            int sn = slot0;                                      // slot of the entering row
            int so = slot0 - WIN; if (so < 0) so += NRS;         // slot of the leaving row (2h+1 rows behind)
            const uint2* Hb = Hs + (kB * GT) * TWP + xlB;
            for (int rb = 0; rb < RB; ++rb) {
                const int rel = batch * RB + rb;
                const bool has_old = rel >= WIN;
                const bool emit = rel >= 2 * HALF && (r0 + rel - HALF) < yb1;
                const uint2* pn = Hb + sn * HROW;
                const uint2* po = Hb + so * HROW;
                uint32_t best = 0xFFFFFFFFu;
#pragma unroll
                for (int j = 0; j < GT; ++j) {
                    const uint2 n = pn[j * TWP];
                    uint2 o = make_uint2(0u, 0u);
                    if (has_old) o = po[j * TWP];
                    const uint32_t tE = n.x + 0x20002000u - o.x;        // per-lane n - o + 8192, no borrow (H < 8192)
                    const uint32_t tO = n.y + 0x20002000u - o.y;
                    V[j][0] = V[j][0] + (tE & 0xFFFFu) - 8192u;
                    V[j][1] = V[j][1] + (tE >> 16) - 8192u;
                    V[j][2] = V[j][2] + (tO & 0xFFFFu) - 8192u;
                    V[j][3] = V[j][3] + (tO >> 16) - 8192u;
                    if (emit) {
                        const uint32_t k0 = V[j][0] * k512 + (4u * j + 3u), k1 = V[j][1] * k512 + (4u * j + 1u);
                        const uint32_t k2 = V[j][2] * k512 + (4u * j + 2u), k3 = V[j][3] * k512 + (4u * j + 0u);
                        best = min(best, min(k0, k1));
                        best = min(best, min(k2, k3));
                    }
                }
                if (emit) pk[(rb * K + kB) * TW + xlB] = best + keybase;
                if (++sn == NRS) sn = 0;
                if (++so == NRS) so = 0;
            }
        }
        slot0 += RB; if (slot0 >= NRS) slot0 -= NRS;
        if (batch + 1 < nbatches) commit_tiles();          // phase A of this batch is behind the barrier above: the tiles are free
        __syncthreads();
    }
    phaseC(nbatches - 1);
}

}  // namespace sadgpu
