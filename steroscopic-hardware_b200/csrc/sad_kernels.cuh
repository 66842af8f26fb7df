// sad_kernels.cuh — sm_100a kernels for the SAD block-matching disparity path.
//
// Computes, bit-exactly, what pkg/despair/sad.go:55-95 (per-pixel disparity scan) and
// :205-244 (SumAbsoluteDifferences) compute, through the equivalent zero-padded separable
// box filter:
//     AD_d(x,y) = |L(x,y) - R(x-d,y)|   (0 outside the image)
//     S_d(X,Y)  = sum over the (2h+1)^2 window of AD_d,  h = block_size/2
//     out(X,Y)  = (argmin_{d in [0, min(D, X-h)]} S_d, lowest d on ties) * 255 / D ;  0 for X < h
//
// Design: integer cost volume, no tensor cores.  Four disparities are packed
// in the four bytes of a word ("group" g holds d = 4g+3-byte), so one VABSDIFF4 evaluates
// four candidates of one pixel against a replicated left pixel.  A CTA owns a column strip x
// a row band x a chunk of groups and iterates over batches of RB rows:
//   phase A  thread = (row, group): walks x with a running horizontal window sum held in
//            registers (16x2-packed, plain IADD3 — no carry can cross the 16-bit lanes), the
//            2h+1 old terms come from a register ring; result H goes to shared memory.
//   phase B  thread = (column, GT groups): vertical running sum of H over 2h+1 rows
//            (new row added, row 2h+1 behind subtracted), 32-bit (sum<<16 | d) keys and a
//            VIMNMX3 running minimum reproduce the strict-< / ascending-d tie-break.
//   phase C  min over the K threads of a pixel, d*255/D through a LUT, store (or atomicMin
//            into a global key map when the disparity range is split over several CTAs).
// Candidates the reference never evaluates (d > X-h, and the padding slots d > D) are
// replaced by a poison H value 255*(2h+1)+1 per row, which makes their window sum strictly
// larger than any evaluated candidate of the same pixel.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sadgpu {

struct SadArgs {
    const uint8_t* L; const uint8_t* R; uint8_t* out; uint32_t* gkey;
    int pitchL, pitchR, pitchOut;
    int W, H, y0, y1;
    int D, NG, NC, NGc, NGP, K;
    int NSTEP, TW, TWp, RB, NR, BH, RW;
    int offL, offR, offPk, offLut;      // shared-memory byte offsets (H ring at 0)
};

template <int I> __device__ __forceinline__ uint32_t get4(const uint4& v) {
    if constexpr (I == 0) return v.x; else if constexpr (I == 1) return v.y;
    else if constexpr (I == 2) return v.z; else return v.w;
}

// ---------------------------------------------------------------------------------------
// Phase A: one (row, group) item.  Lr: replicated left pixels of the row (word i = pixel
// x0-h+i in all four bytes).  Rr: aligned right words such that the bytes needed at step i
// start at byte i of Rr[0].  Hout[xl*hstride] receives (E,O) = 16x2 packed window sums:
// E = (d=4g+3 | d=4g+1 << 16), O = (d=4g+2 | d=4g << 16).
// ---------------------------------------------------------------------------------------
template <int HALF, bool EDGE>
__device__ __forceinline__ void phaseA_walk(const uint32_t* __restrict__ Lr, const uint32_t* __restrict__ Rr,
                                            uint2* __restrict__ Hout, int hstride, int nblk, int nvalid)
{
    constexpr int WIN = 2 * HALF + 1;
    constexpr int U = HALF <= 7 ? 16 : 32;
    uint32_t rE[U], rO[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { rE[u] = 0; rO[u] = 0; }
    uint32_t hE = 0, hO = 0;
    uint32_t wprev = Rr[0];
    for (int blk = 0; blk < nblk; ++blk) {
        uint32_t w[U / 4 + 1];
        w[0] = wprev;
#pragma unroll
        for (int q = 1; q <= U / 4; ++q) w[q] = Rr[blk * (U / 4) + q];
        uint4 lv[U / 4];
#pragma unroll
        for (int q = 0; q < U / 4; ++q) lv[q] = *reinterpret_cast<const uint4*>(Lr + blk * U + 4 * q);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = blk * U + u;
            uint32_t lw;
            if ((u & 3) == 0) lw = lv[u / 4].x; else if ((u & 3) == 1) lw = lv[u / 4].y;
            else if ((u & 3) == 2) lw = lv[u / 4].z; else lw = lv[u / 4].w;
            const uint32_t rw = (u & 3) == 0 ? w[u / 4] : __funnelshift_r(w[u / 4], w[u / 4 + 1], 8 * (u & 3));
            uint32_t ad = __vabsdiffu4(lw, rw);
            if (EDGE) ad = (i < nvalid) ? ad : 0u;       // columns x' >= W contribute nothing (sad.go:231-233)
            const uint32_t e = __byte_perm(ad, 0u, 0x4240);
            const uint32_t o = __byte_perm(ad, 0u, 0x4341);
            hE = hE + e - rE[(u + U - WIN) % U];
            hO = hO + o - rO[(u + U - WIN) % U];
            rE[u] = e; rO[u] = o;
            if (blk > 0 || u >= 2 * HALF) Hout[(i - 2 * HALF) * hstride] = make_uint2(hE, hO);
        }
        wprev = w[U / 4];
    }
}

template <int HALF, int GT>
__global__ void __launch_bounds__(256, 1) sad_generic_kernel(const SadArgs a)
{
    constexpr bool WIDE = HALF >= 8;                  // (2h+1)^2*255 >= 65536: 32-bit window sums
    constexpr int WIN = 2 * HALF + 1;
    constexpr int U = HALF <= 7 ? 16 : 32;
    constexpr uint32_t P = 255u * WIN + 1u;           // poison H value
    constexpr uint32_t PP = P | (P << 16);
    extern __shared__ __align__(16) unsigned char smem[];
    uint2* Hring = reinterpret_cast<uint2*>(smem);
    uint32_t* Lrep = reinterpret_cast<uint32_t*>(smem + a.offL);
    uint32_t* Ral = reinterpret_cast<uint32_t*>(smem + a.offR);
    uint32_t* pk = reinterpret_cast<uint32_t*>(smem + a.offPk);
    uint8_t* lut = smem + a.offLut;

    const int tid = threadIdx.x, nt = blockDim.x;
    const int x0 = blockIdx.x * a.TW;
    const int yb0 = a.y0 + blockIdx.y * a.BH;
    const int yb1 = min(a.y1, yb0 + a.BH);
    const int g0 = blockIdx.z * a.NGc;
    const int ngc = min(a.NGc, a.NG - g0);
    const int Kc = (ngc + GT - 1) / GT;
    if (yb0 >= yb1 || ngc <= 0) return;

    for (int d = tid; d < 4 * a.NG; d += nt) lut[d] = d <= a.D ? (uint8_t)((d * 255) / a.D) : 0;

    // phase-B identity and state
    const int kB = tid / a.TWp, xlB = tid - kB * a.TWp;
    const bool activeB = kB < Kc && xlB < a.TW;
    const int ngB = activeB ? min(GT, ngc - kB * GT) : 0;
    uint32_t V[GT][WIDE ? 4 : 2];
#pragma unroll
    for (int j = 0; j < GT; ++j)
#pragma unroll
        for (int q = 0; q < (WIDE ? 4 : 2); ++q) V[j][q] = 0;

    const int r0 = yb0 - HALF;
    const int nrows = (yb1 - yb0) + 2 * HALF;
    const int nbatches = (nrows + a.RB - 1) / a.RB;
    const int xr0 = x0 - HALF - 3 - 4 * (g0 + a.NGc - 1);
    const int nvalid = a.W - (x0 - HALF);            // steps with x' < W
    const int nblk = a.NSTEP / U;
    const int RWB = a.RW * 4;

    auto phaseC = [&](int batch) {
        const int rbase = r0 + batch * a.RB;
        for (int idx = tid; idx < a.RB * a.TW; idx += nt) {
            const int rb = idx / a.TW, xl = idx - rb * a.TW;
            const int r = rbase + rb, y = r - HALF, x = x0 + xl;
            if (r - r0 < 2 * HALF || y >= yb1 || x >= a.W) continue;
            uint32_t best = 0xFFFFFFFFu;
            for (int k = 0; k < Kc; ++k) best = min(best, pk[(rb * a.K + k) * a.TW + xl]);
            if (a.NC == 1) {
                const uint32_t d = WIDE ? (best & 511u) : (best & 0xFFFFu);
                a.out[(size_t)y * a.pitchOut + x] = lut[d];
            } else {
                const uint32_t ukey = WIDE ? best : (((best >> 16) << 9) | (best & 511u));
                atomicMin(a.gkey + (size_t)y * a.W + x, ukey);
            }
        }
    };

    for (int batch = 0; batch < nbatches; ++batch) {
        const int rbase = r0 + batch * a.RB;
        // ---- tile load: replicated L pixels, aligned R bytes (zero outside the image) ----
        for (int idx = tid; idx < a.RB * a.NSTEP; idx += nt) {
            const int rb = idx / a.NSTEP, i = idx - rb * a.NSTEP;
            const int y = rbase + rb, x = x0 - HALF + i;
            uint32_t v = 0;
            if ((unsigned)y < (unsigned)a.H && (unsigned)x < (unsigned)a.W) v = a.L[(size_t)y * a.pitchL + x];
            Lrep[idx] = v * 0x01010101u;
        }
        for (int idx = tid; idx < a.RB * RWB; idx += nt) {
            const int rb = idx / RWB, bi = idx - rb * RWB;
            const int y = rbase + rb, x = xr0 + bi;
            uint8_t v = 0;
            if ((unsigned)y < (unsigned)a.H && (unsigned)x < (unsigned)a.W) v = a.R[(size_t)y * a.pitchR + x];
            reinterpret_cast<uint8_t*>(Ral)[idx] = v;
        }
        if (batch > 0) phaseC(batch - 1);
        __syncthreads();

        // ---- phase A (+ poison fix-up by the same thread) ----
        for (int item = tid; item < a.RB * ngc; item += nt) {
            const int rb = item / ngc, gl = item - rb * ngc;
            const int G = g0 + gl;
            const int rel = batch * a.RB + rb;
            const int y = rbase + rb;
            uint2* Hout = Hring + ((size_t)(rel % a.NR) * a.TW) * a.NGP + gl;
            const bool row_in = (unsigned)y < (unsigned)a.H;
            const bool all_invalid = (x0 + a.TW - 1 - HALF) < 4 * G;     // every lane, every column
            if (!row_in) {
                for (int xl = 0; xl < a.TW; ++xl) Hout[xl * a.NGP] = make_uint2(0u, 0u);
                continue;
            }
            if (!all_invalid) {
                const uint32_t* Lr = Lrep + rb * a.NSTEP;
                const uint32_t* Rr = Ral + rb * a.RW + (a.NGc - 1 - gl);
                if (nvalid >= a.NSTEP) phaseA_walk<HALF, false>(Lr, Rr, Hout, a.NGP, nblk, nvalid);
                else                   phaseA_walk<HALF, true>(Lr, Rr, Hout, a.NGP, nblk, nvalid);
            }
            // poison: lanes with d > D (padding) or d > x - h (never evaluated, sad.go:64-67 + :212-218)
            const bool pad = 4 * G + 3 > a.D;
            const int xlim = pad ? a.TW : min(a.TW, 4 * G + 4 + HALF - x0);
            for (int xl = 0; xl < xlim; ++xl) {
                const int t = x0 + xl - HALF;                 // largest evaluated d for this column
                const int dmax = min(t, a.D);
                const uint32_t mE = (4 * G + 3 > dmax ? 0x0000FFFFu : 0u) | (4 * G + 1 > dmax ? 0xFFFF0000u : 0u);
                const uint32_t mO = (4 * G + 2 > dmax ? 0x0000FFFFu : 0u) | (4 * G + 0 > dmax ? 0xFFFF0000u : 0u);
                uint2 v = all_invalid ? make_uint2(0u, 0u) : Hout[xl * a.NGP];
                v.x = (v.x & ~mE) | (PP & mE);
                v.y = (v.y & ~mO) | (PP & mO);
                Hout[xl * a.NGP] = v;
            }
        }
        __syncthreads();

        // ---- phase B: vertical running sums + running argmin ----
        if (activeB) {
            for (int rb = 0; rb < a.RB; ++rb) {
                const int rel = batch * a.RB + rb;
                const bool has_old = rel >= WIN;
                const bool emit = rel >= 2 * HALF && (r0 + rel - HALF) < yb1;
                const uint2* pn = Hring + ((size_t)(rel % a.NR) * a.TW + xlB) * a.NGP + kB * GT;
                const uint2* po = Hring + ((size_t)((has_old ? rel - WIN : 0) % a.NR) * a.TW + xlB) * a.NGP + kB * GT;
                uint32_t best = 0xFFFFFFFFu;
#pragma unroll
                for (int j = 0; j < GT; ++j) {
                    if (j < ngB) {
                        const uint2 n = pn[j];
                        uint2 o = make_uint2(0u, 0u);
                        if (has_old) o = po[j];
                        if constexpr (!WIDE) {
                            V[j][0] = V[j][0] + n.x - o.x;
                            V[j][1] = V[j][1] + n.y - o.y;
                            if (emit) {
                                constexpr uint32_t cE = (4u * 0 + 1u) | ((4u * 0 + 3u) << 16);
                                const uint32_t ce = cE + (4u * j) * 0x00010001u;       // (4j+1) | (4j+3)<<16
                                const uint32_t co = ce - 0x00010001u;                  // (4j)   | (4j+2)<<16
                                const uint32_t kEl = __byte_perm(V[j][0], ce, 0x1076);
                                const uint32_t kEh = __byte_perm(V[j][0], ce, 0x3254);
                                const uint32_t kOl = __byte_perm(V[j][1], co, 0x1076);
                                const uint32_t kOh = __byte_perm(V[j][1], co, 0x3254);
                                best = min(best, min(kEl, kEh));
                                best = min(best, min(kOl, kOh));
                            }
                        } else {
                            V[j][0] += (n.x & 0xFFFFu) - (o.x & 0xFFFFu);
                            V[j][1] += (n.x >> 16) - (o.x >> 16);
                            V[j][2] += (n.y & 0xFFFFu) - (o.y & 0xFFFFu);
                            V[j][3] += (n.y >> 16) - (o.y >> 16);
                            if (emit) {
                                best = min(best, min(V[j][0] * 512u + (4u * j + 3u), V[j][1] * 512u + (4u * j + 1u)));
                                best = min(best, min(V[j][2] * 512u + (4u * j + 2u), V[j][3] * 512u + (4u * j + 0u)));
                            }
                        }
                    }
                }
                if (emit) pk[(rb * a.K + kB) * a.TW + xlB] = best + 4u * (uint32_t)(g0 + kB * GT);
            }
        }
        __syncthreads();
    }
    phaseC(nbatches - 1);
}

// Key map -> disparity bytes, used only when the disparity range was split over several CTAs.
__global__ void sad_finalize_kernel(const uint32_t* __restrict__ gkey, uint8_t* __restrict__ out,
                                    int W, int H, int y0, int y1, int pitchOut, long long frameOut, int D)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = y0 + blockIdx.y;
    const int f = blockIdx.z;
    if (x < W && y < y1) {
        const uint32_t d = gkey[((size_t)f * H + y) * W + x] & 511u;
        out[(long long)f * frameOut + (size_t)y * pitchOut + x] = (uint8_t)((d * 255u) / (uint32_t)D);
    }
}

__global__ void sad_fill_kernel(uint32_t* __restrict__ p, size_t n, uint32_t v)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace sadgpu
