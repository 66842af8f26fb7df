// sad_fast.cuh — compile-time specialised fast path of the SAD disparity kernel (block_size <= 15,
// i.e. h = block_size/2 <= 7, where every window sum fits 16 bits).  Same arithmetic as
// sad_kernels.cuh (pkg/despair/sad.go:55-95, :205-244 through the separable box filter), but:
//   * every stride is a template constant (immediate-offset LDS/STS, no address arithmetic),
//   * phase A is a fully unrolled TW+2h-step walk (the 2h+1 old terms are SSA values),
//   * phase B keeps the 2h+1 previous rows of H in a REGISTER ring: shared memory carries each
//     H value exactly once (one STS.64 in, one LDS.64 out),
//   * never-evaluated candidates (d > X-h, d > D) are excluded by a constant bias folded into
//     the running vertical sum (0x8000 per 16-bit lane; h <= 5) or a lane mask (h = 6,7),
//   * several frames are processed by one launch (blockIdx.z = frame x chunk) so that the grid
//     is many waves deep and the single-wave tail disappears.
#pragma once
#include "sad_common.cuh"

namespace sadgpu {

template <int HALF> struct FastTraits {
    static constexpr int WIN = 2 * HALF + 1;
    static constexpr int TW = 64;                                   // output columns per CTA
    static constexpr int TWP = 65;                                  // odd row stride: conflict-free 64-bit access
    static constexpr int NSTEP = TW + 2 * HALF;                     // phase-A walk length
    static constexpr int LW = (NSTEP + 3) & ~3;                     // words per replicated-L row
    static constexpr int RB = WIN >= 7 ? WIN : WIN * ((8 + WIN - 1) / WIN);   // rows per batch (multiple of WIN)
    static constexpr bool BIAS = WIN * WIN * 255 + 32768 < 65536;   // h <= 5
    static constexpr int OFF = walk_off(HALF);                      // byte phase of the R walk in its aligned word
    static constexpr int NWALKW = walk_words(HALF, NSTEP);          // aligned R words one walk touches
};

template <int HALF, int NGC> struct FastCfg : FastTraits<HALF> {
    using T = FastTraits<HALF>;
    // groups per phase-B thread: bounded by the register ring (2*GT*WIN registers)
    static constexpr int GT = HALF <= 4 ? (NGC == 33 ? 5 : 3) : (NGC == 18 ? 3 : 2);
    static constexpr int K = (NGC + GT - 1) / GT;                   // phase-B threads per column
    static constexpr int NGP = K * GT;                              // group slots in H
    static constexpr int NT = T::TW * K;                            // threads per CTA
    static constexpr int RW = NGC - 1 + T::NWALKW;                  // words per aligned-R row
    static constexpr int H_BYTES = ((T::RB * NGP * T::TWP * 8 + 15) / 16) * 16;
    static constexpr int L_BYTES = T::RB * T::LW * 4;
    static constexpr int R_BYTES = T::RB * RW * 4;
    static constexpr int PK_BYTES = T::RB * K * T::TW * 4;
    static constexpr int LUT_BYTES = ((4 * NGC * 8 + 15) / 16) * 16 > 1040 ? 1040 : 1040;   // up to 4*65+ entries
    static constexpr int OFF_L = H_BYTES;
    static constexpr int OFF_R = OFF_L + L_BYTES;
    static constexpr int OFF_PK = OFF_R + R_BYTES;
    static constexpr int OFF_LUT = OFF_PK + PK_BYTES;
    static constexpr int SMEM = OFF_LUT + LUT_BYTES;
};

template <int HALF, int NGC>
__global__ void __launch_bounds__(FastCfg<HALF, NGC>::NT, 1) sad_fast_kernel(const FastArgs a)
{
    using C = FastCfg<HALF, NGC>;
    constexpr int WIN = C::WIN, TW = C::TW, TWP = C::TWP, RB = C::RB, GT = C::GT, K = C::K, NGP = C::NGP, NT = C::NT;
    extern __shared__ __align__(16) unsigned char smem[];
    uint2* Hs = reinterpret_cast<uint2*>(smem);
    uint32_t* Lrep = reinterpret_cast<uint32_t*>(smem + C::OFF_L);
    uint32_t* Ral = reinterpret_cast<uint32_t*>(smem + C::OFF_R);
    uint32_t* pk = reinterpret_cast<uint32_t*>(smem + C::OFF_PK);
    uint8_t* lut = smem + C::OFF_LUT;

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
    // group slots >= NGC are never written by phase A: keep them zero (their lanes are biased/masked out)
    if (NGP > NGC)
        for (int idx = tid; idx < RB * (NGP - NGC) * TWP; idx += NT) {
            const int rb = idx / ((NGP - NGC) * TWP), rem = idx - rb * ((NGP - NGC) * TWP);
            Hs[(rb * NGP + NGC) * TWP + rem] = make_uint2(0u, 0u);
        }

    // ---- phase-B identity, running sums, lane validity ------------------------------------
    const int kB = tid / TW, xlB = tid - kB * TW;
    const int xB = x0 + xlB;
    uint32_t VE[GT], VO[GT], mE[GT], mO[GT];
    uint32_t ringE[WIN][GT], ringO[WIN][GT];
#pragma unroll
    for (int j = 0; j < GT; ++j) {
        const int dbase = 4 * (g0 + kB * GT + j);
        // largest evaluated disparity of this column; padding slots (>= NGC) never hold a candidate
        const int dmax = (kB * GT + j < NGC) ? min(a.D, xB - HALF) : -1;
        const uint32_t iE = (dbase + 3 > dmax ? 0x0000FFFFu : 0u) | (dbase + 1 > dmax ? 0xFFFF0000u : 0u);
        const uint32_t iO = (dbase + 2 > dmax ? 0x0000FFFFu : 0u) | (dbase + 0 > dmax ? 0xFFFF0000u : 0u);
        mE[j] = iE; mO[j] = iO;
        VE[j] = C::BIAS ? (iE & 0x80008000u) : 0u;
        VO[j] = C::BIAS ? (iO & 0x80008000u) : 0u;
#pragma unroll
        for (int r = 0; r < WIN; ++r) { ringE[r][j] = 0; ringO[r][j] = 0; }
    }
    const uint32_t keybase = 4u * (uint32_t)(g0 + kB * GT);
    const uint32_t k16 = opaque(a.k65536), mhi = opaque(a.k65536 * 0xFFFFu);      // 65536, 0xFFFF0000

    const int r0 = yb0 - HALF;
    const int nrows = (yb1 - yb0) + 2 * HALF;
    const int nbatches = (nrows + RB - 1) / RB;
    const int xr0 = x0 - HALF - 3 - 4 * (g0 + NGC - 1) - C::OFF;   // aligned origin of the R tile (multiple of 4)
    const int nvalid = a.W - (x0 - HALF);

    // Tile loads are split in two: the global loads of the next batch are issued into registers at the start of phase B and
    // written to shared memory at its end, so that their latency (HBM when the frames stream) overlaps phase B.
    constexpr int NLE = (RB * C::LW + NT - 1) / NT, NRE = (RB * C::RW + NT - 1) / NT;
    static_assert(NLE <= 4, "left pixels of a thread are packed into one register");
    constexpr bool ASYNC_R = HALF >= 6;       // the register ring of h = 6, 7 leaves no room for right words in flight: cp.async instead
    uint32_t tl = 0, tr[NRE];
    auto issue_tiles = [&](int rbase) {
        tl = 0;
#pragma unroll
        for (int q = 0; q < NLE; ++q) {
            const int idx = tid + q * NT;
            const int rb = idx / C::LW, i = idx - rb * C::LW;
            const int y = rbase + rb, x = x0 - HALF + i;
            uint32_t v = 0;
            if (idx < RB * C::LW && (unsigned)y < (unsigned)a.H && (unsigned)x < (unsigned)a.W) v = Lg[(size_t)y * a.pitchL + x];
            tl |= v << (8 * q);
        }
#pragma unroll
        for (int q = 0; q < NRE; ++q) {
            const int idx = tid + q * NT;
            const int rb = idx / C::RW, j = idx - rb * C::RW;
            const int y = rbase + rb, x = xr0 + 4 * j;
            uint32_t v = 0;
            bool async = false;
            if (idx < RB * C::RW && (unsigned)y < (unsigned)a.H && x + 3 >= 0 && x < a.W) {
                const uint8_t* p = Rg + (size_t)y * a.pitchR;
                if (a.aligned && x >= 0 && x + 3 < a.W) {
                    if (ASYNC_R) {          // no register is held across phase B: the word goes straight to shared memory
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((uint32_t)__cvta_generic_to_shared(Ral + idx)), "l"(p + x) : "memory");
                        async = true;
                    } else v = *reinterpret_cast<const uint32_t*>(p + x);
                } else {
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if ((unsigned)(x + b) < (unsigned)a.W) v |= (uint32_t)p[x + b] << (8 * b);
                }
            }
            if (ASYNC_R) { if (!async && idx < RB * C::RW) Ral[idx] = v; }      // borders and zero rows: stored at once (the tiles are free during phase B)
            else tr[q] = v;
        }
        if (ASYNC_R) asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto commit_tiles = [&]() {
#pragma unroll
        for (int q = 0; q < NLE; ++q) { const int idx = tid + q * NT; if (idx < RB * C::LW) Lrep[idx] = ((tl >> (8 * q)) & 0xFFu) * 0x01010101u; }
#pragma unroll
        if (ASYNC_R) asm volatile("cp.async.wait_group 0;" ::: "memory");
        else {
#pragma unroll
            for (int q = 0; q < NRE; ++q) { const int idx = tid + q * NT; if (idx < RB * C::RW) Ral[idx] = tr[q]; }
        }
    };

    auto phaseC = [&](int batch) {
        const int rbase = r0 + batch * RB;
        for (int idx = tid; idx < RB * TW; idx += NT) {
            const int rb = idx / TW, xl = idx - rb * TW;
            const int rel = batch * RB + rb, y = rbase + rb - HALF, x = x0 + xl;
            if (rel < 2 * HALF || y >= yb1 || x >= a.W) continue;
            uint32_t best = 0xFFFFFFFFu;
#pragma unroll
            for (int k = 0; k < K; ++k) best = min(best, pk[(rb * K + k) * TW + xl]);
            if (x < HALF) best = 0;                                // sad.go:212-218: both windows clamp, d = 0 wins
            if (a.NC == 1) {
                (a.out + (long long)frame * a.frameOut)[(size_t)y * a.pitchOut + x] = lut[best & 0xFFFFu];
            } else {
                atomicMin(a.gkey + ((size_t)frame * a.H + y) * a.W + x, ((best >> 16) << 9) | (best & 511u));
            }
        }
    };

    issue_tiles(r0); commit_tiles();
    __syncthreads();
    for (int batch = 0; batch < nbatches; ++batch) {
        const int rbase = r0 + batch * RB;
        // ---- phase A (and phase C of the previous batch) ----
        if (batch > 0) phaseC(batch - 1);
        for (int item = tid; item < RB * NGC; item += NT) {
            const int rb = item / NGC, gl = item - rb * NGC;
            const uint32_t* Lr = Lrep + rb * C::LW;
            const uint32_t* Rr = Ral + rb * C::RW + (NGC - 1 - gl);
            uint2* Hout = Hs + (rb * NGP + gl) * TWP;
            if (nvalid >= C::NSTEP) sad_walk<HALF, TW, false>(Lr, Rr, Hout, nvalid);
            else                    sad_walk<HALF, TW, true>(Lr, Rr, Hout, nvalid);
        }
        __syncthreads();
        // ---- phase B (and the tile load of the next batch) ----
        if (batch + 1 < nbatches) issue_tiles(rbase + RB);
        {
            const uint2* Hp = Hs + (kB * GT) * TWP + xlB;
#pragma unroll
            for (int rb = 0; rb < RB; ++rb) {
                const int rel = batch * RB + rb;
                const bool emit = (batch > 0 || rb >= 2 * HALF) && (r0 + rel - HALF) < yb1;
                uint32_t best = 0xFFFFFFFFu;
#pragma unroll
                for (int j = 0; j < GT; ++j) {
                    const uint2 n = Hp[(rb * NGP + j) * TWP];
                    VE[j] = VE[j] + n.x - ringE[rb % WIN][j]; ringE[rb % WIN][j] = n.x;
                    VO[j] = VO[j] + n.y - ringO[rb % WIN][j]; ringO[rb % WIN][j] = n.y;
                }
                if (emit) {
#pragma unroll
                    for (int j = 0; j < GT; ++j) {
                        const uint32_t ve = C::BIAS ? VE[j] : (VE[j] | mE[j]);
                        const uint32_t vo = C::BIAS ? VO[j] : (VO[j] | mO[j]);
                        const uint32_t kEl = key_lo(ve, k16, 4u * j + 3u);
                        const uint32_t kEh = key_hi(ve, mhi, 4u * j + 1u);
                        const uint32_t kOl = key_lo(vo, k16, 4u * j + 2u);
                        const uint32_t kOh = key_hi(vo, mhi, 4u * j + 0u);
                        best = min(best, min(kEl, kEh));
                        best = min(best, min(kOl, kOh));
                    }
                    pk[(rb * K + kB) * TW + xlB] = best + keybase;
                }
            }
        }
        if (batch + 1 < nbatches) commit_tiles();          // phase A of this batch is behind the barrier above: the tiles are free
        __syncthreads();
    }
    phaseC(nbatches - 1);
}

}  // namespace sadgpu
