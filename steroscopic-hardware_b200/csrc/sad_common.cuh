// sad_common.cuh — pieces shared by every SAD kernel of libsadgpu.so: the launch arguments, the packed-key helpers,
// the (row, disparity group) walk and the two trivial kernels around a chunked disparity range.
//
// What all kernels compute, bit-exactly (pkg/despair/sad.go:55-95 scan + :205-244 SumAbsoluteDifferences, through the
// equivalent zero-padded separable box filter):
//     AD_d(x,y) = |L(x,y) - R(x-d,y)|   (0 outside the image)
//     S_d(X,Y)  = sum over the (2h+1)^2 window of AD_d,  h = block_size/2
//     out(X,Y)  = (argmin_{d in [0, min(D, X-h)]} S_d, lowest d on ties) * 255 / D ;  0 for X < h
// Four disparities share a 32-bit word: group g holds d = 4g+3-byte, so one VABSDIFF4 of a replicated left pixel
// against the right word at x-4g-3 evaluates four candidates of one pixel.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

namespace sadgpu {

struct FastArgs {
    CUtensorMap tmapL, tmapR;                // TMA descriptors (warp-specialised kernel, use_tma != 0); 64-byte aligned, first members
    const uint8_t* L; const uint8_t* R; uint8_t* out; uint32_t* gkey;
    long long frameL, frameR, frameOut;      // byte strides between frames of a batch
    int pitchL, pitchR, pitchOut;
    int W, H, y0, y1;
    int D, NG, NC, BH;
    int aligned;                             // R rows may be fetched with aligned 32-bit loads
    int use_tma;                             // tile loads by cp.async.bulk.tensor (needs 16-byte aligned base / pitch / frame stride)
    int debug_skip;                          // developer builds (-DSADGPU_PROFILE) only: 1 = walkers idle, 2 = consumers idle, 4 = cycle counters
    unsigned k65536;                         // = 65536, passed at run time so that v*65536+idx stays an IMAD (FMA pipe)
};

// Keys (sum << 16 | index).  Low lane: one IMAD (FMA pipe) with the multiplier 65536 held in a register and the
// index as immediate addend; high lane: one LOP3 (ALU pipe), (v & mask) | index, mask held in a register.
// Both constants are made opaque so that ptxas keeps them in registers instead of re-materialising them.
__device__ __forceinline__ uint32_t opaque(uint32_t v) { asm volatile("" : "+r"(v)); return v; }
__device__ __forceinline__ uint32_t key_lo(uint32_t v, uint32_t k65536, uint32_t idx) { return v * k65536 + idx; }
__device__ __forceinline__ uint32_t key_hi(uint32_t v, uint32_t maskhi, uint32_t idx) { return (v & maskhi) | idx; }

// Byte phase of the right-image walk inside its aligned word: the walk of group g starts at column x0-h-3-4g.
__host__ __device__ constexpr int walk_off(int half) { return ((-(half + 3)) % 4 + 4) % 4; }
// Aligned right words one walk of `nstep` steps touches.
__host__ __device__ constexpr int walk_words(int half, int nstep) { return ((nstep - 1 + walk_off(half)) >> 2) + 2; }

// One (row, group) walk: NOUT outputs, NOUT + 2h steps, fully unrolled (the 2h+1 old terms are SSA values).
//   Lr   replicated left pixels of the row (word i = pixel of step i in all four bytes), 16-byte aligned
//   Rr   aligned right words such that the four bytes needed at step i start at byte i + walk_off(h) of Rr[0]
//   Hout receives (E, O) = 16x2-packed horizontal window sums: E = (d=4g+3 | d=4g+1 << 16), O = (d=4g+2 | d=4g << 16)
// EDGE: steps i >= nvalid lie at columns x >= W and contribute nothing (sad.go:231-233).
// PF: the loads of the next four steps are issued four steps early.  A shared-memory load cannot be hoisted over the H stores of the
// steps before it (they may alias), so without PF every load is issued right where its value is needed.  It pays where a sub-partition
// has ONE walker warp (sad_wsr.cuh); with three walker warps per sub-partition (sad_ws.cuh) they cover each other and the five extra
// registers spill at setmaxnreg 40 (measured: no gain there).
// FMAS: 0 = both running sums are three-input adds (ALU pipe); 1 = the E sum, 2 = both sums are two multiply-adds by +1 / -1 held in
// registers (FMA pipe): two instructions instead of one, but where a sub-partition's ALU pipe is the bound (one walker warp issuing
// nothing but ALU instructions, sad_wsr.cuh) they come off the critical pipe.
template <int HALF, int NOUT, bool EDGE, bool PF = false, int FMAS = 0>
__device__ __forceinline__ void sad_walk(const uint32_t* __restrict__ Lr, const uint32_t* __restrict__ Rr,
                                         uint2* __restrict__ Hout, int nvalid, bool store = true, uint32_t one = 1u, uint32_t mone = 0xFFFFFFFFu)
{
    constexpr int WIN = 2 * HALF + 1, NS = NOUT + 2 * HALF, OFF = walk_off(HALF);
    uint32_t e[NS], o[NS];
    uint32_t hE = 0, hO = 0, w0 = 0, w1 = 0, wn = 0;
    uint4 lv = make_uint4(0, 0, 0, 0), lvn = make_uint4(0, 0, 0, 0);
    if (PF) {
        lvn = *reinterpret_cast<const uint4*>(Lr);
        w0 = Rr[0]; w1 = Rr[1];
        if (NS > 4 - OFF) wn = Rr[2];
    }
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        const int bi = i + OFF;
        if (PF) {
            if ((i & 3) == 0) { lv = lvn; if (i + 4 < NS) lvn = *reinterpret_cast<const uint4*>(Lr + i + 4); }
            if (i > 0 && (bi & 3) == 0) { w0 = w1; w1 = wn; if (i + 4 < NS) wn = Rr[(bi >> 2) + 2]; }
        } else {
            if ((i & 3) == 0) lv = *reinterpret_cast<const uint4*>(Lr + i);
            if (i == 0) { w0 = Rr[bi >> 2]; w1 = Rr[(bi >> 2) + 1]; }
            else if ((bi & 3) == 0) { w0 = w1; w1 = Rr[(bi >> 2) + 1]; }
        }
        const uint32_t lw = (i & 3) == 0 ? lv.x : (i & 3) == 1 ? lv.y : (i & 3) == 2 ? lv.z : lv.w;
        const uint32_t rw = (bi & 3) == 0 ? w0 : __funnelshift_r(w0, w1, 8 * (bi & 3));
        uint32_t ad = __vabsdiffu4(lw, rw);
        if (EDGE) ad = (i < nvalid) ? ad : 0u;
        e[i] = __byte_perm(ad, 0u, 0x4240);                     // (d=4g+3 | d=4g+1 << 16)
        o[i] = __byte_perm(ad, 0u, 0x4341);                     // (d=4g+2 | d=4g   << 16)
        if (FMAS >= 1) { hE = e[i] * one + hE; if (i >= WIN) hE = e[i - WIN] * mone + hE; }
        else           { if (i >= WIN) hE = hE + e[i] - e[i - WIN]; else hE += e[i]; }
        if (FMAS >= 2) { hO = o[i] * one + hO; if (i >= WIN) hO = o[i - WIN] * mone + hO; }
        else           { if (i >= WIN) hO = hO + o[i] - o[i - WIN]; else hO += o[i]; }
        if (i >= 2 * HALF && store) Hout[i - 2 * HALF] = make_uint2(hE, hO);
    }
}

// ---- chunked disparity ranges: the chunks of a pixel meet in a global key map (sum << 9 | d) through atomicMin ----
// d * 255 / D without a division: magic = 2^32 / D + 1 is exact for every numerator below 2^24 (D >= 2).
__device__ __forceinline__ uint32_t scale_disparity(uint32_t d, uint32_t D, uint32_t magic) { return D == 1 ? d * 255u : __umulhi(d * 255u, magic); }

// Four pixels per thread (16-byte key loads, one 4-byte store) when the planes allow it (vec4 != 0), else one.
__global__ void sad_finalize_kernel(const uint32_t* __restrict__ gkey, uint8_t* __restrict__ out,
                                    int W, int H, int y0, int y1, int pitchOut, long long frameOut, int D, uint32_t magic, int vec4)
{
    const int y = y0 + blockIdx.y;
    const int f = blockIdx.z;
    if (y >= y1) return;
    const uint32_t* k = gkey + ((size_t)f * H + y) * W;
    uint8_t* o = out + (long long)f * frameOut + (size_t)y * pitchOut;
    if (vec4) {
        const int x = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
        if (x < W) {
            const uint4 v = *reinterpret_cast<const uint4*>(k + x);
            const uint32_t a = scale_disparity(v.x & 511u, D, magic), b = scale_disparity(v.y & 511u, D, magic);
            const uint32_t c = scale_disparity(v.z & 511u, D, magic), d = scale_disparity(v.w & 511u, D, magic);
            *reinterpret_cast<uint32_t*>(o + x) = a | (b << 8) | (c << 16) | (d << 24);
        }
    } else {
        const int x = blockIdx.x * blockDim.x + threadIdx.x;
        if (x < W) o[x] = (uint8_t)scale_disparity(k[x] & 511u, D, magic);
    }
}

__global__ void sad_fill_kernel(uint32_t* __restrict__ p, size_t n, uint32_t v)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace sadgpu
