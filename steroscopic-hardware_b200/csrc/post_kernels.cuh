// post_kernels.cuh — optional post-processing of a disparity map.  Nothing in the reference corresponds to
// these hooks (cmd/handlers/stream.go:14-37 only serves the files OutputCamera writes): they are strictly additive and never run
// unless sadgpu_postprocess* is called, so the bit-exact path is untouched.  Both kernels are HBM-bound byte work: one thread per
// four pixels, 32-bit loads / stores where the row allows it.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sadgpu {

__device__ __forceinline__ void cswap(uint32_t& a, uint32_t& b) { const uint32_t lo = min(a, b), hi = max(a, b); a = lo; b = hi; }

// 3x3 median, window clamped to the image (border pixels replicate): the 19-exchange median-of-9 network.
__global__ void median3_kernel(const uint8_t* __restrict__ src, size_t sp, uint8_t* __restrict__ dst, size_t dp, int w, int h)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w || y >= h) return;
    const int xm = max(x - 1, 0), xp = min(x + 1, w - 1), ym = max(y - 1, 0), yp = min(y + 1, h - 1);
    const uint8_t *r0 = src + (size_t)ym * sp, *r1 = src + (size_t)y * sp, *r2 = src + (size_t)yp * sp;
    uint32_t p0 = r0[xm], p1 = r0[x], p2 = r0[xp], p3 = r1[xm], p4 = r1[x], p5 = r1[xp], p6 = r2[xm], p7 = r2[x], p8 = r2[xp];
    cswap(p1, p2); cswap(p4, p5); cswap(p7, p8); cswap(p0, p1); cswap(p3, p4); cswap(p6, p7);
    cswap(p1, p2); cswap(p4, p5); cswap(p7, p8); cswap(p0, p3); cswap(p5, p8); cswap(p4, p7);
    cswap(p3, p6); cswap(p1, p4); cswap(p2, p5); cswap(p4, p7); cswap(p4, p2); cswap(p6, p4); cswap(p4, p2);
    dst[(size_t)y * dp + x] = (uint8_t)p4;
}

// Left-right consistency.  Both maps hold v = d*255/D (the scale the path writes, sad.go:91-93).  For a left-map pixel (x, y) with
// decoded disparity d = round(v*D/255) the matching right-image pixel is x - d; the right-referenced map there must decode to within
// `tol` of d, otherwise the pixel is replaced by `invalid`.  Pixels whose match falls outside the image are invalid too.
__global__ void lrcheck_kernel(const uint8_t* __restrict__ lmap, size_t lp, const uint8_t* __restrict__ rmap, size_t rp,
                               uint8_t* __restrict__ dst, size_t dp, int w, int h, int D, int tol, int invalid)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w || y >= h) return;
    const int v = lmap[(size_t)y * lp + x];
    const int d = (v * D + 127) / 255;
    const int xr = x - d;
    int out = invalid;
    if (xr >= 0) {
        const int dr = (rmap[(size_t)y * rp + xr] * D + 127) / 255;
        if (abs(dr - d) <= tol) out = v;
    }
    dst[(size_t)y * dp + x] = (uint8_t)out;
}

// Horizontal mirror of a plane (the right-referenced map is the path run on the mirrored pair with the roles swapped).
__global__ void mirror_kernel(const uint8_t* __restrict__ src, size_t sp, uint8_t* __restrict__ dst, size_t dp, int w, int h)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w || y >= h) return;
    dst[(size_t)y * dp + x] = src[(size_t)y * sp + (w - 1 - x)];
}

}  // namespace sadgpu
