// gray_kernels.cuh — Go-exact 8-bit luma on the GPU: what the reference does to a decoded
// PNG before the SAD path sees it.  HBM-bound elementwise work: 4 pixels per thread, 16-byte loads for RGBA.
//   NRGBA8  (8-bit RGBA PNG -> *image.NRGBA -> convertGenericToGray, pkg/despair/gray.go:43-58; the same value as
//            color.GrayModel.Convert in pkg/camera/output.go:145,160):
//            c16 = (c8*0x101)*a8/0xff  (Go color.NRGBA.RGBA()),  y = (19595 r16 + 38470 g16 + 7471 b16 + 1<<15) >> 24
//   RGB8 "intended" (8-bit RGB PNG, opaque): c16 = c8*0x101, same luma formula
//   RGB8/RGBA8 "loadpng" (*image.RGBA through convertRGBAToGray, gray.go:20-40): the formula is applied to 8-bit
//            values and shifted by 24, so every pixel becomes 0 — reproduced on request, bit-exactly.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sadgpu {

enum GrayMode { GRAY_NRGBA8 = 0, GRAY_RGB8_INTENDED = 1, GRAY_RGBX8_LOADPNG = 2 };

__device__ __forceinline__ uint32_t go_luma16(uint32_t r, uint32_t g, uint32_t b)
{
    return (19595u * r + 38470u * g + 7471u * b + (1u << 15)) >> 24;     // max 65536*65535+32768 < 2^32
}

// Opaque pixel (alpha 255, or no alpha channel): c16 = c8*257 exactly, so the three multiplications by 257 fold into the
// luma weights.  The weights sum to 65536, hence the sum stays below 65536*257*255 + 2^15 < 2^32.
__device__ __forceinline__ uint32_t go_luma_opaque(uint32_t r, uint32_t g, uint32_t b)
{
    return (19595u * 257u * r + 38470u * 257u * g + 7471u * 257u * b + (1u << 15)) >> 24;
}

template <int MODE, int CH>
__device__ __forceinline__ uint32_t gray_of(const uint8_t* p)
{
    const uint32_t r = p[0], g = p[1], b = p[2];
    if (MODE == GRAY_RGBX8_LOADPNG) return (19595u * r + 38470u * g + 7471u * b + (1u << 15)) >> 24;   // always 0
    if (MODE == GRAY_RGB8_INTENDED || CH == 3) return go_luma_opaque(r, g, b);
    const uint32_t a = p[3];
    if (a == 255u) return go_luma_opaque(r, g, b);          // the common case: three divisions by 255 saved (the kernel is ALU-bound with them)
    const uint32_t a257 = a * 257u;
    return go_luma16(r * a257 / 255u, g * a257 / 255u, b * a257 / 255u);
}

constexpr int kGrayRows = 4;      // rows per thread: four independent 16-byte loads in flight (the kernel is latency-, then HBM-bound)

template <int MODE, int CH>
__global__ void gray_kernel(const uint8_t* __restrict__ src, size_t src_pitch, uint8_t* __restrict__ dst, size_t dst_pitch,
                            int w, int h, int vec_ok)
{
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y0 = blockIdx.y * kGrayRows;
    if (x4 >= w) return;
    if (x4 + 3 < w && vec_ok) {
        uint32_t px[kGrayRows][CH];                         // 4 pixels = CH 32-bit words
#pragma unroll
        for (int k = 0; k < kGrayRows; ++k) {
            const int y = min(y0 + k, h - 1);               // clamped: surplus rows reload the last one
            const uint8_t* row = src + (size_t)y * src_pitch;
            if (CH == 4) {
                const uint4 v = *reinterpret_cast<const uint4*>(row + (size_t)x4 * 4);
                px[k][0] = v.x; px[k][1] = v.y; px[k][2] = v.z; px[k][CH - 1] = v.w;
            } else {
                const uint32_t* q = reinterpret_cast<const uint32_t*>(row + (size_t)x4 * 3);
#pragma unroll
                for (int j = 0; j < 3; ++j) px[k][j] = q[j];
            }
        }
#pragma unroll
        for (int k = 0; k < kGrayRows; ++k) {
            if (y0 + k >= h) break;
            const uint8_t* p = reinterpret_cast<const uint8_t*>(px[k]);
            uint32_t v = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) v |= gray_of<MODE, CH>(p + j * CH) << (8 * j);
            *reinterpret_cast<uint32_t*>(dst + (size_t)(y0 + k) * dst_pitch + x4) = v;
        }
    } else {
        for (int k = 0; k < kGrayRows && y0 + k < h; ++k)
            for (int j = 0; j < 4 && x4 + j < w; ++j)
                dst[(size_t)(y0 + k) * dst_pitch + x4 + j] = (uint8_t)gray_of<MODE, CH>(src + (size_t)(y0 + k) * src_pitch + (size_t)(x4 + j) * CH);
    }
}

}  // namespace sadgpu
