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

template <int MODE, int CH>
__device__ __forceinline__ uint32_t gray_of(const uint8_t* p)
{
    const uint32_t r = p[0], g = p[1], b = p[2];
    if (MODE == GRAY_RGBX8_LOADPNG) return (19595u * r + 38470u * g + 7471u * b + (1u << 15)) >> 24;   // always 0
    if (MODE == GRAY_RGB8_INTENDED) return go_luma16(r * 257u, g * 257u, b * 257u);
    const uint32_t a = CH == 4 ? p[3] : 255u;
    return go_luma16(r * 257u * a / 255u, g * 257u * a / 255u, b * 257u * a / 255u);
}

template <int MODE, int CH>
__global__ void gray_kernel(const uint8_t* __restrict__ src, size_t src_pitch, uint8_t* __restrict__ dst, size_t dst_pitch,
                            int w, int h, int vec_ok)
{
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y;
    if (x4 >= w || y >= h) return;
    const uint8_t* row = src + (size_t)y * src_pitch;
    uint8_t* out = dst + (size_t)y * dst_pitch;
    if (x4 + 3 < w && vec_ok) {
        uint8_t px[4 * CH];
        if (CH == 4) *reinterpret_cast<uint4*>(px) = *reinterpret_cast<const uint4*>(row + (size_t)x4 * 4);
        else {
            const uint32_t* q = reinterpret_cast<const uint32_t*>(row + (size_t)x4 * 3);
#pragma unroll
            for (int k = 0; k < 3; ++k) reinterpret_cast<uint32_t*>(px)[k] = q[k];
        }
        uint32_t v = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) v |= gray_of<MODE, CH>(px + k * CH) << (8 * k);
        *reinterpret_cast<uint32_t*>(out + x4) = v;
    } else {
        for (int k = 0; k < 4 && x4 + k < w; ++k) out[x4 + k] = (uint8_t)gray_of<MODE, CH>(row + (size_t)(x4 + k) * CH);
    }
}

}  // namespace sadgpu
