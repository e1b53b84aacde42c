// zb_convert.cu -- Image.convert / convertInto (the pixel-format stage either side of the path; reference
// image.zig:396-421): out[r, c] = convertColor(TargetType, self[r, c]) (color.zig:108-151) between u8, f32, Rgb(u8), Rgba(u8) and
// Rgba(f32).  One elementwise kernel, templated on the (source, destination) pair; the component rules are
//   u8 -> float component   v / 255                      (Gray / Rgb / Rgba .as(f32), color.zig:369-373, 488-493, 550)
//   float -> u8 component   round(255 * clamp(v, 0, 1))  in f32 (.as(u8), :376-379, 496-501, 553); the scalar f32 -> u8 conversion alone
//                           goes through f64 (:114-118)
//   colour -> gray          Rgb(u8): the 16.16 fixed-point BT.709 luma (:1031-1041); float: clamp(lr r + lg g + lb b, 0, 1) (:1043-1045);
//                           Rgba drops its alpha first (:477-479)
//   gray -> colour          replicate, alpha 255 / 1.0 (:1049-1051, :355)
// with every float product and sum separately rounded, so the result is bit-identical to the reference's.
#include "zb_device.cuh"
#include "zb_internal.h"

namespace zb {
namespace {

__device__ __forceinline__ float clamp01(float v) { return fmaxf(0.0f, fminf(v, 1.0f)); }
__device__ __forceinline__ uint8_t unit_to_u8(float v) { return (uint8_t)roundf(__fmul_rn(255.0f, clamp01(v))); }
__device__ __forceinline__ float u8_to_unit(uint8_t v) { return __fdiv_rn((float)v, 255.0f); }
__device__ __forceinline__ uint8_t luma_u8(uint8_t r, uint8_t g, uint8_t b) {
    const int y = (13933 * (int)r + 46871 * (int)g + 4732 * (int)b + 32768) >> 16;
    return (uint8_t)min(max(y, 0), 255);
}
__device__ __forceinline__ float luma_f32(float r, float g, float b) {
    return clamp01(__fadd_rn(__fadd_rn(__fmul_rn(0.2126f, r), __fmul_rn(0.7152f, g)), __fmul_rn(0.0722f, b)));
}

template <int FMT> struct Fmt;
template <> struct Fmt<ZB_PIX_U8> { using CT = uint8_t; static constexpr int N = 1; };
template <> struct Fmt<ZB_PIX_F32> { using CT = float; static constexpr int N = 1; };
template <> struct Fmt<ZB_PIX_RGB8> { using CT = uint8_t; static constexpr int N = 3; };
template <> struct Fmt<ZB_PIX_RGBA8> { using CT = uint8_t; static constexpr int N = 4; };
template <> struct Fmt<ZB_PIX_RGBAF32> { using CT = float; static constexpr int N = 4; };

template <int SF, int DF>
__global__ void __launch_bounds__(256) convert_kernel(const void* __restrict__ src, size_t src_stride, void* __restrict__ dst, size_t dst_stride,
                                                      int rows, int cols) {
    using S = Fmt<SF>;
    using D = Fmt<DF>;
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = ZB_GRID_ROW();
    if (c >= cols || r >= rows) return;
    const typename S::CT* sp = (const typename S::CT*)src + ((size_t)r * src_stride + c) * S::N;
    typename D::CT* dp = (typename D::CT*)dst + ((size_t)r * dst_stride + c) * D::N;
    constexpr bool s_float = sizeof(typename S::CT) == 4, d_float = sizeof(typename D::CT) == 4;
    typename S::CT s[S::N];
#pragma unroll
    for (int k = 0; k < S::N; ++k) s[k] = sp[k];
    typename D::CT d[D::N];
    if constexpr (S::N == 1 && D::N == 1) {                       // scalar <-> scalar, color.zig:113-118
        if constexpr (!s_float) {
            d[0] = __fdiv_rn((float)s[0], 255.0f);
        } else {
            double v = (double)s[0];
            v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
            d[0] = (uint8_t)round(v * 255.0);
        }
    } else if constexpr (S::N == 1) {                             // scalar -> colour, :121-131
        if constexpr (d_float) {
            float y;
            if constexpr (s_float) y = s[0]; else y = u8_to_unit(s[0]);
            d[0] = d[1] = d[2] = y;
            d[3] = 1.0f;
        } else {
            uint8_t y;
            if constexpr (s_float) y = unit_to_u8(s[0]); else y = s[0];
            d[0] = d[1] = d[2] = y;
            if constexpr (D::N == 4) d[3] = 255;
        }
    } else if constexpr (D::N == 1) {                             // colour -> scalar, :133-136
        if constexpr (s_float) {
            const float y = luma_f32(s[0], s[1], s[2]);
            if constexpr (d_float) d[0] = y; else d[0] = unit_to_u8(y);
        } else {
            const uint8_t y = luma_u8(s[0], s[1], s[2]);
            if constexpr (d_float) d[0] = u8_to_unit(y); else d[0] = y;
        }
    } else if constexpr (d_float) {                               // 8-bit colour -> Rgba(f32): source.as(f32).to(.rgba), :144-148
#pragma unroll
        for (int k = 0; k < 3; ++k) d[k] = u8_to_unit(s[k]);
        if constexpr (S::N == 4) d[3] = u8_to_unit(s[3]); else d[3] = 1.0f;
    } else if constexpr (s_float) {                               // Rgba(f32) -> 8-bit colour: source.to(space).as(u8), :150
#pragma unroll
        for (int k = 0; k < D::N; ++k) d[k] = unit_to_u8(s[k]);
    } else {                                                      // Rgb(u8) <-> Rgba(u8)
#pragma unroll
        for (int k = 0; k < 3; ++k) d[k] = s[k];
        if constexpr (D::N == 4) d[3] = 255;
    }
#pragma unroll
    for (int k = 0; k < D::N; ++k) dp[k] = d[k];
}

template <int SF, int DF>
int launch(const zb_image* src, zb_image* dst, cudaStream_t s) {
    const dim3 grid = row_grid(div_up(src->cols, 256), src->rows);
    convert_kernel<SF, DF><<<grid, 256, 0, s>>>(src->data, src->stride, dst->data, dst->stride, (int)src->rows, (int)src->cols);
    ZB_LAUNCHED();
    return ZB_OK;
}

template <int SF>
int launch_from(const zb_image* src, zb_image* dst, int df, cudaStream_t s) {
    switch (df) {
        case ZB_PIX_U8: if constexpr (SF != ZB_PIX_U8) return launch<SF, ZB_PIX_U8>(src, dst, s); break;
        case ZB_PIX_F32: if constexpr (SF != ZB_PIX_F32) return launch<SF, ZB_PIX_F32>(src, dst, s); break;
        case ZB_PIX_RGB8: if constexpr (SF != ZB_PIX_RGB8) return launch<SF, ZB_PIX_RGB8>(src, dst, s); break;
        case ZB_PIX_RGBA8: if constexpr (SF != ZB_PIX_RGBA8) return launch<SF, ZB_PIX_RGBA8>(src, dst, s); break;
        case ZB_PIX_RGBAF32: if constexpr (SF != ZB_PIX_RGBAF32) return launch<SF, ZB_PIX_RGBAF32>(src, dst, s); break;
    }
    return ZB_ERR_UNSUPPORTED;
}

}  // namespace
}  // namespace zb

using namespace zb;

extern "C" int zb_convert(const zb_image* src, int src_pixfmt, zb_image* dst, int dst_pixfmt, zb_stream stream) {
    if (!src || !dst) return ZB_ERR_INVALID_ARGUMENT;
    if (channels_of(src_pixfmt) == 0 || channels_of(dst_pixfmt) == 0) return ZB_ERR_UNSUPPORTED;
    if (src->rows != dst->rows || src->cols != dst->cols) return ZB_ERR_DIMENSION_MISMATCH;      // image.zig:397
    if (src_pixfmt == dst_pixfmt) return zb_copy(src, dst, src_pixfmt, stream);                  // :398-399
    if (src->rows == 0 || src->cols == 0) return ZB_OK;
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    switch (src_pixfmt) {
        case ZB_PIX_U8: rc = launch_from<ZB_PIX_U8>(src, dst, dst_pixfmt, s); break;
        case ZB_PIX_F32: rc = launch_from<ZB_PIX_F32>(src, dst, dst_pixfmt, s); break;
        case ZB_PIX_RGB8: rc = launch_from<ZB_PIX_RGB8>(src, dst, dst_pixfmt, s); break;
        case ZB_PIX_RGBA8: rc = launch_from<ZB_PIX_RGBA8>(src, dst, dst_pixfmt, s); break;
        default: rc = launch_from<ZB_PIX_RGBAF32>(src, dst, dst_pixfmt, s); break;
    }
    if (rc) return rc;
    t_last_kernel = "convert";
    return ZB_OK;
}
