// zb_edges.cu -- Image.sobel (reference image.zig:999-1009 -> edges.zig:33-73), a composition of the hot path:
// gray f32 plane (as(f32, convertColor(u8, px)), color.zig:1031-1041 for the luma; float scalars pass through), two dense 3x3
// convolutions with .replicate through zb_convolve's f32 kernel (the reference's accumulation order), and the magnitude
// sqrt(gx^2 + gy^2) / 4 truncated into a u8 image.  Everything stays on the device; three scratch planes from the pool.
#include "zb_device.cuh"
#include "zb_internal.h"

namespace zb {

int convolve_generic(const zb_image* src, zb_image* dst, int pixfmt, const float* kernel, int kh, int kw, int border, cudaStream_t s);

namespace {

template <int CH, bool IS_FLOAT>
__global__ void __launch_bounds__(256) to_gray_f32_kernel(const void* __restrict__ src, size_t src_stride, float* __restrict__ gray, int rows,
                                                          int cols) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (c >= cols) return;
    float v;
    if constexpr (IS_FLOAT) {
        v = ((const float*)src)[(size_t)r * src_stride + c];
    } else if constexpr (CH == 1) {
        v = (float)((const uint8_t*)src)[(size_t)r * src_stride + c];
    } else {
        const uint8_t* px = (const uint8_t*)src + ((size_t)r * src_stride + c) * CH;
        const int y = (13933 * (int)px[0] + 46871 * (int)px[1] + 4732 * (int)px[2] + 32768) >> 16;   // rgbToGray(u8)
        v = (float)min(max(y, 0), 255);
    }
    gray[(size_t)r * cols + c] = v;
}

__global__ void __launch_bounds__(256) sobel_magnitude_kernel(const float* __restrict__ gx, const float* __restrict__ gy, uint8_t* __restrict__ dst,
                                                              size_t dst_stride, int rows, int cols) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (c >= cols) return;
    const float a = gx[(size_t)r * cols + c], b = gy[(size_t)r * cols + c];
    const float magnitude = __fsqrt_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)));   // edges.zig:65
    const float scaled = __fdiv_rn(magnitude, 4.0f);                                        // :68
    dst[(size_t)r * dst_stride + c] = (uint8_t)truncf(fmaxf(0.0f, fminf(255.0f, scaled))); // :69
}

// The whole of edges.zig:33-73 in one pass over the image: 9 border-replicated luma samples, both 3x3 correlations with the dense
// convolution's accumulation (acc = acc + px*k, separately rounded, taps in row-major order, zero taps included), magnitude.
// Reads the source once and writes one byte per pixel; values are identical to the composition above by construction.
template <int CH, bool IS_FLOAT>
__global__ void __launch_bounds__(256) sobel_fused_kernel(const void* __restrict__ src, size_t src_stride, uint8_t* __restrict__ dst,
                                                          size_t dst_stride, int rows, int cols) {
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int r = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (c >= cols || r >= rows) return;
    auto luma = [&](int y, int x) -> float {
        y = min(max(y, 0), rows - 1);   // .replicate
        x = min(max(x, 0), cols - 1);
        if constexpr (IS_FLOAT) {
            return ((const float*)src)[(size_t)y * src_stride + x];
        } else if constexpr (CH == 1) {
            return (float)((const uint8_t*)src)[(size_t)y * src_stride + x];
        } else {
            const uint8_t* px = (const uint8_t*)src + ((size_t)y * src_stride + x) * CH;
            const int v = (13933 * (int)px[0] + 46871 * (int)px[1] + 4732 * (int)px[2] + 32768) >> 16;   // rgbToGray(u8)
            return (float)min(max(v, 0), 255);
        }
    };
    const float kx[9] = {-1, 0, 1, -2, 0, 2, -1, 0, 1};   // edges.zig:14-18
    const float ky[9] = {-1, -2, -1, 0, 0, 0, 1, 2, 1};   // :21-25
    float gx = 0.0f, gy = 0.0f;
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float v = luma(r + j - 1, c + i - 1);
            gx = mul_add_unfused(v, kx[3 * j + i], gx);
            gy = mul_add_unfused(v, ky[3 * j + i], gy);
        }
    const float magnitude = __fsqrt_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)));
    const float scaled = __fdiv_rn(magnitude, 4.0f);
    dst[(size_t)r * dst_stride + c] = (uint8_t)truncf(fmaxf(0.0f, fminf(255.0f, scaled)));
}

}  // namespace
}  // namespace zb

using namespace zb;

extern "C" int zb_sobel(const zb_image* src, zb_image* dst, int pixfmt, zb_stream stream) {
    if (!src || !dst) return ZB_ERR_INVALID_ARGUMENT;
    if (pixfmt != ZB_PIX_U8 && pixfmt != ZB_PIX_F32 && pixfmt != ZB_PIX_RGB8 && pixfmt != ZB_PIX_RGBA8) return ZB_ERR_UNSUPPORTED;
    if (src->rows != dst->rows || src->cols != dst->cols) return ZB_ERR_DIMENSION_MISMATCH;   // image.zig:1005
    if (src->rows == 0 || src->cols == 0) return ZB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    const int rows = (int)src->rows, cols = (int)src->cols;
    if (!g_force_generic.load()) {   // one pass; the composition below stays as the cross-check (zb_set_force_generic)
        dim3 g2(div_up(cols, 32), div_up(rows, 8));
        uint8_t* dp = (uint8_t*)dst->data;
        switch (pixfmt) {
            case ZB_PIX_F32: sobel_fused_kernel<1, true><<<g2, 256, 0, s>>>(src->data, src->stride, dp, dst->stride, rows, cols); break;
            case ZB_PIX_U8: sobel_fused_kernel<1, false><<<g2, 256, 0, s>>>(src->data, src->stride, dp, dst->stride, rows, cols); break;
            case ZB_PIX_RGB8: sobel_fused_kernel<3, false><<<g2, 256, 0, s>>>(src->data, src->stride, dp, dst->stride, rows, cols); break;
            default: sobel_fused_kernel<4, false><<<g2, 256, 0, s>>>(src->data, src->stride, dp, dst->stride, rows, cols); break;
        }
        ZB_LAUNCHED();
        t_last_kernel = "sobel_fused";
        return ZB_OK;
    }
    const size_t plane = (size_t)rows * cols * sizeof(float);
    Scratch buf;
    if ((rc = buf.alloc(3 * plane, s))) return rc;
    float* gray = buf.as<float>();
    float* gx = gray + (size_t)rows * cols;
    float* gy = gx + (size_t)rows * cols;
    dim3 grid(div_up(cols, 256), rows);
    switch (pixfmt) {
        case ZB_PIX_F32: to_gray_f32_kernel<1, true><<<grid, 256, 0, s>>>(src->data, src->stride, gray, rows, cols); break;
        case ZB_PIX_U8: to_gray_f32_kernel<1, false><<<grid, 256, 0, s>>>(src->data, src->stride, gray, rows, cols); break;
        case ZB_PIX_RGB8: to_gray_f32_kernel<3, false><<<grid, 256, 0, s>>>(src->data, src->stride, gray, rows, cols); break;
        default: to_gray_f32_kernel<4, false><<<grid, 256, 0, s>>>(src->data, src->stride, gray, rows, cols); break;
    }
    ZB_LAUNCHED();
    static const float sobel_x[9] = {-1, 0, 1, -2, 0, 2, -1, 0, 1};   // edges.zig:14-18
    static const float sobel_y[9] = {-1, -2, -1, 0, 0, 0, 1, 2, 1};   // :21-25
    zb_image g{gray, src->rows, src->cols, src->cols}, ix{gx, src->rows, src->cols, src->cols}, iy{gy, src->rows, src->cols, src->cols};
    if ((rc = convolve_generic(&g, &ix, ZB_PIX_F32, sobel_x, 3, 3, ZB_BORDER_REPLICATE, s))) return rc;
    if ((rc = convolve_generic(&g, &iy, ZB_PIX_F32, sobel_y, 3, 3, ZB_BORDER_REPLICATE, s))) return rc;
    sobel_magnitude_kernel<<<grid, 256, 0, s>>>(gx, gy, (uint8_t*)dst->data, dst->stride, rows, cols);
    ZB_LAUNCHED();
    t_last_kernel = "sobel";
    return ZB_OK;
}
