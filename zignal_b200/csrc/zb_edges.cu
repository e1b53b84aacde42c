// zb_edges.cu -- Image.sobel (reference image.zig:999-1009 -> edges.zig:33-73), a composition of the hot path:
// gray f32 plane (as(f32, convertColor(u8, px)), color.zig:1031-1041 for the luma; float scalars pass through), two dense 3x3
// convolutions with .replicate through zb_convolve's f32 kernel (the reference's accumulation order), and the magnitude
// sqrt(gx^2 + gy^2) / 4 truncated into a u8 image.  Everything stays on the device; three scratch planes from the pool.
#include <cmath>
#include <vector>

#include "zb_conv.h"
#include "zb_device.cuh"
#include "zb_internal.h"

namespace zb {

int gaussian_taps_host(float sigma, std::vector<float>& taps);   // zb_api_conv.cu (image.zig:972-990; edges.zig:663-681 is the same formula)

namespace {

template <int CH, bool IS_FLOAT>
__global__ void __launch_bounds__(256) to_gray_f32_kernel(const void* __restrict__ src, size_t src_stride, float* __restrict__ gray, int rows,
                                                          int cols) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = ZB_GRID_ROW();
    if (c >= cols || r >= rows) return;
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
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = ZB_GRID_ROW();
    if (c >= cols || r >= rows) return;
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

// ---- Canny (edges.zig:212-274) ----------------------------------------------------------------------------------------------------

// as(f32, convertColor(u8, v)) for a float scalar: round(clamp(v, 0, 1) * 255) evaluated in f64 (color.zig:114-118).
__global__ void __launch_bounds__(256) canny_quantize_f32_kernel(const float* __restrict__ src, size_t src_stride, float* __restrict__ gray,
                                                                 int rows, int cols) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = ZB_GRID_ROW();
    if (c >= cols || r >= rows) return;
    double d = (double)src[(size_t)r * src_stride + c];
    d = d < 0.0 ? 0.0 : (d > 1.0 ? 1.0 : d);
    gray[(size_t)r * cols + c] = (float)(uint8_t)round(d * 255.0);
}

__global__ void __launch_bounds__(256) canny_magnitude_kernel(const float* __restrict__ gx, const float* __restrict__ gy, float* __restrict__ mag,
                                                              size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a = gx[i], b = gy[i];
    mag[i] = __fsqrt_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)));   // edges.zig:263
}

// Non-maximum suppression along the quantised gradient direction (edges.zig:691-763) fused with the double threshold of
// applyHysteresis' first pass (:540-547): dst = 255 for a surviving pixel with magnitude >= high (a seed), 1 for a surviving
// pixel with magnitude >= low (a candidate), 0 otherwise.  The outermost ring of pixels is never marked (:713).
constexpr uint8_t kWeak = 1, kEdge = 255;
__global__ void __launch_bounds__(256) canny_nms_kernel(const float* __restrict__ gx, const float* __restrict__ gy, const float* __restrict__ mag,
                                                        uint8_t* __restrict__ dst, size_t dst_stride, int rows, int cols, float low, float high) {
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int r = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (c >= cols || r >= rows) return;
    uint8_t v = 0;
    if (r >= 1 && c >= 1 && r + 1 < rows && c + 1 < cols) {
        const size_t i = (size_t)r * cols + c;
        const float vx = gx[i], vy = gy[i];
        const float ax = fabsf(vx), ay = fabsf(vy);
        const float K = 0.414213562f;   // tan(22.5 deg), :709
        int dr, dc;                     // first neighbour; the second is the opposite one
        if (ay <= __fmul_rn(K, ax)) { dr = 0; dc = -1; }
        else if (ax <= __fmul_rn(K, ay)) { dr = -1; dc = 0; }
        else if (__fmul_rn(vx, vy) > 0.0f) { dr = -1; dc = 1; }
        else { dr = -1; dc = -1; }
        const float m = mag[i];
        const float n1 = mag[(size_t)(r + dr) * cols + (c + dc)];
        const float n2 = mag[(size_t)(r - dr) * cols + (c - dc)];
        if (m >= n1 && m >= n2) v = m >= high ? kEdge : (m >= low ? kWeak : (uint8_t)0);
    }
    dst[(size_t)r * dst_stride + c] = v;
}

// Hysteresis (edges.zig:549-575).  The reference grows the edge set breadth-first from the seeds through 8-connected candidates; the
// result is the closure "candidate connected to a seed through candidates", which does not depend on the visiting order.  Each
// block relaxes a 64x64 tile (plus a one-pixel ring read from its neighbours) to its local fixed point in shared memory; the
// host repeats the pass until no block promoted a pixel.  Promotions are monotone (1 -> 255), so a ring value read while the
// neighbouring block is still writing is merely early or late, and a late one is caught by the next pass.
constexpr int kHystTile = 64;
__global__ void __launch_bounds__(1024) canny_hysteresis_kernel(uint8_t* img, size_t stride, int rows, int cols, int* __restrict__ promoted) {
    __shared__ uint8_t t[kHystTile + 2][kHystTile + 4];
    const int r0 = blockIdx.y * kHystTile - 1, c0 = blockIdx.x * kHystTile - 1;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    int weak_here = 0;
    for (int i = tid; i < (kHystTile + 2) * (kHystTile + 2); i += 1024) {
        const int y = i / (kHystTile + 2), x = i - y * (kHystTile + 2);
        const int gr = r0 + y, gc = c0 + x;
        uint8_t v = 0;
        if (gr >= 0 && gr < rows && gc >= 0 && gc < cols) v = ((const volatile uint8_t*)img)[(size_t)gr * stride + gc];
        t[y][x] = v;
        weak_here |= (v == kWeak);
    }
    if (!__syncthreads_or(weak_here)) return;     // nothing left to decide in this tile
    int any = 0, changed;
    do {
        changed = 0;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const int y = 1 + 2 * threadIdx.y + dy, x = 1 + 2 * threadIdx.x + dx;
                if (t[y][x] != kWeak) continue;
                const bool nb = t[y - 1][x - 1] == kEdge || t[y - 1][x] == kEdge || t[y - 1][x + 1] == kEdge || t[y][x - 1] == kEdge ||
                                t[y][x + 1] == kEdge || t[y + 1][x - 1] == kEdge || t[y + 1][x] == kEdge || t[y + 1][x + 1] == kEdge;
                if (nb) { t[y][x] = kEdge; changed = 1; }
            }
        changed = __syncthreads_or(changed);
        any |= changed;
    } while (changed);
    if (!any) return;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            const int y = 1 + 2 * threadIdx.y + dy, x = 1 + 2 * threadIdx.x + dx;
            const int gr = r0 + y, gc = c0 + x;
            if (gr < rows && gc < cols && t[y][x] == kEdge) img[(size_t)gr * stride + gc] = kEdge;
        }
    if (tid == 0) *promoted = 1;
}

// Candidates that were never reached stay 0 in the reference's output.
__global__ void __launch_bounds__(256) canny_finalize_kernel(uint8_t* __restrict__ img, size_t stride, int rows, int cols) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = ZB_GRID_ROW();
    if (c >= cols || r >= rows) return;
    uint8_t* p = img + (size_t)r * stride + c;
    if (*p == kWeak) *p = 0;
}

}  // namespace
}  // namespace zb

using namespace zb;

extern "C" int zb_canny(const zb_image* src, zb_image* dst, int pixfmt, float sigma, float low_threshold, float high_threshold,
                        zb_stream stream) {
    if (!src || !dst) return ZB_ERR_INVALID_ARGUMENT;
    if (!std::isfinite(sigma) || !std::isfinite(low_threshold) || !std::isfinite(high_threshold)) return ZB_ERR_INVALID_ARGUMENT;   // :221
    if (sigma < 0) return ZB_ERR_INVALID_SIGMA;                                                                                     // :224
    if (low_threshold < 0 || high_threshold < 0 || low_threshold >= high_threshold) return ZB_ERR_INVALID_THRESHOLD;               // :225-226
    if (pixfmt != ZB_PIX_U8 && pixfmt != ZB_PIX_F32 && pixfmt != ZB_PIX_RGB8 && pixfmt != ZB_PIX_RGBA8) return ZB_ERR_UNSUPPORTED;
    if (src->rows != dst->rows || src->cols != dst->cols) return ZB_ERR_DIMENSION_MISMATCH;
    if (src->rows == 0 || src->cols == 0) return ZB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    const int rows = (int)src->rows, cols = (int)src->cols;
    const size_t n = (size_t)rows * cols, plane = n * sizeof(float);
    const int planes = sigma == 0 ? 4 : 5;
    Scratch buf;
    if ((rc = buf.alloc(planes * plane + 256, s))) return rc;
    float* gray = buf.as<float>();
    float* gx = gray + n;
    float* gy = gx + n;
    float* mag = gy + n;
    float* blurred = sigma == 0 ? gray : mag + n;                      // :241-242 (sigma == 0: no blur)
    int* promoted = (int*)((char*)buf.p + planes * plane);

    const dim3 grid = row_grid(div_up(cols, 256), (size_t)rows);
    switch (pixfmt) {                                                  // :229-236
        case ZB_PIX_F32: canny_quantize_f32_kernel<<<grid, 256, 0, s>>>((const float*)src->data, src->stride, gray, rows, cols); break;
        case ZB_PIX_U8: to_gray_f32_kernel<1, false><<<grid, 256, 0, s>>>(src->data, src->stride, gray, rows, cols); break;
        case ZB_PIX_RGB8: to_gray_f32_kernel<3, false><<<grid, 256, 0, s>>>(src->data, src->stride, gray, rows, cols); break;
        default: to_gray_f32_kernel<4, false><<<grid, 256, 0, s>>>(src->data, src->stride, gray, rows, cols); break;
    }
    ZB_LAUNCHED();
    zb_image g{gray, src->rows, src->cols, src->cols}, b{blurred, src->rows, src->cols, src->cols}, ix{gx, src->rows, src->cols, src->cols},
        iy{gy, src->rows, src->cols, src->cols};
    if (sigma != 0) {                                                  // blurGaussian, :663-687
        std::vector<float> taps;
        if ((rc = gaussian_taps_host(sigma, taps))) return rc;
        if ((int)taps.size() > kMaxTaps) return ZB_ERR_UNSUPPORTED;
        if ((rc = conv_separable_generic(&g, &b, ZB_PIX_F32, taps.data(), (int)taps.size(), taps.data(), (int)taps.size(), ZB_BORDER_REPLICATE, s)))
            return rc;
    }
    static const float sobel_x[9] = {-1, 0, 1, -2, 0, 2, -1, 0, 1};   // edges.zig:14-18
    static const float sobel_y[9] = {-1, -2, -1, 0, 0, 0, 1, 2, 1};   // :21-25
    if ((rc = convolve_generic(&b, &ix, ZB_PIX_F32, sobel_x, 3, 3, ZB_BORDER_REPLICATE, s))) return rc;   // :253
    if ((rc = convolve_generic(&b, &iy, ZB_PIX_F32, sobel_y, 3, 3, ZB_BORDER_REPLICATE, s))) return rc;   // :254
    canny_magnitude_kernel<<<div_up(n, 256), 256, 0, s>>>(gx, gy, mag, n);
    ZB_LAUNCHED();
    uint8_t* out = (uint8_t*)dst->data;
    canny_nms_kernel<<<dim3(div_up(cols, 32), div_up(rows, 8)), 256, 0, s>>>(gx, gy, mag, out, dst->stride, rows, cols, low_threshold,
                                                                             high_threshold);
    ZB_LAUNCHED();
    const dim3 hgrid(div_up(cols, kHystTile), div_up(rows, kHystTile));
    for (;;) {
        int h = 0;
        if (cudaMemsetAsync(promoted, 0, sizeof(int), s) != cudaSuccess) return ZB_ERR_DEVICE_FAILURE;
        canny_hysteresis_kernel<<<hgrid, dim3(32, 32), 0, s>>>(out, dst->stride, rows, cols, promoted);
        ZB_LAUNCHED();
        if (cudaMemcpyAsync(&h, promoted, sizeof(int), cudaMemcpyDeviceToHost, s) != cudaSuccess) return ZB_ERR_DEVICE_FAILURE;
        if (cudaStreamSynchronize(s) != cudaSuccess) return ZB_ERR_DEVICE_FAILURE;
        if (!h) break;
    }
    canny_finalize_kernel<<<grid, 256, 0, s>>>(out, dst->stride, rows, cols);
    ZB_LAUNCHED();
    t_last_kernel = "canny";
    return ZB_OK;
}

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
        if (pixfmt != ZB_PIX_F32 && g_tune_sobel_tile.load()) {   // 8-bit input: shared-memory byte tiles (luma on the way in), integer gradients (zb_conv_tile_u8.cu)
            rc = sobel_tile_u8(src, dst, channels_of(pixfmt), s);
            if (rc != ZB_ERR_UNSUPPORTED) {
                if (rc == ZB_OK) t_last_kernel = "sobel_tile_u8";
                return rc;
            }
        }
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
    const dim3 grid = row_grid(div_up(cols, 256), (size_t)rows);
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
