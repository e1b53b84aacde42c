// zb_resize.cu -- Image.resize.
// Reference: interpolation.zig:89-191 (dispatcher: same-shape copy :91-108; Rgb/Rgba u8 -> integer
// plane resizers; everything else -> resizeGeneric :194-214) and channel_ops.zig:144-493 (plane
// resizers: nearest :193, bilinear :144 (truncating /65536), bicubic :217, Catmull-Rom :292,
// Mitchell :368 (B=C=1/3 hard-coded), Lanczos3 :438 (f32 weights, direct sin, 6x6)).
//
// The plane resizers map destination index d to source s = (d+0.5)*ratio-0.5 in f32 and derive tap
// indices (mirror-resolved) and Q8 / f32 weights from it.  Those per-row and per-column quantities
// depend on one coordinate only, so the host computes them once per call (exactly the reference's
// f32 expressions, including libm sin for Lanczos) and uploads two small tables; the kernel is then
// pure integer (or f32 for Lanczos) accumulate in the reference's ky-outer / kx-inner order.  Pixels
// stay interleaved: the reference's split -> plane -> merge computes the same per-channel values.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "zb_host_stage.h"
#include "zb_internal.h"
#include "zb_sample.cuh"
#include "zb_warp.h"

namespace zb {

namespace {

struct TapEntry {   // per destination row or column
    int idx[6];     // mirror-resolved source indices
    int w[6];       // Q8 weights (int) or f32 bit patterns (Lanczos); bilinear: w[0] = frac (Q8)
};

// host replicas of the reference's integer kernels ---------------------------------------------
static inline int32_t cubic_q8(int32_t t) {  // channel_ops.zig:228-244
    const int32_t S = 256, at = t < 0 ? -t : t;
    if (at <= S) { const int32_t t2 = (at * at) / S, t3 = (t2 * at) / S; return S - 2 * t2 + t3; }
    if (at <= 2 * S) { const int32_t t2 = (at * at) / S, t3 = (t2 * at) / S; return 4 * S - 8 * at + 5 * t2 - t3; }
    return 0;
}
static inline int32_t catmull_q8(int32_t t) {  // channel_ops.zig:303-319
    const int32_t S = 256, at = t < 0 ? -t : t;
    if (at <= S) { const int32_t t2 = (at * at) / S, t3 = (t2 * at) / S; return S - (5 * t2) / 2 + (3 * t3) / 2; }
    if (at <= 2 * S) { const int32_t t2 = (at * at) / S, t3 = (t2 * at) / S; return 2 * S - 4 * at + (5 * t2) / 2 - t3 / 2; }
    return 0;
}
static inline int32_t mitchell_q8(int32_t t) {  // channel_ops.zig:378-395
    const int64_t s = 256, s2 = s * s, s3 = s2 * s, at = t < 0 ? -(int64_t)t : (int64_t)t;
    if (at < s) { const int64_t a2 = at * at, a3 = a2 * at; return (int32_t)((21 * a3 - 36 * a2 * s + 16 * s3) / (18 * s2)); }
    if (at < 2 * s) { const int64_t a2 = at * at, a3 = a2 * at; return (int32_t)((-7 * a3 + 36 * a2 * s - 60 * at * s2 + 32 * s3) / (18 * s2)); }
    return 0;
}
static inline float lanczos_direct(float x) {  // channel_ops.zig:446-454
    if (x == 0) return 1.0f;
    const float a = 3.0f;
    if (std::fabs(x) >= a) return 0.0f;
    const float pi_x = 3.14159265358979323846f * x;
    return (a * std::sin(pi_x) * std::sin(pi_x / a)) / (pi_x * pi_x);
}
static inline int mirror_idx(long long i, long long n) {  // border.zig:53-60
    if (i >= 0 && i < n) return (int)i;
    if (n == 1) return 0;
    const long long period = 2 * (n - 1);
    long long m = i % period;
    if (m < 0) m += period;
    return (int)(m >= n ? period - m : m);
}

static void build_table(std::vector<TapEntry>& tab, uint32_t src_n, uint32_t dst_n, int method) {
    tab.resize(dst_n);
    const float ratio = (float)src_n / (float)dst_n;
    for (uint32_t d = 0; d < dst_n; ++d) {
        TapEntry e;
        memset(&e, 0, sizeof(e));
        const float sf = ((float)d + 0.5f) * ratio - 0.5f;
        switch (method) {
            case ZB_INTERP_NEAREST: {  // channel_ops.zig:205-206
                const uint32_t v = (uint32_t)std::round(sf);
                e.idx[0] = (int)(v < src_n - 1 ? v : src_n - 1);
                break;
            }
            case ZB_INTERP_BILINEAR: {  // :161-166
                const long long si = (long long)std::floor(sf);
                e.w[0] = (int)std::trunc((sf - std::floor(sf)) * 256.0f);
                e.idx[0] = mirror_idx(si, src_n);
                e.idx[1] = mirror_idx(si + 1, src_n);
                break;
            }
            case ZB_INTERP_BICUBIC:
            case ZB_INTERP_CATMULL_ROM:
            case ZB_INTERP_MITCHELL: {  // :250-272
                const long long si = (long long)std::floor(sf);
                const int32_t f = (int32_t)std::trunc((sf - std::floor(sf)) * 256.0f);
                for (int k = 0; k < 4; ++k) {
                    e.idx[k] = mirror_idx(si + k - 1, src_n);
                    const int32_t t = k * 256 - 256 - f;
                    e.w[k] = method == ZB_INTERP_BICUBIC ? cubic_q8(t) : (method == ZB_INTERP_CATMULL_ROM ? catmull_q8(t) : mitchell_q8(t));
                }
                break;
            }
            default: {  // Lanczos :460-481
                const long long si = (long long)std::floor(sf);
                const float f = sf - std::floor(sf);
                for (int k = 0; k < 6; ++k) {
                    e.idx[k] = mirror_idx(si + k - 2, src_n);
                    const float w = lanczos_direct((float)(k - 2) - f);
                    memcpy(&e.w[k], &w, 4);
                }
                break;
            }
        }
        tab[d] = e;
    }
}

// MODE: 0 nearest, 1 bilinear, 2 4x4 integer cubic family, 3 Lanczos 6x6 f32
template <int CH, int MODE>
__global__ void __launch_bounds__(256) resize_plane_kernel(const uint8_t* __restrict__ src, size_t src_row_b, uint8_t* __restrict__ dst,
                                                           size_t dst_row_b, int dst_rows, int dst_cols,
                                                           const TapEntry* __restrict__ xt, const TapEntry* __restrict__ yt) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c >= dst_cols) return;
    const TapEntry ex = xt[c];
    const TapEntry ey = yt[r];
    uint8_t* out = dst + (size_t)r * dst_row_b + (size_t)c * CH;
    if constexpr (MODE == 0) {
        const uint8_t* p = src + (size_t)ey.idx[0] * src_row_b + (size_t)ex.idx[0] * CH;
#pragma unroll
        for (int k = 0; k < CH; ++k) out[k] = p[k];
    } else if constexpr (MODE == 1) {
        const int fx = ex.w[0], fy = ey.w[0];
        const uint8_t* r0 = src + (size_t)ey.idx[0] * src_row_b;
        const uint8_t* r1 = src + (size_t)ey.idx[1] * src_row_b;
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            const int tl = r0[(size_t)ex.idx[0] * CH + k], tr = r0[(size_t)ex.idx[1] * CH + k];
            const int bl = r1[(size_t)ex.idx[0] * CH + k], br = r1[(size_t)ex.idx[1] * CH + k];
            const int top = tl * (256 - fx) + tr * fx;
            const int bottom = bl * (256 - fx) + br * fx;
            const int result = (top * (256 - fy) + bottom * fy) / 65536;  // channel_ops.zig:185, truncating
            out[k] = (uint8_t)(result < 0 ? 0 : (result > 255 ? 255 : result));
        }
    } else if constexpr (MODE == 2) {
        int sum[CH];
#pragma unroll
        for (int k = 0; k < CH; ++k) sum[k] = 0;
        int weight_sum = 0;
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) {
            const uint8_t* row = src + (size_t)ey.idx[ky] * src_row_b;
            const int wy = ey.w[ky];
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) {
                const int w = (ex.w[kx] * wy) / 256;  // @divTrunc(wx * wy, SCALE)
                const uint8_t* p = row + (size_t)ex.idx[kx] * CH;
#pragma unroll
                for (int k = 0; k < CH; ++k) sum[k] += (int)p[k] * w;
                weight_sum += w;
            }
        }
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            const int result = weight_sum != 0 ? sum[k] / weight_sum : 0;
            out[k] = (uint8_t)(result < 0 ? 0 : (result > 255 ? 255 : result));
        }
    } else {
        float sum[CH];
#pragma unroll
        for (int k = 0; k < CH; ++k) sum[k] = 0;
        float weight_sum = 0;
#pragma unroll
        for (int ky = 0; ky < 6; ++ky) {
            const uint8_t* row = src + (size_t)ey.idx[ky] * src_row_b;
            const float wy = __int_as_float(ey.w[ky]);
#pragma unroll
            for (int kx = 0; kx < 6; ++kx) {
                const float w = __int_as_float(ex.w[kx]) * wy;
                const uint8_t* p = row + (size_t)ex.idx[kx] * CH;
#pragma unroll
                for (int k = 0; k < CH; ++k) sum[k] += (float)p[k] * w;
                weight_sum += w;
            }
        }
#pragma unroll
        for (int k = 0; k < CH; ++k) out[k] = clamp_u8_from_float(weight_sum != 0 ? sum[k] / weight_sum : 0.0f);
    }
}

// Cubic family with a UNIFORM phase: when src/dst is an integer ratio every destination column has the same fractional source
// position (frac = 0.5 for even ratios), so the 4 x-weights are the same for all columns, likewise the 4 y-weights for all rows,
// and the 16 products w = @divTrunc(wx * wy, 256) and their sum are per-launch constants instead of 16 multiplies + truncating
// divides and 3 integer divisions by a runtime value per pixel.  Values are identical to the general kernel by construction.
struct UniformCubic {
    int w[16];        // [ky][kx]
    int weight_sum;   // > 0
    float rcp;        // 1 / weight_sum
};

template <int CH>
__global__ void __launch_bounds__(256) resize_cubic_uniform_kernel(const uint8_t* __restrict__ src, size_t src_row_b, size_t src_bytes,
                                                                   uint8_t* __restrict__ dst, size_t dst_row_b, int dst_cols,
                                                                   const TapEntry* __restrict__ xt, const TapEntry* __restrict__ yt,
                                                                   const __grid_constant__ UniformCubic u) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c >= dst_cols) return;
    const int4 ix = __ldg(reinterpret_cast<const int4*>(&xt[c]));   // idx[0..3] lead the entry
    const int4 iy = __ldg(reinterpret_cast<const int4*>(&yt[r]));
    const int iyv[4] = {iy.x, iy.y, iy.z, iy.w};
    int sum[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) sum[k] = 0;
    const bool consecutive = ix.y == ix.x + 1 && ix.z == ix.x + 2 && ix.w == ix.x + 3;   // false only where the mirror border folds
    if (consecutive) {
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) {
            const uint8_t* p = src + (size_t)iyv[ky] * src_row_b + (size_t)ix.x * CH;   // 4 * CH contiguous bytes
            uint32_t wds[CH];                                                            // the 4*CH bytes as CH words
            const uintptr_t a = (uintptr_t)p & 3u;
            const uint32_t* q = reinterpret_cast<const uint32_t*>(p - a);
            if (a == 0) {
#pragma unroll
                for (int i = 0; i < CH; ++i) wds[i] = __ldg(q + i);
            } else if ((const uint8_t*)(q + CH + 1) <= src + src_bytes && (const uint8_t*)q >= src) {
                uint32_t t[CH + 1];
#pragma unroll
                for (int i = 0; i <= CH; ++i) t[i] = __ldg(q + i);
#pragma unroll
                for (int i = 0; i < CH; ++i) wds[i] = __funnelshift_r(t[i], t[i + 1], 8 * (unsigned)a);
            } else {   // the aligned window would leave the image: bytes
#pragma unroll
                for (int i = 0; i < CH; ++i) wds[i] = (uint32_t)p[4 * i] | ((uint32_t)p[4 * i + 1] << 8) | ((uint32_t)p[4 * i + 2] << 16) | ((uint32_t)p[4 * i + 3] << 24);
            }
#pragma unroll
            for (int kx = 0; kx < 4; ++kx)
#pragma unroll
                for (int k = 0; k < CH; ++k) {
                    const int b = kx * CH + k;
                    sum[k] += (int)((wds[b >> 2] >> (8 * (b & 3))) & 0xFFu) * u.w[ky * 4 + kx];
                }
        }
    } else {
        const int ixv[4] = {ix.x, ix.y, ix.z, ix.w};
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) {
            const uint8_t* row = src + (size_t)iyv[ky] * src_row_b;
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) {
                const uint8_t* p = row + (size_t)ixv[kx] * CH;
#pragma unroll
                for (int k = 0; k < CH; ++k) sum[k] += (int)p[k] * u.w[ky * 4 + kx];
            }
        }
    }
    uint8_t* out = dst + (size_t)r * dst_row_b + (size_t)c * CH;
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        // @divTrunc(sum, weight_sum) clamped to 0..255: non-positive sums give 0; for positive sums (< 2^24, exact in f32) the
        // float estimate is within one of the quotient and the remainder test fixes it
        int q = 0;
        if (sum[k] > 0) {
            q = __float2int_rz((float)sum[k] * u.rcp);
            const int rem = sum[k] - q * u.weight_sum;
            q += rem >= u.weight_sum ? 1 : (rem < 0 ? -1 : 0);
            q = min(q, 255);
        }
        out[k] = (uint8_t)q;
    }
}

template <int CH>
int launch_plane(const zb_image* src, zb_image* dst, int method, const TapEntry* xt, const TapEntry* yt, cudaStream_t s) {
    dim3 grid(div_up(dst->cols, 256), dst->rows);
    const uint8_t* sp = (const uint8_t*)src->data;
    uint8_t* dp = (uint8_t*)dst->data;
    const size_t sb = (size_t)src->stride * CH, db = (size_t)dst->stride * CH;
    const int dr = (int)dst->rows, dc = (int)dst->cols;
    switch (method) {
        case ZB_INTERP_NEAREST: resize_plane_kernel<CH, 0><<<grid, 256, 0, s>>>(sp, sb, dp, db, dr, dc, xt, yt); break;
        case ZB_INTERP_BILINEAR: resize_plane_kernel<CH, 1><<<grid, 256, 0, s>>>(sp, sb, dp, db, dr, dc, xt, yt); break;
        case ZB_INTERP_LANCZOS: resize_plane_kernel<CH, 3><<<grid, 256, 0, s>>>(sp, sb, dp, db, dr, dc, xt, yt); break;
        default: resize_plane_kernel<CH, 2><<<grid, 256, 0, s>>>(sp, sb, dp, db, dr, dc, xt, yt); break;
    }
    ZB_LAUNCHED();
    return ZB_OK;
}

// resizeGeneric, interpolation.zig:194-214
template <typename CT, int N, int METHOD>
__global__ void __launch_bounds__(256) resize_generic_kernel(SrcView img, CT* __restrict__ dst, size_t dst_stride, int dst_rows,
                                                             int dst_cols, float scale_x, float scale_y, float mb, float mc,
                                                             const float* __restrict__ lut) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c >= dst_cols) return;
    const float src_y = ((float)r + 0.5f) * scale_y - 0.5f;
    const float src_x = ((float)c + 0.5f) * scale_x - 0.5f;
    Pix<CT, N> val;
    if (!interpolate<CT, N, METHOD, ZB_BORDER_MIRROR>(img, src_x, src_y, mb, mc, ZB_BORDER_MIRROR, lut, val)) val = zero_px<CT, N>();
    store_px<CT, N>(dst, (size_t)r * dst_stride + c, val);
}

template <typename CT, int N>
int launch_generic(const zb_image* src, zb_image* dst, int method, float mb, float mc, const float* lut, cudaStream_t s) {
    SrcView v{src->data, (int)src->rows, (int)src->cols, src->stride};
    const float scale_x = (float)src->cols / (float)dst->cols;
    const float scale_y = (float)src->rows / (float)dst->rows;
    dim3 grid(div_up(dst->cols, 256), dst->rows);
    return dispatch_method(method, [&](auto m) -> int {
        resize_generic_kernel<CT, N, decltype(m)::value><<<grid, 256, 0, s>>>(v, (CT*)dst->data, (size_t)dst->stride, (int)dst->rows,
                                                                              (int)dst->cols, scale_x, scale_y, mb, mc, lut);
        ZB_LAUNCHED();
        return ZB_OK;
    });
}

}  // namespace

int resize_dispatch(const zb_image* src, zb_image* dst, int pixfmt, int method, float mb, float mc, cudaStream_t s) {
    if (!src || !dst) return ZB_ERR_INVALID_ARGUMENT;
    if (channels_of(pixfmt) == 0) return ZB_ERR_UNSUPPORTED;
    if (method < ZB_INTERP_NEAREST || method > ZB_INTERP_LANCZOS) return ZB_ERR_INVALID_ARGUMENT;
    if (src->rows == dst->rows && src->cols == dst->cols) return zb_copy(src, dst, pixfmt, (zb_stream)s);  // :91-108
    if (dst->rows == 0 || dst->cols == 0) return ZB_OK;
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    if (src->rows == 0 || src->cols == 0) {  // every sample is null -> zeroes (:205-210)
        ZB_CUDA(cudaMemset2DAsync(dst->data, dst->stride * pixel_bytes(pixfmt), 0, (size_t)dst->cols * pixel_bytes(pixfmt), dst->rows, s));
        return ZB_OK;
    }
    if (pixfmt == ZB_PIX_RGB8 || pixfmt == ZB_PIX_RGBA8) {  // meta.isRgb(T), :111
        std::vector<TapEntry> xt, yt;
        build_table(xt, src->cols, dst->cols, method);
        build_table(yt, src->rows, dst->rows, method);
        Scratch tab;
        if ((rc = tab.alloc((xt.size() + yt.size()) * sizeof(TapEntry), s))) return rc;
        TapEntry* dxt = tab.as<TapEntry>();
        TapEntry* dyt = dxt + xt.size();
        ZB_CUDA(cudaMemcpyAsync(dxt, xt.data(), xt.size() * sizeof(TapEntry), cudaMemcpyHostToDevice, s));
        ZB_CUDA(cudaMemcpyAsync(dyt, yt.data(), yt.size() * sizeof(TapEntry), cudaMemcpyHostToDevice, s));
        if (method == ZB_INTERP_BICUBIC || method == ZB_INTERP_CATMULL_ROM || method == ZB_INTERP_MITCHELL) {
            bool uniform = true;
            for (size_t c = 1; c < xt.size() && uniform; ++c) uniform = memcmp(xt[c].w, xt[0].w, 4 * sizeof(int)) == 0;
            for (size_t r = 1; r < yt.size() && uniform; ++r) uniform = memcmp(yt[r].w, yt[0].w, 4 * sizeof(int)) == 0;
            UniformCubic u;
            u.weight_sum = 0;
            for (int ky = 0; ky < 4; ++ky)
                for (int kx = 0; kx < 4; ++kx) {
                    u.w[ky * 4 + kx] = (xt[0].w[kx] * yt[0].w[ky]) / 256;   // @divTrunc(wx * wy, SCALE), channel_ops.zig:262
                    u.weight_sum += u.w[ky * 4 + kx];
                }
            long long abs_sum = 0;
            for (int i = 0; i < 16; ++i) abs_sum += std::llabs((long long)u.w[i]);
            if (uniform && u.weight_sum > 0 && 255 * abs_sum < (1 << 24)) {
                u.rcp = 1.0f / (float)u.weight_sum;
                dim3 grid(div_up(dst->cols, 256), dst->rows);
                const int ch = pixfmt == ZB_PIX_RGB8 ? 3 : 4;
                const size_t sb = (size_t)src->stride * ch, db = (size_t)dst->stride * ch;
                const size_t src_bytes = (size_t)(src->rows - 1) * sb + (size_t)src->cols * ch;
                if (ch == 3)
                    resize_cubic_uniform_kernel<3><<<grid, 256, 0, s>>>((const uint8_t*)src->data, sb, src_bytes, (uint8_t*)dst->data, db,
                                                                        (int)dst->cols, dxt, dyt, u);
                else
                    resize_cubic_uniform_kernel<4><<<grid, 256, 0, s>>>((const uint8_t*)src->data, sb, src_bytes, (uint8_t*)dst->data, db,
                                                                        (int)dst->cols, dxt, dyt, u);
                ZB_LAUNCHED();
                t_last_kernel = "resize_cubic_uniform_u8";
                return ZB_OK;
            }
        }
        t_last_kernel = "resize_plane_u8";
        return pixfmt == ZB_PIX_RGB8 ? launch_plane<3>(src, dst, method, dxt, dyt, s) : launch_plane<4>(src, dst, method, dxt, dyt, s);
    }
    const float* lut = nullptr;
    if (method == ZB_INTERP_LANCZOS && (rc = lanczos_lut_device(&lut, s))) return rc;
    t_last_kernel = "resize_generic";
    switch (pixfmt) {
        case ZB_PIX_U8: return launch_generic<uint8_t, 1>(src, dst, method, mb, mc, lut, s);
        case ZB_PIX_F32: return launch_generic<float, 1>(src, dst, method, mb, mc, lut, s);
        case ZB_PIX_RGBAF32: return launch_generic<float, 4>(src, dst, method, mb, mc, lut, s);
    }
    return ZB_ERR_UNSUPPORTED;
}

}  // namespace zb

using namespace zb;

extern "C" {

int zb_resize(const zb_image* src, zb_image* dst, int pixfmt, int method, float mb, float mc, zb_stream s) {
    return resize_dispatch(src, dst, pixfmt, method, mb, mc, (cudaStream_t)s);
}

int zb_host_resize(const zb_image* src, zb_image* dst, int pixfmt, int method, float mb, float mc) {
    if (!src || !dst) return ZB_ERR_INVALID_ARGUMENT;
    HostStage st;
    int rc;
    if ((rc = st.begin(src, dst, pixfmt))) return rc;
    if ((rc = resize_dispatch(&st.dsrc, &st.ddst, pixfmt, method, mb, mc, st.stream))) return rc;
    return st.finish(dst, pixfmt);
}

}  // extern "C"
