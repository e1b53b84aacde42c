// zb_fdm.cu -- FeatureDistributionMatching (reference fdm.zig:19-275) on the device.
//
// The reference makes two sequential full passes per image: a Welford mean/co-moment stream
// (stats.zig:261-280, a ~10-flop dependency chain per pixel) and the per-pixel affine colour map.
// Here the statistics pass is a parallel reduction of EXACT integer moments of the u8 pixels
// (n, sum x_i, sum x_i x_j as u64; integer atomics, so the result is order-independent), finished on
// the host in f64: cov_ij = (n*Sij - Si*Sj) / (n (n-1) 255^2) with the numerator exact in 128 bits.
// That differs from Welford only by Welford's own rounding noise (~1e-16 relative).  The 3x3 SVDs,
// W = Us * diag(sqrt(lt/ls)) * Ut^T and the bias follow fdm.zig:199-254 verbatim (host, f64); the map
// (fdm.zig:257-271) is one f64 kernel: res = r*w0 + g*w1 + b*w2 + bias, round(255*clamp(res,0,1)).
// The 11 moment sums are also the quantities one all-reduce combines when an image is sharded
// across GPUs (zb_fdm_moments / zb_fdm_update_with_moments).
#include <cmath>
#include <cstring>

#include "zb_host_stage.h"
#include "zb_internal.h"
#include "zb_linalg.h"

struct zb_fdm {
    int pixfmt;
    double target_mean[3];
    double target_u[9];
    double target_s[3];
    bool target_is_gray;
    bool has_target, has_source;
    zb_image source;
};

namespace zb {
namespace {

// color.zig:1031-1041 rgbToGray(u8)
__device__ __forceinline__ unsigned rgb_to_gray(unsigned r, unsigned g, unsigned b) {
    int y = (int)((13933u * r + 46871u * g + 4732u * b + 32768u) >> 16);
    return (unsigned)min(max(y, 0), 255);
}

// sums: {n, Sr, Sg, Sb, Srr, Srg, Srb, Sgg, Sgb, Sbb, non_gray}
template <int CH>
__global__ void __launch_bounds__(256) moments_kernel(const uint8_t* __restrict__ img, size_t n_px, int as_luma,
                                                      unsigned long long* __restrict__ sums) {
    unsigned long long acc[11];
#pragma unroll
    for (int i = 0; i < 11; ++i) acc[i] = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t base = (size_t)blockIdx.x * blockDim.x + threadIdx.x; base < n_px; base += stride * 16) {
        unsigned a[11];  // 16 pixels: sums of products <= 16*255^2 fit easily in u32
#pragma unroll
        for (int i = 0; i < 11; ++i) a[i] = 0;
#pragma unroll 4
        for (int j = 0; j < 16; ++j) {
            const size_t px = base + (size_t)j * stride;
            if (px >= n_px) break;
            unsigned r, g, b;
            if constexpr (CH == 1) {
                r = g = b = img[px];
            } else if constexpr (CH == 4) {
                const uchar4 q = reinterpret_cast<const uchar4*>(img)[px];
                r = q.x; g = q.y; b = q.z;
            } else {
                r = img[px * 3]; g = img[px * 3 + 1]; b = img[px * 3 + 2];
            }
            if (CH != 1 && (r != g || g != b)) a[10] += 1;
            if (CH != 1 && as_luma) r = g = b = rgb_to_gray(r, g, b);
            a[0] += 1;
            a[1] += r; a[2] += g; a[3] += b;
            a[4] += r * r; a[5] += r * g; a[6] += r * b;
            a[7] += g * g; a[8] += g * b; a[9] += b * b;
        }
#pragma unroll
        for (int i = 0; i < 11; ++i) acc[i] += a[i];
    }
    __shared__ unsigned long long sh[8][11];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 11; ++i) {
        unsigned long long v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) sh[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < 11) {
        unsigned long long v = 0;
        for (int w = 0; w < 8; ++w) v += sh[w][threadIdx.x];
        atomicAdd(&sums[threadIdx.x], v);
    }
}

struct MapParams {
    double w[9];
    double bias[3];
    double scale, offset;  // gray branch
    int mode;              // 0 colour (fdm.zig:257-271), 1 scalar on u8 (:185-189), 2 gray target on colour source (:191-197)
};

template <int CH>
__global__ void __launch_bounds__(256) fdm_map_kernel(uint8_t* __restrict__ img, size_t n_px, MapParams p) {
    const size_t px = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (px >= n_px) return;
    if constexpr (CH == 1) {
        const double val = (double)img[px] / 255.0;
        const double result = fmax(0.0, fmin(val * p.scale + p.offset, 1.0));
        img[px] = (uint8_t)round(255.0 * result);
    } else {
        uint8_t* q = img + px * CH;
        if (p.mode == 2) {
            const double val = (double)rgb_to_gray(q[0], q[1], q[2]) / 255.0;
            const double result = fmax(0.0, fmin(val * p.scale + p.offset, 1.0));
            const uint8_t res = (uint8_t)round(255.0 * result);
            q[0] = res; q[1] = res; q[2] = res;
            if (CH == 4) q[3] = 0;  // `.{ .r, .g, .b }`: alpha takes its default 0 (color.zig:405, fdm.zig:196)
        } else {
            const double r = (double)q[0] / 255.0, g = (double)q[1] / 255.0, b = (double)q[2] / 255.0;
            const double res0 = r * p.w[0] + g * p.w[3] + b * p.w[6] + p.bias[0];
            const double res1 = r * p.w[1] + g * p.w[4] + b * p.w[7] + p.bias[1];
            const double res2 = r * p.w[2] + g * p.w[5] + b * p.w[8] + p.bias[2];
            q[0] = (uint8_t)round(255.0 * fmax(0.0, fmin(res0, 1.0)));
            q[1] = (uint8_t)round(255.0 * fmax(0.0, fmin(res1, 1.0)));
            q[2] = (uint8_t)round(255.0 * fmax(0.0, fmin(res2, 1.0)));
        }
    }
}

int moments_device(const zb_image* img, int pixfmt, int as_luma, uint64_t* sums11, cudaStream_t s) {
    if (!img || !sums11) return ZB_ERR_INVALID_ARGUMENT;
    if (pixfmt != ZB_PIX_U8 && pixfmt != ZB_PIX_RGB8 && pixfmt != ZB_PIX_RGBA8) return ZB_ERR_UNSUPPORTED;  // fdm.zig:20
    if (img->stride != img->cols) return ZB_ERR_UNSUPPORTED;  // the reference walks image.data linearly (fdm.zig:82)
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    const size_t n_px = (size_t)img->rows * img->cols;
    Scratch d;
    if ((rc = d.alloc(11 * sizeof(uint64_t), s))) return rc;
    ZB_CUDA(cudaMemsetAsync(d.p, 0, 11 * sizeof(uint64_t), s));
    if (n_px > 0) {
        const unsigned blocks = (unsigned)std::min<size_t>((size_t)di.sm_count * 8, (n_px + 255) / 256);
        const uint8_t* p = (const uint8_t*)img->data;
        auto* ds = d.as<unsigned long long>();
        switch (pixfmt) {
            case ZB_PIX_U8: moments_kernel<1><<<blocks, 256, 0, s>>>(p, n_px, as_luma, ds); break;
            case ZB_PIX_RGB8: moments_kernel<3><<<blocks, 256, 0, s>>>(p, n_px, as_luma, ds); break;
            default: moments_kernel<4><<<blocks, 256, 0, s>>>(p, n_px, as_luma, ds); break;
        }
        ZB_LAUNCHED();
    }
    ZB_CUDA(cudaMemcpyAsync(sums11, d.p, 11 * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    ZB_CUDA(cudaStreamSynchronize(s));
    return ZB_OK;
}

// mean (stats.zig:283-286) and unbiased covariance (:301-320) of x/255 from the integer moments
void stats_from_moments(const uint64_t* m, double mean[3], double cov[9]) {
    const double n = (double)m[0];
    for (int i = 0; i < 3; ++i) mean[i] = m[0] ? ((double)m[1 + i] / n) / 255.0 : 0.0;
    for (int i = 0; i < 9; ++i) cov[i] = 0.0;
    if (m[0] <= 1) return;
    static const int idx[3][3] = {{4, 5, 6}, {5, 7, 8}, {6, 8, 9}};
    const long double denom = (long double)m[0] * (long double)(m[0] - 1) * 65025.0L;
    for (int i = 0; i < 3; ++i)
        for (int j = i; j < 3; ++j) {
            const __int128 num = (__int128)m[0] * (__int128)m[idx[i][j]] - (__int128)m[1 + i] * (__int128)m[1 + j];
            const double c = (double)((long double)num / denom);
            cov[i * 3 + j] = c;
            cov[j * 3 + i] = c;
        }
}

// fdm.zig:92-121
int set_target_from_moments(zb_fdm* f, const uint64_t* m) {
    double cov[9];
    stats_from_moments(m, f->target_mean, cov);
    f->target_is_gray = (f->pixfmt == ZB_PIX_U8) || (m[10] == 0);
    memset(f->target_u, 0, sizeof(f->target_u));
    f->target_s[0] = f->target_s[1] = f->target_s[2] = 0;
    if (f->target_is_gray) {
        f->target_s[0] = cov[0];
    } else {
        double v[9];
        if (svd_golub_reinsch<double>(cov, 3, 3, ZB_SVD_SKINNY_U, false, f->target_u, 3, f->target_s, v) != 0) return ZB_ERR_NOT_CONVERGED;
    }
    f->has_target = true;
    return ZB_OK;
}

// scalar 3x3 product in the reference's gemm order (Matrix.zig:806-817: 27 ops < 512 -> scalar path)
void matmul3(const double* a, const double* b, double* out) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double acc = 0;
            for (int k = 0; k < 3; ++k) acc += a[i * 3 + k] * b[k * 3 + j];
            out[i * 3 + j] = 0.0 + 1.0 * acc;
        }
}

// fdm.zig:174-272 given the source moments
int update_from_moments(zb_fdm* f, const uint64_t* m, cudaStream_t s) {
    double source_mean[3], scov[9];
    stats_from_moments(m, source_mean, scov);
    MapParams p;
    memset(&p, 0, sizeof(p));
    const int ch = channels_of(f->pixfmt);
    if (f->pixfmt == ZB_PIX_U8 || f->target_is_gray) {  // :177-198
        const double source_var = scov[0];
        p.scale = source_var > 1e-10 ? std::sqrt(f->target_s[0] / source_var) : 1.0;
        p.offset = f->target_mean[0] - source_mean[0] * p.scale;
        p.mode = f->pixfmt == ZB_PIX_U8 ? 1 : 2;
    } else {  // :199-254
        double us[9], ss[3], v[9];
        if (svd_golub_reinsch<double>(scov, 3, 3, ZB_SVD_SKINNY_U, false, us, 3, ss, v) != 0) return ZB_ERR_NOT_CONVERGED;
        double sigma[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 3; ++i)
            if (ss[i] > 1e-10) sigma[i * 3 + i] = std::sqrt(f->target_s[i] / ss[i]);
        double ut_t[9], w_temp[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) ut_t[i * 3 + j] = f->target_u[j * 3 + i];
        matmul3(us, sigma, w_temp);
        matmul3(w_temp, ut_t, p.w);
        for (int j = 0; j < 3; ++j) {
            double sum = 0;
            for (int k = 0; k < 3; ++k) sum += source_mean[k] * p.w[k * 3 + j];
            p.bias[j] = f->target_mean[j] - sum;
        }
        p.mode = 0;
    }
    const size_t n_px = (size_t)f->source.rows * f->source.cols;
    if (n_px == 0) return ZB_OK;
    uint8_t* img = (uint8_t*)f->source.data;
    const unsigned blocks = div_up(n_px, 256);
    switch (ch) {
        case 1: fdm_map_kernel<1><<<blocks, 256, 0, s>>>(img, n_px, p); break;
        case 3: fdm_map_kernel<3><<<blocks, 256, 0, s>>>(img, n_px, p); break;
        default: fdm_map_kernel<4><<<blocks, 256, 0, s>>>(img, n_px, p); break;
    }
    ZB_LAUNCHED();
    t_last_kernel = "fdm_map";
    return ZB_OK;
}

}  // namespace
}  // namespace zb

using namespace zb;

extern "C" {

int zb_fdm_create(zb_fdm** out, int pixfmt) {
    if (!out) return ZB_ERR_INVALID_ARGUMENT;
    if (pixfmt != ZB_PIX_U8 && pixfmt != ZB_PIX_RGB8 && pixfmt != ZB_PIX_RGBA8) return ZB_ERR_UNSUPPORTED;  // fdm.zig:20
    zb_fdm* f = new zb_fdm();
    memset(f, 0, sizeof(*f));
    f->pixfmt = pixfmt;
    *out = f;
    return ZB_OK;
}
int zb_fdm_destroy(zb_fdm* f) { delete f; return ZB_OK; }

int zb_fdm_moments(const zb_image* img, int pixfmt, int as_luma, uint64_t* sums11, zb_stream s) {
    return moments_device(img, pixfmt, as_luma, sums11, (cudaStream_t)s);
}

int zb_fdm_set_target_moments(zb_fdm* f, const uint64_t* sums11) {
    if (!f || !sums11) return ZB_ERR_INVALID_ARGUMENT;
    return set_target_from_moments(f, sums11);
}

int zb_fdm_set_target(zb_fdm* f, const zb_image* target, zb_stream s) {
    if (!f || !target) return ZB_ERR_INVALID_ARGUMENT;
    uint64_t m[11];
    int rc = moments_device(target, f->pixfmt, 0, m, (cudaStream_t)s);
    if (rc) return rc;
    return set_target_from_moments(f, m);
}

int zb_fdm_set_source(zb_fdm* f, zb_image* source) {
    if (!f || !source) return ZB_ERR_INVALID_ARGUMENT;
    if (source->stride != source->cols) return ZB_ERR_UNSUPPORTED;
    f->source = *source;
    f->has_source = true;
    return ZB_OK;
}

int zb_fdm_update_with_moments(zb_fdm* f, const uint64_t* source_sums11, zb_stream s) {
    if (!f || !source_sums11) return ZB_ERR_INVALID_ARGUMENT;
    if (!f->has_target) return ZB_ERR_NO_TARGET_SET;  // fdm.zig:142
    if (!f->has_source) return ZB_ERR_NO_SOURCE_SET;  // fdm.zig:143
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    return update_from_moments(f, source_sums11, (cudaStream_t)s);
}

int zb_fdm_update(zb_fdm* f, zb_stream s) {
    if (!f) return ZB_ERR_INVALID_ARGUMENT;
    if (!f->has_target) return ZB_ERR_NO_TARGET_SET;
    if (!f->has_source) return ZB_ERR_NO_SOURCE_SET;
    uint64_t m[11];
    const int as_luma = (f->pixfmt != ZB_PIX_U8 && f->target_is_gray) ? 1 : 0;  // fdm.zig:155-162
    int rc = moments_device(&f->source, f->pixfmt, as_luma, m, (cudaStream_t)s);
    if (rc) return rc;
    return update_from_moments(f, m, (cudaStream_t)s);
}

int zb_fdm_match(zb_fdm* f, zb_image* source, const zb_image* target, zb_stream s) {  // fdm.zig:133-137
    int rc;
    if ((rc = zb_fdm_set_target(f, target, s))) return rc;
    if ((rc = zb_fdm_set_source(f, source))) return rc;
    return zb_fdm_update(f, s);
}

int zb_host_fdm_match(zb_image* source, const zb_image* target, int pixfmt) {
    if (!source || !target) return ZB_ERR_INVALID_ARGUMENT;
    if (source->stride != source->cols || target->stride != target->cols) return ZB_ERR_UNSUPPORTED;
    zb_fdm* f = nullptr;
    int rc = zb_fdm_create(&f, pixfmt);
    if (rc) return rc;
    HostStage st;  // src slot = target image, dst slot = source image (modified in place)
    const size_t pb = pixel_bytes(pixfmt);
    rc = st.begin(target, source, pixfmt);
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(st.ddst.data, source->data, (size_t)source->rows * source->cols * pb, cudaMemcpyHostToDevice, st.stream);
        if (e != cudaSuccess) rc = set_cuda_error(e, "upload source", __FILE__, __LINE__);
    }
    if (!rc) rc = zb_fdm_match(f, &st.ddst, &st.dsrc, (zb_stream)st.stream);
    if (!rc) rc = st.finish(source, pixfmt);
    zb_fdm_destroy(f);
    return rc;
}

}  // extern "C"
