// zo_fdm.cpp -- CPU oracle for FeatureDistributionMatching and Pca.
// TEST INFRASTRUCTURE ONLY (see zignal_oracle.h).  Restates stats.zig:234-322 (CovarianceStats),
// fdm.zig:68-124 (setTarget), :141-273 (update), color.zig:1031-1041 (rgb -> gray, Q16 BT.709),
// pca.zig:104-164 (fit), :291-312 (transform), :331-362 (covariance path), :380-425 (Gram path).
#include "zo_linalg.h"

namespace zo {

// stats.zig:234-322 CovarianceStats(3, f64): streaming Welford mean + co-moment.
struct CovStats3 {
    size_t count = 0;
    double mean_vec[3] = {0, 0, 0};
    double m2[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    inline void add(const double sample[3]) {  // :261-280
        count += 1;
        const double n = (double)count;
        double delta[3];
        for (int i = 0; i < 3; ++i) {
            delta[i] = sample[i] - mean_vec[i];
            mean_vec[i] += delta[i] / n;
        }
        for (int i = 0; i < 3; ++i)
            for (int j = i; j < 3; ++j) {
                const double term = delta[i] * (sample[j] - mean_vec[j]);
                m2[i][j] += term;
                if (i != j) m2[j][i] += term;
            }
    }
    inline void mean(double out[3]) const {  // :283-286
        for (int i = 0; i < 3; ++i) out[i] = count == 0 ? 0 : mean_vec[i];
    }
    inline void variance_vector(double out[3]) const {  // :289-298
        if (count <= 1) { out[0] = out[1] = out[2] = 0; return; }
        const double n_1 = (double)(count - 1);
        for (int i = 0; i < 3; ++i) out[i] = m2[i][i] / n_1;
    }
    inline void covariance_matrix(double out[9]) const {  // :301-320
        if (count <= 1) { for (int i = 0; i < 9; ++i) out[i] = 0; return; }
        const double n_1 = (double)(count - 1);
        for (int i = 0; i < 3; ++i)
            for (int j = i; j < 3; ++j) {
                const double cv = m2[i][j] / n_1;
                out[i * 3 + j] = cv;
                if (i != j) out[j * 3 + i] = cv;
            }
    }
};

// color.zig:1031-1041 rgbToGray(u8): Q16 BT.709 luma, luma_r/g/b = 0.2126/0.7152/0.0722.
static inline uint8_t rgb_to_gray_u8(uint8_t r8, uint8_t g8, uint8_t b8) {
    const int32_t r = r8, g = g8, b = b8;
    const int32_t yr = 13933, yg = 46871, yb = 4732;  // @round(luma * 65536)
    int32_t y = (yr * r + yg * g + yb * b + 32768) >> 16;
    y = std::min(std::max(y, 0), 255);
    return (uint8_t)y;
}

// fdm.zig walks image.data linearly (:82,150,257) -- i.e. the first rows*cols elements of a contiguous
// image.  Views are not supported by the reference; the oracle requires stride == cols.
static bool contiguous(const zo_image* im) { return im->stride == im->cols; }

// Statistics pass of setTarget (:74-90) or update (:148-172).  as_luma selects the gray-target branch (:157-162).
static void fdm_stats(const zo_image* img, int pixfmt, bool as_luma, CovStats3& stats, bool* is_gray) {
    const size_t n = (size_t)img->rows * img->cols;
    const uint8_t* p = (const uint8_t*)img->data;
    bool gray = true;
    if (pixfmt == ZO_PIX_U8) {
        for (size_t i = 0; i < n; ++i) {
            const double v = (double)p[i] / 255.0;
            const double s[3] = {v, v, v};
            stats.add(s);
        }
    } else {
        const int ch = channels_of(pixfmt);
        for (size_t i = 0; i < n; ++i) {
            const uint8_t* px = p + i * ch;
            if (as_luma) {
                const double v = (double)rgb_to_gray_u8(px[0], px[1], px[2]) / 255.0;
                const double s[3] = {v, v, v};
                stats.add(s);
            } else {
                const double s[3] = {(double)px[0] / 255.0, (double)px[1] / 255.0, (double)px[2] / 255.0};
                stats.add(s);
            }
            if (px[0] != px[1] || px[1] != px[2]) gray = false;
        }
    }
    if (is_gray) *is_gray = (pixfmt == ZO_PIX_U8) ? true : gray;
}

static inline double clamp01(double v) { return zig_max(0.0, zig_min(v, 1.0)); }  // std.math.clamp(v, 0, 1)

// fdm.zig:257-271
static void fdm_apply_color(zo_image* source, int pixfmt, const double* w, const double* bias) {
    const size_t n = (size_t)source->rows * source->cols;
    const int ch = channels_of(pixfmt);
    uint8_t* p = (uint8_t*)source->data;
#pragma omp parallel for schedule(static) num_threads(g_threads) if (g_threads > 1)
    for (size_t i = 0; i < n; ++i) {
        uint8_t* px = p + i * ch;
        const double r = (double)px[0] / 255.0, g = (double)px[1] / 255.0, b = (double)px[2] / 255.0;
        const double res0 = r * w[0] + g * w[3] + b * w[6] + bias[0];
        const double res1 = r * w[1] + g * w[4] + b * w[7] + bias[1];
        const double res2 = r * w[2] + g * w[5] + b * w[8] + bias[2];
        px[0] = (uint8_t)std::round(255.0 * clamp01(res0));
        px[1] = (uint8_t)std::round(255.0 * clamp01(res1));
        px[2] = (uint8_t)std::round(255.0 * clamp01(res2));
    }
}

// fdm.zig:133-273 match() = setTarget + setSource + update
static int fdm_match(zo_image* source, const zo_image* target, int pixfmt) {
    if (pixfmt != ZO_PIX_U8 && pixfmt != ZO_PIX_RGB8 && pixfmt != ZO_PIX_RGBA8) return ZO_ERR_UNSUPPORTED;  // fdm.zig:20
    if (!contiguous(source) || !contiguous(target)) return ZO_ERR_UNSUPPORTED;
    // ---- setTarget (:68-124)
    CovStats3 tstats;
    bool target_is_gray = false;
    fdm_stats(target, pixfmt, false, tstats, &target_is_gray);
    double target_mean[3], target_cov_s[3] = {0, 0, 0}, target_u[9] = {0};
    tstats.mean(target_mean);
    if (target_is_gray) {
        double var[3];
        tstats.variance_vector(var);
        target_cov_s[0] = var[0];
    } else {
        double cov[9], v_unused[9];
        tstats.covariance_matrix(cov);
        if (svd<double>(cov, 3, 3, ZO_SVD_SKINNY_U, false, target_u, target_cov_s, v_unused) != 0) return ZO_ERR_NOT_CONVERGED;
    }
    // ---- update (:141-273)
    CovStats3 sstats;
    fdm_stats(source, pixfmt, pixfmt != ZO_PIX_U8 && target_is_gray, sstats, nullptr);
    double source_mean[3];
    sstats.mean(source_mean);
    const size_t n = (size_t)source->rows * source->cols;
    uint8_t* p = (uint8_t*)source->data;
    if (pixfmt == ZO_PIX_U8 || target_is_gray) {  // :177-198
        double svar[3];
        sstats.variance_vector(svar);
        const double source_var = svar[0];
        const double scale = source_var > 1e-10 ? std::sqrt(target_cov_s[0] / source_var) : 1.0;
        const double offset = target_mean[0] - source_mean[0] * scale;
        if (pixfmt == ZO_PIX_U8) {
            for (size_t i = 0; i < n; ++i) {
                const double val = (double)p[i] / 255.0;
                const double result = clamp01(val * scale + offset);
                p[i] = (uint8_t)std::round(255.0 * result);
            }
        } else {
            const int ch = channels_of(pixfmt);
            for (size_t i = 0; i < n; ++i) {
                uint8_t* px = p + i * ch;
                const double val = (double)rgb_to_gray_u8(px[0], px[1], px[2]) / 255.0;
                const double result = clamp01(val * scale + offset);
                const uint8_t res = (uint8_t)std::round(255.0 * result);
                px[0] = res; px[1] = res; px[2] = res;
                if (ch == 4) px[3] = 0;  // `.{ .r, .g, .b }` literal: alpha takes its default 0 (color.zig:405)
            }
        }
        return ZO_OK;
    }
    // colour branch (:199-272)
    double scov[9], us[9], ss[3], v_unused[9];
    sstats.covariance_matrix(scov);
    if (svd<double>(scov, 3, 3, ZO_SVD_SKINNY_U, false, us, ss, v_unused) != 0) return ZO_ERR_NOT_CONVERGED;
    double sigma[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 3; ++i)
        if (ss[i] > 1e-10) sigma[i * 3 + i] = std::sqrt(target_cov_s[i] / ss[i]);
    double ut_t[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) ut_t[i * 3 + j] = target_u[j * 3 + i];
    double w_temp[9], w[9];
    gemm<double>(us, 3, 3, false, sigma, 3, 3, false, 1.0, 0.0, nullptr, w_temp, 0);  // Matrix.dot, 27 ops < 512 -> scalar
    gemm<double>(w_temp, 3, 3, false, ut_t, 3, 3, false, 1.0, 0.0, nullptr, w, 0);
    double bias[3];
    for (int j = 0; j < 3; ++j) {
        double sum = 0;
        for (int k = 0; k < 3; ++k) sum += source_mean[k] * w[k * 3 + j];
        bias[j] = target_mean[j] - sum;
    }
    fdm_apply_color(source, pixfmt, w, bias);
    return ZO_OK;
}

// pca.zig:104-164 + :331-425
template <typename T>
static int pca_fit(const T* data, uint32_t n_samples, uint32_t dim, uint32_t k_req, T* mean, T* components, T* eigenvalues) {
    if (n_samples == 0) return ZO_ERR_INVALID_ARGUMENT;   // error.NoVectors
    if (n_samples == 1) return ZO_ERR_INVALID_ARGUMENT;   // error.InsufficientData
    if (k_req == 0) return ZO_ERR_INVALID_ARGUMENT;       // error.InvalidComponents
    const uint32_t max_components = std::min(n_samples - 1, dim);
    const uint32_t k = std::min(k_req, max_components);
    for (uint32_t j = 0; j < dim; ++j) mean[j] = 0;
    for (uint32_t i = 0; i < n_samples; ++i)
        for (uint32_t j = 0; j < dim; ++j) mean[j] += data[(size_t)i * dim + j];
    const T nf = (T)n_samples;
    for (uint32_t j = 0; j < dim; ++j) mean[j] /= nf;
    std::vector<T> centered((size_t)n_samples * dim);
    for (uint32_t i = 0; i < n_samples; ++i)
        for (uint32_t j = 0; j < dim; ++j) centered[(size_t)i * dim + j] = data[(size_t)i * dim + j] - mean[j];
    const T scale = (T)1.0 / (T)(n_samples - 1);
    if (n_samples <= dim) {  // Gram path :380-425
        const uint32_t n = n_samples;
        std::vector<T> gram((size_t)n * n), u((size_t)n * n), s(n), v(1);
        gemm<T>(centered.data(), n_samples, dim, false, centered.data(), n_samples, dim, true, scale, (T)0, nullptr, gram.data(), 0);
        if (svd<T>(gram.data(), n, n, ZO_SVD_SKINNY_U, false, u.data(), s.data(), v.data()) != 0) return ZO_ERR_NOT_CONVERGED;
        for (uint32_t i = 0; i < k; ++i) {
            const T eigenval = s[i];
            eigenvalues[i] = eigenval;
            if (eigenval > (T)1e-12) {
                for (uint32_t j = 0; j < dim; ++j) {
                    T sum = 0;
                    for (uint32_t kk = 0; kk < n; ++kk) sum += centered[(size_t)kk * dim + j] * u[(size_t)kk * n + i];
                    components[(size_t)j * k + i] = sum / std::sqrt(eigenval * (T)(n_samples - 1));
                }
            } else {
                for (uint32_t j = 0; j < dim; ++j) components[(size_t)j * k + i] = 0;
            }
        }
    } else {  // covariance path :331-362
        std::vector<T> cov((size_t)dim * dim), u((size_t)dim * dim), s(dim), v(1);
        gemm<T>(centered.data(), n_samples, dim, true, centered.data(), n_samples, dim, false, scale, (T)0, nullptr, cov.data(), 0);
        if (svd<T>(cov.data(), dim, dim, ZO_SVD_SKINNY_U, false, u.data(), s.data(), v.data()) != 0) return ZO_ERR_NOT_CONVERGED;
        for (uint32_t i = 0; i < k; ++i) {
            eigenvalues[i] = s[i];
            for (uint32_t j = 0; j < dim; ++j) components[(size_t)j * k + i] = u[(size_t)j * dim + i];
        }
    }
    return ZO_OK;
}

// pca.zig:291-312
template <typename T>
static int pca_transform(const T* data, uint32_t n, uint32_t dim, uint32_t k, const T* mean, const T* components, T* out) {
    if (n == 0) return ZO_ERR_INVALID_ARGUMENT;
    std::vector<T> centered((size_t)n * dim);
    for (uint32_t i = 0; i < n; ++i)
        for (uint32_t j = 0; j < dim; ++j) centered[(size_t)i * dim + j] = data[(size_t)i * dim + j] - mean[j];
    return gemm<T>(centered.data(), n, dim, false, components, dim, k, false, (T)1, (T)0, nullptr, out, 0);
}

}  // namespace zo

extern "C" {

int zo_fdm_stats(const zo_image* img, int pixfmt, int as_luma, double* mean3, double* cov9, int* is_gray) {
    if (!zo::contiguous(img)) return ZO_ERR_UNSUPPORTED;
    zo::CovStats3 st;
    bool gray = false;
    zo::fdm_stats(img, pixfmt, as_luma != 0, st, &gray);
    st.mean(mean3);
    st.covariance_matrix(cov9);
    if (is_gray) *is_gray = gray ? 1 : 0;
    return ZO_OK;
}
int zo_fdm_match(zo_image* source, const zo_image* target, int pixfmt) { return zo::fdm_match(source, target, pixfmt); }
int zo_fdm_apply(zo_image* source, int pixfmt, const double* w9, const double* bias3) {
    if (pixfmt != ZO_PIX_RGB8 && pixfmt != ZO_PIX_RGBA8) return ZO_ERR_UNSUPPORTED;
    if (!zo::contiguous(source)) return ZO_ERR_UNSUPPORTED;
    zo::fdm_apply_color(source, pixfmt, w9, bias3);
    return ZO_OK;
}
int zo_pca_fit_f64(const double* data, uint32_t n, uint32_t dim, uint32_t k, double* mean, double* comps, double* eig) {
    return zo::pca_fit<double>(data, n, dim, k, mean, comps, eig);
}
int zo_pca_fit_f32(const float* data, uint32_t n, uint32_t dim, uint32_t k, float* mean, float* comps, float* eig) {
    return zo::pca_fit<float>(data, n, dim, k, mean, comps, eig);
}
int zo_pca_transform_f64(const double* data, uint32_t n, uint32_t dim, uint32_t k, const double* mean, const double* comps, double* out) {
    return zo::pca_transform<double>(data, n, dim, k, mean, comps, out);
}
int zo_pca_transform_f32(const float* data, uint32_t n, uint32_t dim, uint32_t k, const float* mean, const float* comps, float* out) {
    return zo::pca_transform<float>(data, n, dim, k, mean, comps, out);
}

}  // extern "C"
