// zb_api_conv.cu -- C entry points for Image.convolveSeparable / convolve / gaussianBlur
// (reference image.zig:917-994, convolution.zig:198-438) and their host-pointer twins.
#include <cmath>
#include <vector>

#include "zb_conv.h"
#include "zb_host_stage.h"

using namespace zb;

namespace zb {

static int check_shapes(const zb_image* src, const zb_image* dst, int pixfmt) {
    if (!src || !dst) return ZB_ERR_INVALID_ARGUMENT;
    if (channels_of(pixfmt) == 0) return ZB_ERR_UNSUPPORTED;
    if (src->rows != dst->rows || src->cols != dst->cols) return ZB_ERR_DIMENSION_MISMATCH;  // image.zig:927,947,962
    return ZB_OK;
}

int conv_separable_dispatch(const zb_image* src, zb_image* dst, int pixfmt, const float* kx, int nx, const float* ky, int ny,
                            int border, cudaStream_t s) {
    int rc = check_shapes(src, dst, pixfmt);
    if (rc) return rc;
    if (nx <= 0 || ny <= 0 || !kx || !ky) return ZB_ERR_INVALID_ARGUMENT;
    if (border < ZB_BORDER_ZERO || border > ZB_BORDER_WRAP) return ZB_ERR_INVALID_ARGUMENT;
    if (src->rows == 0 || src->cols == 0) return ZB_OK;
    DeviceInfo di;
    if ((rc = device_info(&di))) return rc;
    if (pixfmt == ZB_PIX_RGBAF32 && !g_force_generic.load()) {
        rc = conv_separable_fused_rgbaf32(src, dst, kx, nx, ky, ny, border, g_exact_f32.load() != 0, s);
        if (rc != ZB_ERR_UNSUPPORTED) return rc;
    }
    if (pixfmt == ZB_PIX_RGBA8 && !g_force_generic.load()) {
        rc = conv_separable_fused_rgba8(src, dst, kx, nx, ky, ny, border, s);
        if (rc != ZB_ERR_UNSUPPORTED) return rc;
    }
    return conv_separable_generic(src, dst, pixfmt, kx, nx, ky, ny, border, s);
}

// image.zig:972-990
int gaussian_taps_host(float sigma, std::vector<float>& taps) {
    const size_t radius = (size_t)std::ceil(3.0f * sigma);
    const size_t n = 2 * radius + 1;
    taps.resize(n);
    float sum = 0;
    for (size_t i = 0; i < n; ++i) {
        const float x = (float)i - (float)radius;
        taps[i] = std::exp(-(x * x) / (2.0f * sigma * sigma));
        sum += taps[i];
    }
    for (size_t i = 0; i < n; ++i) taps[i] /= sum;
    return ZB_OK;
}

}  // namespace zb

extern "C" {

int zb_gaussian_taps(float sigma, float* taps, int cap, int* n) {
    if (!n) return ZB_ERR_INVALID_ARGUMENT;
    if (sigma == 0) { *n = 0; return ZB_OK; }
    if (!(sigma > 0)) return ZB_ERR_INVALID_SIGMA;
    std::vector<float> t;
    gaussian_taps_host(sigma, t);
    *n = (int)t.size();
    if (!taps || cap < (int)t.size()) return ZB_ERR_INVALID_ARGUMENT;
    memcpy(taps, t.data(), t.size() * sizeof(float));
    return ZB_OK;
}

int zb_conv_separable(const zb_image* src, zb_image* dst, int pixfmt, const float* kx, int nx, const float* ky, int ny, int border,
                      zb_stream s) {
    return conv_separable_dispatch(src, dst, pixfmt, kx, nx, ky, ny, border, (cudaStream_t)s);
}

int zb_convolve(const zb_image* src, zb_image* dst, int pixfmt, const float* kernel, int kh, int kw, int border, zb_stream s) {
    int rc = check_shapes(src, dst, pixfmt);
    if (rc) return rc;
    if (kh <= 0 || kw <= 0 || !kernel) return ZB_ERR_INVALID_ARGUMENT;
    if (border < ZB_BORDER_ZERO || border > ZB_BORDER_WRAP) return ZB_ERR_INVALID_ARGUMENT;
    if (src->rows == 0 || src->cols == 0) return ZB_OK;
    DeviceInfo di;
    if ((rc = device_info(&di))) return rc;
    return convolve_generic(src, dst, pixfmt, kernel, kh, kw, border, (cudaStream_t)s);
}

int zb_gaussian_blur(const zb_image* src, zb_image* dst, int pixfmt, float sigma, zb_stream s) {
    int rc = check_shapes(src, dst, pixfmt);
    if (rc) return rc;
    if (sigma == 0) return zb_copy(src, dst, pixfmt, s);   // image.zig:966
    if (!(sigma > 0)) return ZB_ERR_INVALID_SIGMA;         // image.zig:970
    std::vector<float> taps;
    gaussian_taps_host(sigma, taps);
    return conv_separable_dispatch(src, dst, pixfmt, taps.data(), (int)taps.size(), taps.data(), (int)taps.size(), ZB_BORDER_MIRROR,
                                   (cudaStream_t)s);     // image.zig:993
}

int zb_host_conv_separable(const zb_image* src, zb_image* dst, int pixfmt, const float* kx, int nx, const float* ky, int ny,
                           int border) {
    int rc = check_shapes(src, dst, pixfmt);
    if (rc) return rc;
    HostStage st;
    if ((rc = st.begin(src, dst, pixfmt))) return rc;
    if ((rc = conv_separable_dispatch(&st.dsrc, &st.ddst, pixfmt, kx, nx, ky, ny, border, st.stream))) return rc;
    return st.finish(dst, pixfmt);
}

int zb_host_convolve(const zb_image* src, zb_image* dst, int pixfmt, const float* kernel, int kh, int kw, int border) {
    int rc = check_shapes(src, dst, pixfmt);
    if (rc) return rc;
    HostStage st;
    if ((rc = st.begin(src, dst, pixfmt))) return rc;
    if ((rc = zb_convolve(&st.dsrc, &st.ddst, pixfmt, kernel, kh, kw, border, st.stream))) return rc;
    return st.finish(dst, pixfmt);
}

int zb_host_gaussian_blur(const zb_image* src, zb_image* dst, int pixfmt, float sigma) {
    int rc = check_shapes(src, dst, pixfmt);
    if (rc) return rc;
    if (sigma < 0) return ZB_ERR_INVALID_SIGMA;
    HostStage st;
    if ((rc = st.begin(src, dst, pixfmt))) return rc;
    if ((rc = zb_gaussian_blur(&st.dsrc, &st.ddst, pixfmt, sigma, st.stream))) return rc;
    return st.finish(dst, pixfmt);
}

}  // extern "C"
