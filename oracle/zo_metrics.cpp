// zo_metrics.cpp -- CPU oracle (TEST INFRASTRUCTURE ONLY, see zignal_oracle.h) for Image.psnr / ssim / meanPixelError
// (reference image.zig:1105-1147 -> image/metrics.zig:10-234).  f64 throughout, the reference's loop and summation order.
#include "zo_common.h"

namespace zo {
namespace {

inline int channels(int pixfmt) {
    switch (pixfmt) {
        case ZO_PIX_U8: case ZO_PIX_F32: return 1;
        case ZO_PIX_RGB8: return 3;
        case ZO_PIX_RGBA8: case ZO_PIX_RGBAF32: return 4;
    }
    return 0;
}
inline bool is_float(int pixfmt) { return pixfmt == ZO_PIX_F32 || pixfmt == ZO_PIX_RGBAF32; }
inline double component_max(int pixfmt) { return is_float(pixfmt) ? 1.0 : 255.0; }     // metrics.zig:176-185

inline double component(const zo_image* im, int pixfmt, size_t r, size_t c, int k) {
    const int ch = channels(pixfmt);
    const size_t i = (r * (size_t)im->stride + c) * ch + k;
    return is_float(pixfmt) ? (double)((const float*)im->data)[i] : (double)((const uint8_t*)im->data)[i];
}

// metrics.zig:187-222 getPixelScalar: scalars as f64; Rgb / Rgba of u8 through color.rgbLuma (color.zig:1021-1027) scaled back to
// 0..255; any other struct (Rgba(f32)) the mean of its fields.
inline double pixel_scalar(const zo_image* im, int pixfmt, size_t r, size_t c) {
    if (pixfmt == ZO_PIX_U8 || pixfmt == ZO_PIX_F32) return component(im, pixfmt, r, c, 0);
    if (pixfmt == ZO_PIX_RGB8 || pixfmt == ZO_PIX_RGBA8) {
        const double rf = component(im, pixfmt, r, c, 0) / 255.0, gf = component(im, pixfmt, r, c, 1) / 255.0,
                     bf = component(im, pixfmt, r, c, 2) / 255.0;
        return (0.2126 * rf + 0.7152 * gf + 0.0722 * bf) * 255.0;
    }
    double sum = 0.0;
    for (int k = 0; k < 4; ++k) sum += component(im, pixfmt, r, c, k);
    return sum / 4.0;
}

}  // namespace
}  // namespace zo

using namespace zo;

extern "C" int zo_psnr(const zo_image* a, const zo_image* b, int pixfmt, double* out) {       // metrics.zig:10-54
    if (a->rows != b->rows || a->cols != b->cols) return ZO_ERR_DIMENSION_MISMATCH;
    const int ch = channels(pixfmt);
    if (!ch) return ZO_ERR_UNSUPPORTED;
    double mse = 0.0;
    size_t count = 0;
    for (size_t r = 0; r < a->rows; ++r)
        for (size_t c = 0; c < a->cols; ++c)
            for (int k = 0; k < ch; ++k) {
                const double diff = component(a, pixfmt, r, c, k) - component(b, pixfmt, r, c, k);
                mse += diff * diff;
                ++count;
            }
    mse /= (double)count;
    if (mse == 0.0) { *out = std::numeric_limits<double>::infinity(); return ZO_OK; }
    *out = 20.0 * std::log10(component_max(pixfmt)) - 10.0 * std::log10(mse);
    return ZO_OK;
}

extern "C" int zo_mean_pixel_error(const zo_image* a, const zo_image* b, int pixfmt, double* out) {   // metrics.zig:116-165
    if (a->rows != b->rows || a->cols != b->cols) return ZO_ERR_DIMENSION_MISMATCH;
    const int ch = channels(pixfmt);
    if (!ch) return ZO_ERR_UNSUPPORTED;
    double total = 0.0;
    size_t count = 0;
    for (size_t r = 0; r < a->rows; ++r)
        for (size_t c = 0; c < a->cols; ++c)
            for (int k = 0; k < ch; ++k) {
                total += std::fabs(component(a, pixfmt, r, c, k) - component(b, pixfmt, r, c, k));
                ++count;
            }
    if (count == 0) { *out = 0.0; return ZO_OK; }
    *out = (total / (double)count) / component_max(pixfmt);
    return ZO_OK;
}

extern "C" void zo_ssim_window(double* w /* 121 */) {                                                  // metrics.zig:232-251
    const double sigma = 1.5;
    double sum = 0.0;
    for (int dy = 0; dy < 11; ++dy)
        for (int dx = 0; dx < 11; ++dx) {
            const double y = (double)dy - 5.0, x = (double)dx - 5.0;
            const double g = std::exp(-(x * x + y * y) / (2.0 * sigma * sigma));
            w[dy * 11 + dx] = g;
            sum += g;
        }
    for (int i = 0; i < 121; ++i) w[i] /= sum;
}

extern "C" int zo_ssim(const zo_image* a, const zo_image* b, int pixfmt, double* out) {                // metrics.zig:56-114
    if (a->rows != b->rows || a->cols != b->cols) return ZO_ERR_DIMENSION_MISMATCH;
    if (a->rows < 11 || a->cols < 11) return ZO_ERR_IMAGE_TOO_SMALL;
    if (!channels(pixfmt)) return ZO_ERR_UNSUPPORTED;
    const double l = component_max(pixfmt);
    const double c1 = (0.01 * l) * (0.01 * l), c2 = (0.03 * l) * (0.03 * l);
    double window[121];
    zo_ssim_window(window);
    double ssim_sum = 0.0, weight_sum = 0.0;
    for (size_t row = 5; row < a->rows - 5; ++row)
        for (size_t col = 5; col < a->cols - 5; ++col) {
            double mu_x = 0.0, mu_y = 0.0, mu_x_sq = 0.0, mu_y_sq = 0.0, mu_xy = 0.0;
            for (int dy = 0; dy < 11; ++dy)
                for (int dx = 0; dx < 11; ++dx) {
                    const double weight = window[dy * 11 + dx];
                    const double vx = pixel_scalar(a, pixfmt, row - 5 + dy, col - 5 + dx);
                    const double vy = pixel_scalar(b, pixfmt, row - 5 + dy, col - 5 + dx);
                    mu_x += weight * vx;
                    mu_y += weight * vy;
                    mu_x_sq += weight * vx * vx;
                    mu_y_sq += weight * vy * vy;
                    mu_xy += weight * vx * vy;
                }
            const double sigma_x_sq = std::max(0.0, mu_x_sq - mu_x * mu_x);
            const double sigma_y_sq = std::max(0.0, mu_y_sq - mu_y * mu_y);
            const double sigma_xy = mu_xy - mu_x * mu_y;
            const double numerator = (2.0 * mu_x * mu_y + c1) * (2.0 * sigma_xy + c2);
            const double denominator = (mu_x * mu_x + mu_y * mu_y + c1) * (sigma_x_sq + sigma_y_sq + c2);
            ssim_sum += numerator / denominator;
            weight_sum += 1.0;
        }
    *out = ssim_sum / weight_sum;
    return ZO_OK;
}
