// zignal.hpp -- C++ host-side mirror of zignal's Image(T) hot methods over the C ABI (header only).
// The reference's toolchain (Zig nightly) is absent from this image, so the host layer above the C ABI
// is provided in C++ (this file) next to the uncompiled Zig shim (zig/zignal_b200.zig).  Method names,
// argument meaning and error behaviour follow reference src/image.zig:523-1147.  Image<T> wraps HOST buffers (each call is the literal
// drop-in: upload, kernel, download); DeviceImage<T> wraps device buffers for resident chains.  tests/test_host_cpp_mirror.py compiles this
// header with -Wall -Wextra -Werror for every pixel type and links it against the library.
#pragma once
#include <cmath>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/zignal_b200.h"

namespace zignal {

struct Rgb8 { uint8_t r, g, b; };
struct Rgba8 { uint8_t r, g, b, a; };
struct RgbaF32 { float r, g, b, a; };

template <typename T> constexpr int pixfmt_of();
template <> constexpr int pixfmt_of<uint8_t>() { return ZB_PIX_U8; }
template <> constexpr int pixfmt_of<float>() { return ZB_PIX_F32; }
template <> constexpr int pixfmt_of<Rgb8>() { return ZB_PIX_RGB8; }
template <> constexpr int pixfmt_of<Rgba8>() { return ZB_PIX_RGBA8; }
template <> constexpr int pixfmt_of<RgbaF32>() { return ZB_PIX_RGBAF32; }

struct Error : std::runtime_error {
    int status;
    explicit Error(int s) : std::runtime_error(std::string("zignal error.") + zb_status_name(s)), status(s) {}
};
inline void check(int status) { if (status != ZB_OK) throw Error(status); }

enum class BorderMode { zero = 0, replicate = 1, mirror = 2, wrap = 3 };
struct Interpolation {
    int tag = ZB_INTERP_BILINEAR;
    float b = 1.0f / 3.0f, c = 1.0f / 3.0f;
    static Interpolation nearest() { return {ZB_INTERP_NEAREST}; }
    static Interpolation bilinear() { return {ZB_INTERP_BILINEAR}; }
    static Interpolation bicubic() { return {ZB_INTERP_BICUBIC}; }
    static Interpolation catmull_rom() { return {ZB_INTERP_CATMULL_ROM}; }
    static Interpolation mitchell(float b, float c) { return {ZB_INTERP_MITCHELL, b, c}; }
    static Interpolation lanczos() { return {ZB_INTERP_LANCZOS}; }
};

// Image<T> over HOST memory: every method is the literal drop-in (H2D + kernel + D2H inside the call).
template <typename T>
struct Image {
    uint32_t rows = 0, cols = 0;
    T* data = nullptr;
    size_t stride = 0;  // pixels

    zb_image raw() const { return zb_image{(void*)data, rows, cols, (uint64_t)stride}; }
    bool hasSameShape(const Image& o) const { return rows == o.rows && cols == o.cols; }

    void gaussianBlur(Image out, float sigma) const { auto a = raw(), d = out.raw(); check(zb_host_gaussian_blur(&a, &d, pixfmt_of<T>(), sigma)); }
    void convolveSeparable(Image out, const std::vector<float>& kx, const std::vector<float>& ky, BorderMode border) const {
        auto a = raw(), d = out.raw();
        check(zb_host_conv_separable(&a, &d, pixfmt_of<T>(), kx.data(), (int)kx.size(), ky.data(), (int)ky.size(), (int)border));
    }
    template <size_t KH, size_t KW>
    void convolve(Image out, const float (&kernel)[KH][KW], BorderMode border) const {
        auto a = raw(), d = out.raw();
        check(zb_host_convolve(&a, &d, pixfmt_of<T>(), &kernel[0][0], (int)KH, (int)KW, (int)border));
    }
    void boxBlur(Image out, uint32_t radius) const { auto a = raw(), d = out.raw(); check(zb_host_box_blur(&a, &d, pixfmt_of<T>(), radius)); }
    void sharpen(Image out, size_t radius) const { auto a = raw(), d = out.raw(); check(zb_host_sharpen(&a, &d, pixfmt_of<T>(), (uint32_t)radius)); }
    void resize(Image out, Interpolation m) const { auto a = raw(), d = out.raw(); zb_host_resize(&a, &d, pixfmt_of<T>(), m.tag, m.b, m.c); }
    void rotateInto(Image out, float angle, Interpolation m, BorderMode border) const {
        auto a = raw(), d = out.raw();
        zb_host_rotate_into(&a, &d, pixfmt_of<T>(), angle, m.tag, m.b, m.c, (int)border);
    }
    struct Bounds { uint32_t rows, cols; };
    Bounds rotateBounds(float angle) const { Bounds b{}; zb_rotate_bounds(rows, cols, angle, &b.rows, &b.cols); return b; }
    // transform: {m00,m01,m10,m11,b0,b1} (similarity/affine) or 9 values (projective)
    void warp(Image out, int xform_kind, const float* m, Interpolation method) const {
        auto a = raw(), d = out.raw();
        zb_host_warp(&a, &d, pixfmt_of<T>(), xform_kind, m, method.tag, method.b, method.c);
    }
};

enum class Blending { none = 0, normal, multiply, screen, overlay, soft_light, hard_light, color_dodge, color_burn, darken, lighten, difference, exclusion };
struct Rect { float l, t, r, b; };   // geometry/Rectangle.zig (f32; r and b exclusive)

// DeviceImage<T>: the same {rows, cols, data, stride} struct over DEVICE memory, for callers that keep a chain of operations resident
// (SURVEY 8(f)).  Non-owning, like the reference's views; every method enqueues on `stream` and mirrors the Zig method it is named after
// (image.zig line in the comment).  Methods returning a double wait for the stream.
template <typename T>
struct DeviceImage {
    uint32_t rows = 0, cols = 0;
    T* data = nullptr;     // device pointer (zb_malloc / cudaMalloc)
    size_t stride = 0;     // pixels
    zb_stream stream = nullptr;

    zb_image raw() const { return zb_image{(void*)data, rows, cols, (uint64_t)stride}; }
    DeviceImage view(uint32_t l, uint32_t t, uint32_t r, uint32_t b) const {                       // :426-430
        return DeviceImage{b - t, r - l, data + (size_t)t * stride + l, stride, stream};
    }
    void copy(DeviceImage out) const { auto a = raw(), d = out.raw(); check(zb_copy(&a, &d, pixfmt_of<T>(), stream)); }              // :375
    template <typename U>
    void convertInto(DeviceImage<U> out) const { auto a = raw(), d = out.raw(); check(zb_convert(&a, pixfmt_of<T>(), &d, pixfmt_of<U>(), stream)); }   // :396
    void gaussianBlur(DeviceImage out, float sigma) const { auto a = raw(), d = out.raw(); check(zb_gaussian_blur(&a, &d, pixfmt_of<T>(), sigma, stream)); }   // :954
    void convolveSeparable(DeviceImage out, const std::vector<float>& kx, const std::vector<float>& ky, BorderMode border) const {     // :935
        auto a = raw(), d = out.raw();
        check(zb_conv_separable(&a, &d, pixfmt_of<T>(), kx.data(), (int)kx.size(), ky.data(), (int)ky.size(), (int)border, stream));
    }
    template <size_t KH, size_t KW>
    void convolve(DeviceImage out, const float (&kernel)[KH][KW], BorderMode border) const {                                          // :917
        auto a = raw(), d = out.raw();
        check(zb_convolve(&a, &d, pixfmt_of<T>(), &kernel[0][0], (int)KH, (int)KW, (int)border, stream));
    }
    void boxBlur(DeviceImage out, uint32_t radius) const { auto a = raw(), d = out.raw(); check(zb_box_blur(&a, &d, pixfmt_of<T>(), radius, stream)); }      // :635
    void sharpen(DeviceImage out, uint32_t radius) const { auto a = raw(), d = out.raw(); check(zb_sharpen(&a, &d, pixfmt_of<T>(), radius, stream)); }       // :785
    void medianBlur(DeviceImage out, uint32_t radius) const { order(out, radius, ZB_ORDER_PERCENTILE, 0.5, BorderMode::mirror); }                             // :650
    void percentileBlur(DeviceImage out, uint32_t radius, double percentile, BorderMode border) const { order(out, radius, ZB_ORDER_PERCENTILE, percentile, border); }   // :672
    void minBlur(DeviceImage out, uint32_t radius, BorderMode border) const { order(out, radius, ZB_ORDER_PERCENTILE, 0.0, border); }                         // :696
    void maxBlur(DeviceImage out, uint32_t radius, BorderMode border) const { order(out, radius, ZB_ORDER_PERCENTILE, 1.0, border); }                         // :719
    void midpointBlur(DeviceImage out, uint32_t radius, BorderMode border) const { order(out, radius, ZB_ORDER_MIDPOINT, 0.0, border); }                      // :742
    void alphaTrimmedMeanBlur(DeviceImage out, uint32_t radius, double trim, BorderMode border) const { order(out, radius, ZB_ORDER_ALPHA_TRIMMED, trim, border); }   // :767
    void motionBlurLinear(DeviceImage out, float angle, uint32_t distance) const {                                                                           // motion_blur.zig:65
        auto a = raw(), d = out.raw();
        check(zb_motion_blur_linear(&a, &d, pixfmt_of<T>(), angle, std::cos(angle), std::sin(angle), distance, stream));
    }
    void motionBlurRadial(DeviceImage out, float center_x, float center_y, float strength, bool spin) const {                                                // motion_blur.zig:252
        auto a = raw(), d = out.raw();
        check(zb_motion_blur_radial(&a, &d, pixfmt_of<T>(), center_x, center_y, strength, spin ? 1 : 0, stream));
    }
    void resize(DeviceImage out, Interpolation m) const { auto a = raw(), d = out.raw(); check(zb_resize(&a, &d, pixfmt_of<T>(), m.tag, m.b, m.c, stream)); }   // :523
    void rotateInto(DeviceImage out, float angle, Interpolation m, BorderMode border) const {                                                               // :564
        auto a = raw(), d = out.raw();
        check(zb_rotate_into(&a, &d, pixfmt_of<T>(), angle, m.tag, m.b, m.c, (int)border, stream));
    }
    void warp(DeviceImage out, int xform_kind, const float* m, Interpolation method) const {                                                                // :621
        auto a = raw(), d = out.raw();
        check(zb_warp(&a, &d, pixfmt_of<T>(), xform_kind, m, method.tag, method.b, method.c, stream));
    }
    void extract(DeviceImage out, Rect rect, float angle, Interpolation m, BorderMode border) const {                                                       // :594
        auto a = raw(), d = out.raw();
        check(zb_extract(&a, &d, pixfmt_of<T>(), rect.l, rect.t, rect.r, rect.b, angle, std::cos(angle), std::sin(angle), m.tag, m.b, m.c, (int)border, stream));
    }
    void insert(DeviceImage source, Rect rect, float angle, Interpolation m, Blending blend) {                                                              // :604
        auto d = raw(), a = source.raw();
        check(zb_insert_blend(&d, &a, pixfmt_of<T>(), rect.l, rect.t, rect.r, rect.b, angle, std::cos(angle), std::sin(angle), m.tag, m.b, m.c, (int)blend, stream));
    }
    void sobel(DeviceImage<uint8_t> out) const { auto a = raw(), d = out.raw(); check(zb_sobel(&a, &d, pixfmt_of<T>(), stream)); }                            // :999
    void canny(DeviceImage<uint8_t> out, float sigma, float low, float high) const {                                                                        // :1041
        auto a = raw(), d = out.raw();
        check(zb_canny(&a, &d, pixfmt_of<T>(), sigma, low, high, stream));
    }
    double psnr(DeviceImage other) const { return metric(zb_psnr, other); }                                                                                  // :1105
    double ssim(DeviceImage other) const { return metric(zb_ssim, other); }                                                                                  // :1126
    double meanPixelError(DeviceImage other) const { return metric(zb_mean_pixel_error, other); }                                                            // :1145

private:
    void order(DeviceImage out, uint32_t radius, int mode, double param, BorderMode border) const {
        auto a = raw(), d = out.raw();
        check(zb_order_blur(&a, &d, pixfmt_of<T>(), radius, mode, param, (int)border, stream));
    }
    template <typename F>
    double metric(F fn, DeviceImage other) const {
        auto a = raw(), b = other.raw();
        double out = 0.0;
        check(fn(&a, &b, pixfmt_of<T>(), &out, stream));
        return out;
    }
};

// Matrix.eigh (matrix/eigen.zig:34): row-major n x n symmetric input -> eigenvalues ascending + eigenvectors as columns.
struct Eigh {
    std::vector<double> values, vectors;
};
inline Eigh eigh(const std::vector<double>& a, uint32_t n) {
    Eigh e{std::vector<double>(n), std::vector<double>((size_t)n * n)};
    check(zb_eigh_f64(a.data(), n, n, e.values.data(), e.vectors.data()));
    return e;
}

}  // namespace zignal
