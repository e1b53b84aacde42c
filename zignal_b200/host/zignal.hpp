// zignal.hpp -- C++ host-side mirror of zignal's Image(T) hot methods over the C ABI (header only).
// The reference's toolchain (Zig nightly) is absent from this image, so the host layer above the C ABI
// is provided in C++ (this file) next to the uncompiled Zig shim (zig/zignal_b200.zig).  Method names,
// argument meaning and error behaviour follow reference src/image.zig:523-994.
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

}  // namespace zignal
