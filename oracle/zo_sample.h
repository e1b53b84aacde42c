// zo_sample.h -- generic per-pixel samplers of the CPU oracle (TEST INFRASTRUCTURE ONLY).
// Restates image/interpolation.zig:72-84 (interpolate), :222-300 (kernels), :306-311 (nearest),
// :313-407 (bilinear), :409-423 (kernel wrappers), :426-519 (interpolateWithKernel).
#pragma once
#include "zo_common.h"

namespace zo {

struct Rgb8 { uint8_t v[3]; };
struct Rgba8 { uint8_t v[4]; };
struct RgbaF32 { float v[4]; };

// Pixel traits: storage type T, channel type CT, channel count N.
struct PxU8 { using T = uint8_t; using CT = uint8_t; static constexpr int N = 1;
    static inline CT get(const T& p, int) { return p; } static inline void set(T& p, int, CT c) { p = c; }
    static inline T zero() { return 0; } };
struct PxF32 { using T = float; using CT = float; static constexpr int N = 1;
    static inline CT get(const T& p, int) { return p; } static inline void set(T& p, int, CT c) { p = c; }
    static inline T zero() { return 0.0f; } };
struct PxRgb8 { using T = Rgb8; using CT = uint8_t; static constexpr int N = 3;
    static inline CT get(const T& p, int k) { return p.v[k]; } static inline void set(T& p, int k, CT c) { p.v[k] = c; }
    static inline T zero() { return T{{0, 0, 0}}; } };
struct PxRgba8 { using T = Rgba8; using CT = uint8_t; static constexpr int N = 4;
    static inline CT get(const T& p, int k) { return p.v[k]; } static inline void set(T& p, int k, CT c) { p.v[k] = c; }
    static inline T zero() { return T{{0, 0, 0, 0}}; } };
struct PxRgbaF32 { using T = RgbaF32; using CT = float; static constexpr int N = 4;
    static inline CT get(const T& p, int k) { return p.v[k]; } static inline void set(T& p, int k, CT c) { p.v[k] = c; }
    static inline T zero() { return T{{0, 0, 0, 0}}; } };

// meta.clamp(CT, f32): ints round-half-away + saturate, floats plain cast (meta.zig:110-135).
template <typename CT> static inline CT clamp_channel(float v);
template <> inline uint8_t clamp_channel<uint8_t>(float v) { return clamp_u8(v); }
template <> inline float clamp_channel<float>(float v) { return v; }

const float* lanczos3_lut();

// interpolation.zig:222-230
static inline float bicubic_kernel(float t) {
    const float at = std::fabs(t);
    if (at <= 1) return 1 - 2 * at * at + at * at * at;
    else if (at <= 2) return 4 - 8 * at + 5 * at * at - at * at * at;
    return 0;
}
// interpolation.zig:234-242
static inline float catmull_rom_kernel(float x) {
    const float ax = std::fabs(x);
    if (ax <= 1) return 1.5f * ax * ax * ax - 2.5f * ax * ax + 1;
    else if (ax <= 2) return -0.5f * ax * ax * ax + 2.5f * ax * ax - 4 * ax + 2;
    return 0;
}
// interpolation.zig:270-280
static inline float lanczos3_kernel_lut(float x) {
    const float ax = std::fabs(x);
    if (ax >= 3.0f) return 0;
    const float step = 1024.0f / 3.0f;
    const float pos = ax * step;
    const size_t idx = (size_t)std::trunc(pos);
    const float frac = pos - (float)idx;
    const float* lut = lanczos3_lut();
    return lut[idx] * (1.0f - frac) + lut[idx + 1] * frac;
}
// interpolation.zig:284-300
static inline float mitchell_kernel(float x, float m_b, float m_c) {
    const float ax = std::fabs(x);
    const float ax2 = ax * ax;
    const float ax3 = ax2 * ax;
    if (ax < 1) {
        return ((12 - 9 * m_b - 6 * m_c) * ax3 + (-18 + 12 * m_b + 6 * m_c) * ax2 + (6 - 2 * m_b)) / 6;
    } else if (ax < 2) {
        return ((-m_b - 6 * m_c) * ax3 + (6 * m_b + 30 * m_c) * ax2 + (-12 * m_b - 48 * m_c) * ax + (8 * m_b + 24 * m_c)) / 6;
    }
    return 0;
}

// interpolation.zig:306-311
template <typename PX>
static bool interpolate_nearest(const Img<typename PX::T>& img, float x, float y, int border, typename PX::T* out) {
    const int64_t col = resolve_index((int64_t)std::round(x), (int64_t)img.cols, border);
    if (col < 0) return false;
    const int64_t row = resolve_index((int64_t)std::round(y), (int64_t)img.rows, border);
    if (row < 0) return false;
    *out = img.at((size_t)row, (size_t)col);
    return true;
}

// interpolation.zig:353-368 lerpInt for <= 8-bit fields (Intermediate = i32)
static inline uint8_t lerp_int_u8(uint8_t tl, uint8_t tr, uint8_t bl, uint8_t br, int32_t fx, int32_t fy) {
    const int32_t scale = 256;
    const int32_t top_val = (int32_t)tl * (scale - fx) + (int32_t)tr * fx;
    const int32_t bottom_val = (int32_t)bl * (scale - fx) + (int32_t)br * fx;
    const int32_t result = (top_val * (scale - fy) + bottom_val * fy + (scale * scale / 2)) / (scale * scale);
    return clamp_from_int<uint8_t>(result);
}
// interpolation.zig:370-377 lerpFloat
static inline float lerp_float(float tl, float tr, float bl, float br, float lr, float tb) {
    return (1 - tb) * ((1 - lr) * tl + lr * tr) + tb * ((1 - lr) * bl + lr * br);
}

// interpolation.zig:313-407
template <typename PX>
static bool interpolate_bilinear(const Img<typename PX::T>& img, float x, float y, int border, typename PX::T* out) {
    using T = typename PX::T;
    const int64_t left = (int64_t)std::floor(x), top = (int64_t)std::floor(y);
    const int64_t right = left + 1, bottom = top + 1;
    const int64_t r0 = resolve_index(top, (int64_t)img.rows, border);
    const int64_t r1 = resolve_index(bottom, (int64_t)img.rows, border);
    const int64_t c0 = resolve_index(left, (int64_t)img.cols, border);
    const int64_t c1 = resolve_index(right, (int64_t)img.cols, border);
    auto get = [&](int64_t r, int64_t c) -> T { return (r >= 0 && c >= 0) ? img.at((size_t)r, (size_t)c) : PX::zero(); };
    if (border == ZO_BORDER_MIRROR) {  // :337-339
        if (r0 < 0 || r1 < 0 || c0 < 0 || c1 < 0) return false;
    }
    const T tl = get(r0, c0), tr = get(r0, c1), bl = get(r1, c0), br = get(r1, c1);
    const float lr_frac = x - (float)left;
    const float tb_frac = y - (float)top;
    const int32_t fx = (int32_t)std::round(lr_frac * 256.0f);
    const int32_t fy = (int32_t)std::round(tb_frac * 256.0f);
    T temp = PX::zero();
    for (int k = 0; k < PX::N; ++k) {
        if constexpr (std::is_same<typename PX::CT, uint8_t>::value) {
            PX::set(temp, k, lerp_int_u8(PX::get(tl, k), PX::get(tr, k), PX::get(bl, k), PX::get(br, k), fx, fy));
        } else {
            PX::set(temp, k, lerp_float(PX::get(tl, k), PX::get(tr, k), PX::get(bl, k), PX::get(br, k), lr_frac, tb_frac));
        }
    }
    *out = temp;
    return true;
}

// interpolation.zig:426-519.  KIND: 0 bicubic, 1 catmull-rom, 2 lanczos (LUT, radius 3), 3 mitchell(b,c)
template <typename PX, int KIND>
static bool interpolate_with_kernel(const Img<typename PX::T>& img, float x, float y, float mb, float mc, int border,
                                    typename PX::T* out) {
    using T = typename PX::T;
    constexpr int window_radius = (KIND == 2) ? 3 : 2;
    constexpr int window_size = window_radius * 2;
    const int64_t ix = (int64_t)std::floor(x), iy = (int64_t)std::floor(y);
    const float fx = x - (float)ix, fy = y - (float)iy;
    float xw[6], yw[6];
    auto kern = [&](float t) -> float {
        if (KIND == 0) return bicubic_kernel(t);
        if (KIND == 1) return catmull_rom_kernel(t);
        if (KIND == 2) return lanczos3_kernel_lut(t);
        return mitchell_kernel(t, mb, mc);
    };
    for (int i = 0; i < window_size; ++i) {
        const float off = (float)(i - (window_radius - 1));
        xw[i] = kern(off - fx);
        yw[i] = kern(off - fy);
    }
    float sums[4] = {0, 0, 0, 0};
    float weight_sum = 0;
    for (int j = 0; j < window_size; ++j) {
        const int64_t row_idx = iy - (window_radius - 1) + j;
        const int64_t py = resolve_index(row_idx, (int64_t)img.rows, border);
        if (py < 0) continue;
        for (int i = 0; i < window_size; ++i) {
            const int64_t col_idx = ix - (window_radius - 1) + i;
            const int64_t px = resolve_index(col_idx, (int64_t)img.cols, border);
            if (px < 0) continue;
            const T pixel = img.at((size_t)py, (size_t)px);
            const float weight = xw[i] * yw[j];
            for (int k = 0; k < PX::N; ++k) sums[k] += (float)PX::get(pixel, k) * weight;
            weight_sum += weight;
        }
    }
    T result = PX::zero();
    for (int k = 0; k < PX::N; ++k) {
        const float val = weight_sum != 0 ? sums[k] / weight_sum : 0;
        PX::set(result, k, clamp_channel<typename PX::CT>(val));
    }
    *out = result;
    return true;
}

// interpolation.zig:72-84
template <typename PX>
static bool interpolate(const Img<typename PX::T>& img, float x, float y, int method, float mb, float mc, int border,
                        typename PX::T* out) {
    if (!std::isfinite(x) || !std::isfinite(y)) return false;
    const float range_limit = (float)(INT64_MAX / 2);
    if (std::fabs(x) > range_limit || std::fabs(y) > range_limit) return false;
    switch (method) {
        case ZO_INTERP_NEAREST: return interpolate_nearest<PX>(img, x, y, border, out);
        case ZO_INTERP_BILINEAR: return interpolate_bilinear<PX>(img, x, y, border, out);
        case ZO_INTERP_BICUBIC: return interpolate_with_kernel<PX, 0>(img, x, y, mb, mc, border, out);
        case ZO_INTERP_CATMULL_ROM: return interpolate_with_kernel<PX, 1>(img, x, y, mb, mc, border, out);
        case ZO_INTERP_LANCZOS: return interpolate_with_kernel<PX, 2>(img, x, y, mb, mc, border, out);
        case ZO_INTERP_MITCHELL: return interpolate_with_kernel<PX, 3>(img, x, y, mb, mc, border, out);
    }
    return false;
}

}  // namespace zo
