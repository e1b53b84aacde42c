// zo_common.h -- shared helpers of the CPU oracle (TEST INFRASTRUCTURE ONLY, see zignal_oracle.h).
// Every helper cites the reference lines it restates (paths relative to /root/reference/src/).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <vector>

#include "zignal_oracle.h"

namespace zo {

extern int g_threads;

// image.zig:97-102, :426-430.  Typed view over zo_image.
template <typename T>
struct Img {
    T* data;
    uint32_t rows, cols;
    size_t stride;  // pixels
    Img() : data(nullptr), rows(0), cols(0), stride(0) {}
    Img(T* d, uint32_t r, uint32_t c, size_t s) : data(d), rows(r), cols(c), stride(s) {}
    explicit Img(const zo_image* im) : data((T*)im->data), rows(im->rows), cols(im->cols), stride((size_t)im->stride) {}
    inline T& at(size_t r, size_t c) const { return data[r * stride + c]; }
};

// An owning contiguous plane (what `Image(T).init(allocator, rows, cols)` gives, image.zig:124-132).
// Like the reference's allocator.alloc, the storage is NOT initialised (so a parallel first pass first-touches its own pages).
template <typename T>
struct Plane {
    std::unique_ptr<T[]> store;
    T* buf;
    uint32_t rows, cols;
    Plane(uint32_t r, uint32_t c) : store(new T[(size_t)r * c + 1]), buf(store.get()), rows(r), cols(c) {}
    Img<T> img() { return Img<T>(buf, rows, cols, cols); }
};

// Zig @mod for integers is floored modulo (border.zig:57,61).
static inline int64_t floor_mod(int64_t a, int64_t b) {
    int64_t m = a % b;
    if (m != 0 && ((m < 0) != (b < 0))) m += b;
    return m;
}

// border.zig:46-63 resolveIndex.  -1 stands for null.
static inline int64_t resolve_index(int64_t idx, int64_t length, int border) {
    if (idx >= 0 && idx < length) return idx;
    switch (border) {
        case ZO_BORDER_ZERO: return -1;
        case ZO_BORDER_REPLICATE:
            if (length == 0) return -1;
            return std::min(std::max(idx, (int64_t)0), length - 1);
        case ZO_BORDER_MIRROR: {
            if (length <= 0) return -1;
            if (length == 1) return 0;
            const int64_t period = 2 * (length - 1);
            const int64_t m = floor_mod(idx, period);
            const int64_t i = (m < 0) ? m + period : m;
            return (i >= length) ? period - i : i;
        }
        case ZO_BORDER_WRAP:
            if (length == 0) return -1;
            return floor_mod(idx, length);
    }
    return -1;
}

// Zig @min/@max on floats return the non-NaN operand; std.math.clamp = @max(lower, @min(val, upper)).
static inline double zig_min(double a, double b) { return std::isnan(a) ? b : (std::isnan(b) ? a : (a < b ? a : b)); }
static inline double zig_max(double a, double b) { return std::isnan(a) ? b : (std::isnan(b) ? a : (a > b ? a : b)); }

// meta.zig:121-128: clamp(IntT, float): @trunc(std.math.clamp(@round(as(f64, v)), min, max)).
template <typename IntT>
static inline IntT clamp_from_float(double v) {
    const double lo = std::numeric_limits<IntT>::is_signed ? (double)std::numeric_limits<IntT>::min() : 0.0;
    const double hi = (double)std::numeric_limits<IntT>::max();
    const double r = std::round(v);  // half away from zero, like Zig @round
    const double c = zig_max(lo, zig_min(r, hi));
    return (IntT)c;
}
static inline uint8_t clamp_u8(float v) { return clamp_from_float<uint8_t>((double)v); }

// meta.zig:115-120: clamp(IntT, integer) saturates.
template <typename IntT>
static inline IntT clamp_from_int(int64_t v) {
    const int64_t lo = (int64_t)std::numeric_limits<IntT>::min();
    const int64_t hi = (int64_t)std::numeric_limits<IntT>::max();
    return (IntT)std::min(std::max(v, lo), hi);
}

// convolution.zig:18-22 divClampU8.
static inline uint8_t div_clamp_u8(int64_t accum, int64_t scale) {
    const int64_t half = scale / 2;
    const int64_t rounded = (accum + (accum >= 0 ? half : -half)) / scale;  // C '/' == @divTrunc
    return clamp_from_int<uint8_t>(rounded);
}

// meta.zig:20-25 as(i32, f32): @intFromFloat(@round(x)).
static inline int32_t as_i32(float v) { return (int32_t)std::round(v); }

// channel_ops.zig:56-112 / :122-136: de-interleave / interleave u8 struct pixels.
static inline void split_channels_u8(int channels, const uint8_t* base,
                                     uint32_t rows, uint32_t cols, size_t stride_px,
                                     std::vector<std::vector<uint8_t>>& planes, bool* uniform, uint8_t* uniform_val) {
    planes.assign(channels, std::vector<uint8_t>((size_t)rows * cols));
    bool has_value[4] = {false, false, false, false};
    bool is_uniform[4] = {true, true, true, true};
    uint8_t uv[4] = {0, 0, 0, 0};
    size_t idx = 0;
    for (uint32_t r = 0; r < rows; ++r) {
        for (uint32_t c = 0; c < cols; ++c) {
            const uint8_t* px = base + (r * stride_px + c) * channels;
            for (int i = 0; i < channels; ++i) {
                const uint8_t value = px[i];
                planes[i][idx] = value;
                if (!has_value[i]) { uv[i] = value; has_value[i] = true; }
                else if (is_uniform[i] && value != uv[i]) is_uniform[i] = false;
            }
            ++idx;
        }
    }
    for (int i = 0; i < channels; ++i) {
        if (uniform) uniform[i] = has_value[i] && is_uniform[i];
        if (uniform_val) uniform_val[i] = uv[i];
    }
}

static inline void merge_channels_u8(const uint8_t* const* planes, int channels, uint8_t* base,
                                     uint32_t rows, uint32_t cols, size_t stride_px) {
    size_t idx = 0;
    for (uint32_t r = 0; r < rows; ++r)
        for (uint32_t c = 0; c < cols; ++c) {
            uint8_t* px = base + (r * stride_px + c) * channels;
            for (int i = 0; i < channels; ++i) px[i] = planes[i][idx];
            ++idx;
        }
}

static inline int channels_of(int pixfmt) {
    switch (pixfmt) {
        case ZO_PIX_U8: case ZO_PIX_F32: return 1;
        case ZO_PIX_RGB8: return 3;
        case ZO_PIX_RGBA8: case ZO_PIX_RGBAF32: return 4;
    }
    return 0;
}
static inline size_t pixel_bytes(int pixfmt) {
    switch (pixfmt) {
        case ZO_PIX_U8: return 1; case ZO_PIX_F32: return 4; case ZO_PIX_RGB8: return 3;
        case ZO_PIX_RGBA8: return 4; case ZO_PIX_RGBAF32: return 16;
    }
    return 0;
}

}  // namespace zo
