// zb_sample.cuh -- device samplers: Image.interpolate for every Interpolation variant.
// Reference: interpolation.zig:72-84 (interpolate), :222-300 (kernels; Lanczos via a 1025-entry LUT
// with linear interpolation :256-280), :306-311 (nearest), :313-407 (bilinear: integer lerp with
// fx = round(frac*256) and +32768 rounding for <=16-bit ints, float lerp otherwise), :426-519
// (interpolateWithKernel: f32 weights wx[i]*wy[j], row-major accumulation, / weight_sum, meta.clamp;
// out-of-range taps are skipped, so kernel samplers renormalise at .zero borders).
//
// This translation unit family is compiled with -fmad=false: every f32 expression below is evaluated
// exactly as written (separately rounded mul/add), which is what the reference's Zig does.
#pragma once
#include <type_traits>

#include "../../include/zignal_b200.h"
#include "zb_device.cuh"

namespace zb {

template <typename CT, int N>
struct Pix {
    CT v[N];
};
// Rgba(u8): one 32-bit register; `u` is what loads, stores and the packed bilinear blend touch, `v` serves the per-channel code.
template <>
struct Pix<uint8_t, 4> {
    union {
        uint8_t v[4];
        uint32_t u;
    };
};

template <typename CT, int N>
__device__ __forceinline__ Pix<CT, N> load_px(const CT* __restrict__ base, size_t px) {
    Pix<CT, N> p;
    if constexpr (sizeof(CT) == 1 && N == 4) {
        p.u = *reinterpret_cast<const uint32_t*>(base + px * 4);
    } else if constexpr (sizeof(CT) == 4 && N == 4) {
        const float4 q = *reinterpret_cast<const float4*>(base + px * 4);
        p.v[0] = q.x; p.v[1] = q.y; p.v[2] = q.z; p.v[3] = q.w;
    } else {
#pragma unroll
        for (int k = 0; k < N; ++k) p.v[k] = base[px * N + k];
    }
    return p;
}

template <typename CT, int N>
__device__ __forceinline__ void store_px(CT* __restrict__ base, size_t px, const Pix<CT, N>& p) {
    if constexpr (sizeof(CT) == 1 && N == 4) {
        *reinterpret_cast<uint32_t*>(base + px * 4) = p.u;
    } else if constexpr (sizeof(CT) == 4 && N == 4) {
        *reinterpret_cast<float4*>(base + px * 4) = make_float4(p.v[0], p.v[1], p.v[2], p.v[3]);
    } else {
#pragma unroll
        for (int k = 0; k < N; ++k) base[px * N + k] = p.v[k];
    }
}

template <typename CT, int N>
__device__ __forceinline__ Pix<CT, N> zero_px() {
    Pix<CT, N> p;
    if constexpr (sizeof(CT) == 1 && N == 4) {
        p.u = 0u;
    } else {
#pragma unroll
        for (int k = 0; k < N; ++k) p.v[k] = (CT)0;
    }
    return p;
}

// @round(256 * (x - floor(x))) (interpolation.zig:349-352) without the float round: the fraction f is exact, 256 f >= 0, and
// round-half-away(256 f) = floor(256 f + 1/2) = (floor(512 f) + 1) >> 1 with floor(512 f) = floor(512 x) - 512 floor(x);
// 512 x is an exact product and fits an int32 for |x| < 2^21 (larger coordinates take the float path).
__device__ __forceinline__ unsigned frac_q8(float x, float floor_x, int floor_i) {
    if (fabsf(x) < 2097152.0f) return (unsigned)((__float2int_rd(x * 512.0f) - (floor_i << 9) + 1) >> 1);
    return (unsigned)(int)roundf((x - floor_x) * 256.0f);
}

// interpolation.zig:349-367 on four channels at once.  Horizontal blend on two 16-bit lanes per register
// (channel * 256 <= 65280 never carries into the neighbour lane), vertical blend per channel in 32 bits:
// (top * (256 - fy) + bottom * fy + 32768) < 2^24, so ">> 16" leaves the result (<= 255, no clamp needed) in byte 2.
__device__ __forceinline__ uint32_t bilerp_rgba8(uint32_t tl, uint32_t tr, uint32_t bl, uint32_t br, unsigned fx, unsigned fy) {
    const unsigned gx = 256u - fx, gy = 256u - fy;
    const uint32_t t02 = __byte_perm(tl, 0, 0x4240) * gx + __byte_perm(tr, 0, 0x4240) * fx;
    const uint32_t t13 = __byte_perm(tl, 0, 0x4341) * gx + __byte_perm(tr, 0, 0x4341) * fx;
    const uint32_t b02 = __byte_perm(bl, 0, 0x4240) * gx + __byte_perm(br, 0, 0x4240) * fx;
    const uint32_t b13 = __byte_perm(bl, 0, 0x4341) * gx + __byte_perm(br, 0, 0x4341) * fx;
    const uint32_t r0 = (t02 & 0xFFFFu) * gy + (b02 & 0xFFFFu) * fy + 32768u;
    const uint32_t r2 = (t02 >> 16) * gy + (b02 >> 16) * fy + 32768u;
    const uint32_t r1 = (t13 & 0xFFFFu) * gy + (b13 & 0xFFFFu) * fy + 32768u;
    const uint32_t r3 = (t13 >> 16) * gy + (b13 >> 16) * fy + 32768u;
    return __byte_perm(__byte_perm(r0, r1, 0x0062), __byte_perm(r2, r3, 0x0062), 0x5410);
}

struct SrcView {
    const void* data;
    int rows, cols;
    unsigned long long stride;  // pixels
};

// interpolation.zig:222-230
__device__ __forceinline__ float bicubic_kernel(float t) {
    const float at = fabsf(t);
    if (at <= 1) return 1 - 2 * at * at + at * at * at;
    else if (at <= 2) return 4 - 8 * at + 5 * at * at - at * at * at;
    return 0;
}
// interpolation.zig:234-242
__device__ __forceinline__ float catmull_rom_kernel(float x) {
    const float ax = fabsf(x);
    if (ax <= 1) return 1.5f * ax * ax * ax - 2.5f * ax * ax + 1;
    else if (ax <= 2) return -0.5f * ax * ax * ax + 2.5f * ax * ax - 4 * ax + 2;
    return 0;
}
// interpolation.zig:270-280 (lut: 1025 host-computed entries, :256-267)
__device__ __forceinline__ float lanczos3_kernel_lut(float x, const float* __restrict__ lut) {
    const float ax = fabsf(x);
    if (ax >= 3.0f) return 0;
    const float step = 1024.0f / 3.0f;
    const float pos = ax * step;
    const int idx = (int)truncf(pos);
    const float frac = pos - (float)idx;
    return lut[idx] * (1.0f - frac) + lut[idx + 1] * frac;
}
// interpolation.zig:284-300
__device__ __forceinline__ float mitchell_kernel(float x, float m_b, float m_c) {
    const float ax = fabsf(x);
    const float ax2 = ax * ax;
    const float ax3 = ax2 * ax;
    if (ax < 1) {
        return ((12 - 9 * m_b - 6 * m_c) * ax3 + (-18 + 12 * m_b + 6 * m_c) * ax2 + (6 - 2 * m_b)) / 6;
    } else if (ax < 2) {
        return ((-m_b - 6 * m_c) * ax3 + (6 * m_b + 30 * m_c) * ax2 + (-12 * m_b - 48 * m_c) * ax + (8 * m_b + 24 * m_c)) / 6;
    }
    return 0;
}

// meta.clamp(CT, f32): ints round half away + saturate, floats plain cast
template <typename CT>
__device__ __forceinline__ CT clamp_channel(float v) {
    if constexpr (sizeof(CT) == 1) return clamp_u8_from_float(v);
    else return v;
}

// interpolation.zig:353-368 lerpInt, 8-bit fields
__device__ __forceinline__ uint8_t lerp_int_u8(int tl, int tr, int bl, int br, int fx, int fy) {
    const int scale = 256;
    const int top_val = tl * (scale - fx) + tr * fx;
    const int bottom_val = bl * (scale - fx) + br * fx;
    const int result = (top_val * (scale - fy) + bottom_val * fy + (scale * scale / 2)) / (scale * scale);
    return (uint8_t)(result < 0 ? 0 : (result > 255 ? 255 : result));
}

// Index arithmetic is templated: `int` when the sample coordinates are far inside the i32 range (always,
// in practice), `long long` otherwise -- the reference uses isize (interpolation.zig:314-322).
template <typename I>
__device__ __forceinline__ I resolve_idx(I idx, I length, int border) {
    if (idx >= 0 && idx < length) return idx;
    if (border == 0) return -1;
    if (length <= 0) return -1;
    if (border == 1) return idx < 0 ? 0 : length - 1;
    if (border == 2) {
        if (length == 1) return 0;
        const I period = 2 * (length - 1);
        I m = idx % period;
        if (m < 0) m += period;
        return m >= length ? period - m : m;
    }
    I m = idx % length;
    if (m < 0) m += length;
    return m;
}

// BORDER_T >= 0 fixes the border mode at compile time (the hot combinations); -1 keeps it a runtime value.
template <typename CT, int N, int METHOD, typename I, int BORDER_T = -1>
__device__ __forceinline__ bool interpolate_impl(const SrcView& img, float x, float y, float mb, float mc, int border_rt,
                                                 const float* __restrict__ lut, Pix<CT, N>& out) {
    const CT* base = (const CT*)img.data;
    const I rows = (I)img.rows, cols = (I)img.cols;
    const int border = BORDER_T >= 0 ? BORDER_T : border_rt;

    if constexpr (METHOD == ZB_INTERP_NEAREST) {  // :306-311
        const I col = resolve_idx<I>((I)roundf(x), cols, border);
        if (col < 0) return false;
        const I row = resolve_idx<I>((I)roundf(y), rows, border);
        if (row < 0) return false;
        out = load_px<CT, N>(base, (size_t)row * img.stride + (size_t)col);
        return true;
    } else if constexpr (METHOD == ZB_INTERP_BILINEAR) {  // :313-407
        const float flx = floorf(x), fly = floorf(y);
        const I left = (I)flx, top = (I)fly;
        Pix<CT, N> tl, tr, bl, br;
        if (BORDER_T == ZB_BORDER_ZERO && (left < -1 || left >= cols || top < -1 || top >= rows)) {
            out = zero_px<CT, N>();   // all four neighbours are outside: .zero makes every one of them 0 (most of a rotated frame's margin)
            return true;
        }
        if (left >= 0 && left + 1 < cols && top >= 0 && top + 1 < rows) {  // all four neighbours inside: no border logic
            const CT* q = base + ((size_t)top * img.stride + (size_t)left) * N;
            tl = load_px<CT, N>(q, 0);
            tr = load_px<CT, N>(q, 1);
            bl = load_px<CT, N>(q, img.stride);
            br = load_px<CT, N>(q, img.stride + 1);
        } else {
            const I r0 = resolve_idx<I>(top, rows, border), r1 = resolve_idx<I>(top + 1, rows, border);
            const I c0 = resolve_idx<I>(left, cols, border), c1 = resolve_idx<I>(left + 1, cols, border);
            if (border == ZB_BORDER_MIRROR && (r0 < 0 || r1 < 0 || c0 < 0 || c1 < 0)) return false;  // :337-339
            const Pix<CT, N> z = zero_px<CT, N>();
            if (border == ZB_BORDER_ZERO && (r0 < 0 && r1 < 0 || c0 < 0 && c1 < 0)) { out = z; return true; }  // all four are zero
            tl = (r0 >= 0 && c0 >= 0) ? load_px<CT, N>(base, (size_t)r0 * img.stride + (size_t)c0) : z;
            tr = (r0 >= 0 && c1 >= 0) ? load_px<CT, N>(base, (size_t)r0 * img.stride + (size_t)c1) : z;
            bl = (r1 >= 0 && c0 >= 0) ? load_px<CT, N>(base, (size_t)r1 * img.stride + (size_t)c0) : z;
            br = (r1 >= 0 && c1 >= 0) ? load_px<CT, N>(base, (size_t)r1 * img.stride + (size_t)c1) : z;
        }
        const float lr = x - flx;  // == x - as(f32, left): floor(x) is exactly representable
        const float tb = y - fly;
        if constexpr (sizeof(CT) == 1 && N == 4) {
            const unsigned fx = frac_q8(x, flx, (int)left), fy = frac_q8(y, fly, (int)top);
            out.u = bilerp_rgba8(tl.u, tr.u, bl.u, br.u, fx, fy);
        } else if constexpr (sizeof(CT) == 1) {
            // :349-367.  fx, fy in [0, 256]; every intermediate is non-negative and < 2^25, so unsigned shift == @divTrunc
            const unsigned fx = frac_q8(x, flx, (int)left), fy = frac_q8(y, fly, (int)top);
#pragma unroll
            for (int k = 0; k < N; ++k) {
                const unsigned top_val = (unsigned)tl.v[k] * (256u - fx) + (unsigned)tr.v[k] * fx;
                const unsigned bottom_val = (unsigned)bl.v[k] * (256u - fx) + (unsigned)br.v[k] * fx;
                const unsigned result = (top_val * (256u - fy) + bottom_val * fy + 32768u) >> 16;
                out.v[k] = (uint8_t)(result > 255u ? 255u : result);
            }
        } else {
#pragma unroll
            for (int k = 0; k < N; ++k)
                out.v[k] = (1 - tb) * ((1 - lr) * tl.v[k] + lr * tr.v[k]) + tb * ((1 - lr) * bl.v[k] + lr * br.v[k]);
        }
        return true;
    } else {
        // kernel samplers, :426-519
        constexpr int window_radius = (METHOD == ZB_INTERP_LANCZOS) ? 3 : 2;
        constexpr int window_size = window_radius * 2;
        const float flx = floorf(x), fly = floorf(y);
        const I ix = (I)flx, iy = (I)fly;
        const float fx = x - flx, fy = y - fly;
        float xw[window_size], yw[window_size];
#pragma unroll
        for (int i = 0; i < window_size; ++i) {
            const float off = (float)(i - (window_radius - 1));
            const float ox = off - fx, oy = off - fy;
            if constexpr (METHOD == ZB_INTERP_BICUBIC) { xw[i] = bicubic_kernel(ox); yw[i] = bicubic_kernel(oy); }
            else if constexpr (METHOD == ZB_INTERP_CATMULL_ROM) { xw[i] = catmull_rom_kernel(ox); yw[i] = catmull_rom_kernel(oy); }
            else if constexpr (METHOD == ZB_INTERP_LANCZOS) { xw[i] = lanczos3_kernel_lut(ox, lut); yw[i] = lanczos3_kernel_lut(oy, lut); }
            else { xw[i] = mitchell_kernel(ox, mb, mc); yw[i] = mitchell_kernel(oy, mb, mc); }
        }
        float sums[N];
#pragma unroll
        for (int k = 0; k < N; ++k) sums[k] = 0;
        float weight_sum = 0;
#pragma unroll
        for (int j = 0; j < window_size; ++j) {
            const I py = resolve_idx<I>(iy - (window_radius - 1) + j, rows, border);
            if (py < 0) continue;
#pragma unroll
            for (int i = 0; i < window_size; ++i) {
                const I px = resolve_idx<I>(ix - (window_radius - 1) + i, cols, border);
                if (px < 0) continue;
                const Pix<CT, N> pixel = load_px<CT, N>(base, (size_t)py * img.stride + (size_t)px);
                const float weight = xw[i] * yw[j];
#pragma unroll
                for (int k = 0; k < N; ++k) sums[k] += (float)pixel.v[k] * weight;
                weight_sum += weight;
            }
        }
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const float val = weight_sum != 0 ? sums[k] / weight_sum : 0.0f;
            out.v[k] = clamp_channel<CT>(val);
        }
        return true;
    }
}

// interpolation.zig:72-84.  Returns false for null (caller writes zeroes).
template <typename CT, int N, int METHOD, int BORDER_T = -1>
__device__ __forceinline__ bool interpolate(const SrcView& img, float x, float y, float mb, float mc, int border,
                                            const float* __restrict__ lut, Pix<CT, N>& out) {
    if (fabsf(x) < 1.0e9f && fabsf(y) < 1.0e9f)  // finite and far inside the i32 range: 32-bit index math
        return interpolate_impl<CT, N, METHOD, int, BORDER_T>(img, x, y, mb, mc, border, lut, out);
    if (!isfinite(x) || !isfinite(y)) return false;
    const float range_limit = 4611686018427387904.0f;  // @floatFromInt(maxInt(isize) / 2)
    if (fabsf(x) > range_limit || fabsf(y) > range_limit) return false;
    return interpolate_impl<CT, N, METHOD, long long>(img, x, y, mb, mc, border, lut, out);
}

// Run `f(std::integral_constant<int, METHOD>)` for the runtime interpolation method.
template <typename F>
static inline int dispatch_method(int method, F&& f) {
    switch (method) {
        case ZB_INTERP_NEAREST: return f(std::integral_constant<int, ZB_INTERP_NEAREST>{});
        case ZB_INTERP_BILINEAR: return f(std::integral_constant<int, ZB_INTERP_BILINEAR>{});
        case ZB_INTERP_BICUBIC: return f(std::integral_constant<int, ZB_INTERP_BICUBIC>{});
        case ZB_INTERP_CATMULL_ROM: return f(std::integral_constant<int, ZB_INTERP_CATMULL_ROM>{});
        case ZB_INTERP_MITCHELL: return f(std::integral_constant<int, ZB_INTERP_MITCHELL>{});
        case ZB_INTERP_LANCZOS: return f(std::integral_constant<int, ZB_INTERP_LANCZOS>{});
    }
    return ZB_ERR_INVALID_ARGUMENT;
}

}  // namespace zb
