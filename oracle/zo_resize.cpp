// zo_resize.cpp -- CPU oracle for Image.interpolate / Image.resize.
// TEST INFRASTRUCTURE ONLY (see zignal_oracle.h).  Restates image/channel_ops.zig:144-493 (integer
// plane resizers for Rgb/Rgba u8), image/interpolation.zig:72-84 (interpolate), :89-191 (resize),
// :194-214 (resizeGeneric), :222-300 (kernels, Lanczos LUT), :306-519 (samplers).
#include "zo_sample.h"

namespace zo {

// interpolation.zig:256-267: 1025-entry LUT, lanczosKernel(i / (1024/3), 3) evaluated in f32.
const float* lanczos3_lut() {
    static float lut[1025];
    static bool init = false;
    if (!init) {
        const float step = 1024.0f / 3.0f;
        for (int i = 0; i < 1025; ++i) {
            const float x = (float)i / step;
            float v;
            if (x == 0) v = 1;
            else if (std::fabs(x) >= 3.0f) v = 0;
            else {
                const float pi_x = 3.14159265358979323846f * x;  // std.math.pi coerced to f32
                const float pi_x_over_a = pi_x / 3.0f;
                v = (3.0f * std::sin(pi_x) * std::sin(pi_x_over_a)) / (pi_x * pi_x);
            }
            lut[i] = v;
        }
        init = true;
    }
    return lut;
}

// ---------------------------------------------------------------------------------------------
// channel_ops.zig plane resizers (contiguous u8 planes)
// ---------------------------------------------------------------------------------------------
static inline size_t mirror_idx(int64_t i, uint32_t n) { return (size_t)resolve_index(i, (int64_t)n, ZO_BORDER_MIRROR); }

// channel_ops.zig:144-190
static void resize_plane_bilinear_u8(const uint8_t* src, uint8_t* dst, uint32_t sr, uint32_t sc, uint32_t dr, uint32_t dc) {
    const int32_t s = 256;
    const float sf = 256.0f;
    const float x_ratio = (float)sc / (float)dc;
    const float y_ratio = (float)sr / (float)dr;
#pragma omp parallel for schedule(static) num_threads(g_threads) if (g_threads > 1)
    for (uint32_t r = 0; r < dr; ++r) {
        const float src_y_f = ((float)r + 0.5f) * y_ratio - 0.5f;
        const int64_t src_y_i = (int64_t)std::floor(src_y_f);
        const int32_t fy = (int32_t)std::trunc((src_y_f - std::floor(src_y_f)) * sf);
        const size_t y0 = mirror_idx(src_y_i, sr), y1 = mirror_idx(src_y_i + 1, sr);
        for (uint32_t c = 0; c < dc; ++c) {
            const float src_x_f = ((float)c + 0.5f) * x_ratio - 0.5f;
            const int64_t src_x_i = (int64_t)std::floor(src_x_f);
            const int32_t fx = (int32_t)std::trunc((src_x_f - std::floor(src_x_f)) * sf);
            const size_t x0 = mirror_idx(src_x_i, sc), x1 = mirror_idx(src_x_i + 1, sc);
            const int32_t tl = src[y0 * sc + x0], tr = src[y0 * sc + x1];
            const int32_t bl = src[y1 * sc + x0], br = src[y1 * sc + x1];
            const int32_t top = tl * (s - fx) + tr * fx;
            const int32_t bottom = bl * (s - fx) + br * fx;
            const int32_t result = (top * (s - fy) + bottom * fy) / (s * s);  // @divTrunc, no rounding
            dst[(size_t)r * dc + c] = clamp_from_int<uint8_t>(result);
        }
    }
}

// channel_ops.zig:193-214
static void resize_plane_nearest_u8(const uint8_t* src, uint8_t* dst, uint32_t sr, uint32_t sc, uint32_t dr, uint32_t dc) {
    const float x_ratio = (float)sc / (float)dc;
    const float y_ratio = (float)sr / (float)dr;
#pragma omp parallel for schedule(static) num_threads(g_threads) if (g_threads > 1)
    for (uint32_t r = 0; r < dr; ++r) {
        const float src_y_f = ((float)r + 0.5f) * y_ratio - 0.5f;
        const uint32_t src_y = std::min(sr - 1, (uint32_t)std::round(src_y_f));
        for (uint32_t c = 0; c < dc; ++c) {
            const float src_x_f = ((float)c + 0.5f) * x_ratio - 0.5f;
            const uint32_t src_x = std::min(sc - 1, (uint32_t)std::round(src_x_f));
            dst[(size_t)r * dc + c] = src[(size_t)src_y * sc + src_x];
        }
    }
}

// Q8 kernel evaluators.  channel_ops.zig:228-244 (bicubic, a=-1), :303-319 (Catmull-Rom), :378-395 (Mitchell)
static inline int32_t cubic_q8(int32_t t) {
    const int32_t SCALE = 256;
    const int32_t at = t < 0 ? -t : t;
    if (at <= SCALE) {
        const int32_t t2 = (at * at) / SCALE, t3 = (t2 * at) / SCALE;
        return SCALE - 2 * t2 + t3;
    } else if (at <= 2 * SCALE) {
        const int32_t t2 = (at * at) / SCALE, t3 = (t2 * at) / SCALE;
        return 4 * SCALE - 8 * at + 5 * t2 - t3;
    }
    return 0;
}
static inline int32_t catmull_q8(int32_t t) {
    const int32_t SCALE = 256;
    const int32_t at = t < 0 ? -t : t;
    if (at <= SCALE) {
        const int32_t t2 = (at * at) / SCALE, t3 = (t2 * at) / SCALE;
        return SCALE - (5 * t2) / 2 + (3 * t3) / 2;
    } else if (at <= 2 * SCALE) {
        const int32_t t2 = (at * at) / SCALE, t3 = (t2 * at) / SCALE;
        return 2 * SCALE - 4 * at + (5 * t2) / 2 - t3 / 2;
    }
    return 0;
}
static inline int32_t mitchell_q8(int32_t t) {
    const int64_t s = 256, s2 = s * s, s3 = s2 * s;
    const int64_t at = t < 0 ? -(int64_t)t : (int64_t)t;
    if (at < s) {
        const int64_t at2 = at * at, at3 = at2 * at;
        return (int32_t)((21 * at3 - 36 * at2 * s + 16 * s3) / (18 * s2));
    } else if (at < 2 * s) {
        const int64_t at2 = at * at, at3 = at2 * at;
        return (int32_t)((-7 * at3 + 36 * at2 * s - 60 * at * s2 + 32 * s3) / (18 * s2));
    }
    return 0;
}

// channel_ops.zig:217-289 / :292-365 / :368-435 share this 4x4 integer loop.
template <int32_t (*KERNEL)(int32_t)>
static void resize_plane_cubic_u8(const uint8_t* src, uint8_t* dst, uint32_t sr, uint32_t sc, uint32_t dr, uint32_t dc) {
    const int32_t SCALE = 256;
    const float x_ratio = (float)sc / (float)dc;
    const float y_ratio = (float)sr / (float)dr;
#pragma omp parallel for schedule(static) num_threads(g_threads) if (g_threads > 1)
    for (uint32_t r = 0; r < dr; ++r) {
        const float src_y_f = ((float)r + 0.5f) * y_ratio - 0.5f;
        const int64_t src_y = (int64_t)std::floor(src_y_f);
        const int32_t fy = (int32_t)std::trunc((src_y_f - std::floor(src_y_f)) * (float)SCALE);
        for (uint32_t c = 0; c < dc; ++c) {
            const float src_x_f = ((float)c + 0.5f) * x_ratio - 0.5f;
            const int64_t src_x = (int64_t)std::floor(src_x_f);
            const int32_t fx = (int32_t)std::trunc((src_x_f - std::floor(src_x_f)) * (float)SCALE);
            int32_t sum = 0, weight_sum = 0;
            for (int ky = 0; ky < 4; ++ky) {
                const size_t pixel_y = mirror_idx(src_y + ky - 1, sr);
                const int32_t wy = KERNEL(ky * SCALE - SCALE - fy);
                for (int kx = 0; kx < 4; ++kx) {
                    const size_t pixel_x = mirror_idx(src_x + kx - 1, sc);
                    const int32_t wx = KERNEL(kx * SCALE - SCALE - fx);
                    const int32_t w = (wx * wy) / SCALE;
                    sum += (int32_t)src[pixel_y * sc + pixel_x] * w;
                    weight_sum += w;
                }
            }
            const int32_t result = weight_sum != 0 ? sum / weight_sum : 0;
            dst[(size_t)r * dc + c] = clamp_from_int<uint8_t>(result);
        }
    }
}

// channel_ops.zig:438-493 (f32 weights, direct sin, 6x6)
static inline float lanczos_direct(float x) {
    if (x == 0) return 1.0f;
    const float a = 3.0f;
    if (std::fabs(x) >= a) return 0.0f;
    const float pi_x = 3.14159265358979323846f * x;
    return (a * std::sin(pi_x) * std::sin(pi_x / a)) / (pi_x * pi_x);
}
static void resize_plane_lanczos_u8(const uint8_t* src, uint8_t* dst, uint32_t sr, uint32_t sc, uint32_t dr, uint32_t dc) {
    const float x_ratio = (float)sc / (float)dc;
    const float y_ratio = (float)sr / (float)dr;
#pragma omp parallel for schedule(static) num_threads(g_threads) if (g_threads > 1)
    for (uint32_t r = 0; r < dr; ++r) {
        const float src_y_f = ((float)r + 0.5f) * y_ratio - 0.5f;
        const int64_t src_y = (int64_t)std::floor(src_y_f);
        const float fy = src_y_f - std::floor(src_y_f);
        for (uint32_t c = 0; c < dc; ++c) {
            const float src_x_f = ((float)c + 0.5f) * x_ratio - 0.5f;
            const int64_t src_x = (int64_t)std::floor(src_x_f);
            const float fx = src_x_f - std::floor(src_x_f);
            float sum = 0, weight_sum = 0;
            for (int ky = 0; ky < 6; ++ky) {
                const size_t pixel_y = mirror_idx(src_y + ky - 2, sr);
                const float wy = lanczos_direct((float)(ky - 2) - fy);
                for (int kx = 0; kx < 6; ++kx) {
                    const size_t pixel_x = mirror_idx(src_x + kx - 2, sc);
                    const float wx = lanczos_direct((float)(kx - 2) - fx);
                    const float w = wx * wy;
                    sum += (float)src[pixel_y * sc + pixel_x] * w;
                    weight_sum += w;
                }
            }
            const float result = weight_sum != 0 ? sum / weight_sum : 0;
            dst[(size_t)r * dc + c] = clamp_u8(result);
        }
    }
}

// interpolation.zig:194-214 resizeGeneric
template <typename PX>
static void resize_generic(const zo_image* src, zo_image* dst, int method, float mb, float mc) {
    using T = typename PX::T;
    Img<T> s(src), d(dst);
    const float scale_x = (float)s.cols / (float)d.cols;
    const float scale_y = (float)s.rows / (float)d.rows;
#pragma omp parallel for schedule(static) num_threads(g_threads) if (g_threads > 1)
    for (uint32_t r = 0; r < d.rows; ++r) {
        const float src_y = ((float)r + 0.5f) * scale_y - 0.5f;
        for (uint32_t c = 0; c < d.cols; ++c) {
            const float src_x = ((float)c + 0.5f) * scale_x - 0.5f;
            T val;
            if (!interpolate<PX>(s, src_x, src_y, method, mb, mc, ZO_BORDER_MIRROR, &val)) val = PX::zero();
            d.at(r, c) = val;
        }
    }
}

// interpolation.zig:89-191
static int resize(const zo_image* src, zo_image* dst, int pixfmt, int method, float mb, float mc) {
    if (method < ZO_INTERP_NEAREST || method > ZO_INTERP_LANCZOS) return ZO_ERR_INVALID_ARGUMENT;
    const size_t pb = pixel_bytes(pixfmt);
    if (pb == 0) return ZO_ERR_UNSUPPORTED;
    if (src->rows == dst->rows && src->cols == dst->cols) {  // :91-108
        if (src->data == dst->data) return ZO_OK;
        for (uint32_t r = 0; r < src->rows; ++r)
            std::memcpy((uint8_t*)dst->data + r * dst->stride * pb, (const uint8_t*)src->data + r * src->stride * pb,
                        (size_t)src->cols * pb);
        return ZO_OK;
    }
    if (dst->rows == 0 || dst->cols == 0) return ZO_OK;
    if (pixfmt == ZO_PIX_RGB8 || pixfmt == ZO_PIX_RGBA8) {  // meta.isRgb(T), :111-186
        if (src->rows == 0 || src->cols == 0) {
            // resolveIndex(...).? on an empty axis is a reference panic; the generic path zero-fills.
            for (uint32_t r = 0; r < dst->rows; ++r) std::memset((uint8_t*)dst->data + r * dst->stride * pb, 0, (size_t)dst->cols * pb);
            return ZO_OK;
        }
        const int ch = channels_of(pixfmt);
        std::vector<std::vector<uint8_t>> channels;
        split_channels_u8(ch, (const uint8_t*)src->data, src->rows, src->cols, (size_t)src->stride, channels, nullptr, nullptr);
        std::vector<std::vector<uint8_t>> outs(ch, std::vector<uint8_t>((size_t)dst->rows * dst->cols));
        for (int i = 0; i < ch; ++i) {
            const uint8_t* s = channels[i].data();
            uint8_t* d = outs[i].data();
            switch (method) {
                case ZO_INTERP_NEAREST: resize_plane_nearest_u8(s, d, src->rows, src->cols, dst->rows, dst->cols); break;
                case ZO_INTERP_BILINEAR: resize_plane_bilinear_u8(s, d, src->rows, src->cols, dst->rows, dst->cols); break;
                case ZO_INTERP_BICUBIC: resize_plane_cubic_u8<cubic_q8>(s, d, src->rows, src->cols, dst->rows, dst->cols); break;
                case ZO_INTERP_CATMULL_ROM: resize_plane_cubic_u8<catmull_q8>(s, d, src->rows, src->cols, dst->rows, dst->cols); break;
                case ZO_INTERP_MITCHELL: resize_plane_cubic_u8<mitchell_q8>(s, d, src->rows, src->cols, dst->rows, dst->cols); break;  // ignores (b,c)
                case ZO_INTERP_LANCZOS: resize_plane_lanczos_u8(s, d, src->rows, src->cols, dst->rows, dst->cols); break;
            }
        }
        const uint8_t* finals[4];
        for (int i = 0; i < ch; ++i) finals[i] = outs[i].data();
        merge_channels_u8(finals, ch, (uint8_t*)dst->data, dst->rows, dst->cols, (size_t)dst->stride);
        return ZO_OK;
    }
    switch (pixfmt) {
        case ZO_PIX_U8: resize_generic<PxU8>(src, dst, method, mb, mc); return ZO_OK;
        case ZO_PIX_F32: resize_generic<PxF32>(src, dst, method, mb, mc); return ZO_OK;
        case ZO_PIX_RGBAF32: resize_generic<PxRgbaF32>(src, dst, method, mb, mc); return ZO_OK;
    }
    return ZO_ERR_UNSUPPORTED;
}

}  // namespace zo

extern "C" {

int zo_interpolate(const zo_image* src, int pixfmt, float x, float y, int method, float mb, float mc, int border,
                   void* out_pixel) {
    using namespace zo;
    switch (pixfmt) {
        case ZO_PIX_U8: return interpolate<PxU8>(Img<uint8_t>(src), x, y, method, mb, mc, border, (uint8_t*)out_pixel) ? 1 : 0;
        case ZO_PIX_F32: return interpolate<PxF32>(Img<float>(src), x, y, method, mb, mc, border, (float*)out_pixel) ? 1 : 0;
        case ZO_PIX_RGB8: return interpolate<PxRgb8>(Img<Rgb8>(src), x, y, method, mb, mc, border, (Rgb8*)out_pixel) ? 1 : 0;
        case ZO_PIX_RGBA8: return interpolate<PxRgba8>(Img<Rgba8>(src), x, y, method, mb, mc, border, (Rgba8*)out_pixel) ? 1 : 0;
        case ZO_PIX_RGBAF32: return interpolate<PxRgbaF32>(Img<RgbaF32>(src), x, y, method, mb, mc, border, (RgbaF32*)out_pixel) ? 1 : 0;
    }
    return 0;
}

int zo_resize(const zo_image* src, zo_image* dst, int pixfmt, int method, float mb, float mc) {
    return zo::resize(src, dst, pixfmt, method, mb, mc);
}

}  // extern "C"
