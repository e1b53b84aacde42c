// zo_transform.cpp -- CPU oracle for Image.rotate / rotateInto / rotateBounds / warp / extract.
// TEST INFRASTRUCTURE ONLY (see zignal_oracle.h).  Restates image/transforms.zig:112-149
// (rotateBounds), :163-212 (rotateInto), :385-462 (rotate0/90/180/270), :522-531 (warp),
// image.zig:200-229 (setBorder), :322-327 (getCenter), geometry/transforms.zig:39-42, :147-150,
// :224-231 (project) and matrix/SMatrix.zig:530-560 (the scalar gemm tail project() runs through).
#include "zo_sample.h"

namespace zo {

static const float TAU_F = 6.283185307179586f;   // std.math.tau coerced to f32
static const float PI_F = 3.141592653589793f;

// Zig's float @mod lowers to: a = frem(l, r); if (l < 0) frem(a + r, r) else a  (result has r's sign).
static inline float zig_mod_f32(float l, float r) {
    const float a = std::fmod(l, r);
    if (l < 0) return std::fmod(a + r, r);
    return a;
}

// transforms.zig:114-136 / :165-187: which orthogonal fast path (if any) `angle` selects.
static int rotate_class(float angle) {
    const float n = zig_mod_f32(angle, TAU_F);
    const float eps = 1e-6f;
    if (std::fabs(n) < eps || std::fabs(n - TAU_F) < eps) return 1;
    if (std::fabs(n - PI_F / 2.0f) < eps) return 2;
    if (std::fabs(n - PI_F) < eps) return 3;
    if (std::fabs(n - 3.0f * PI_F / 2.0f) < eps) return 4;
    return 0;
}

// image.zig:200-229 setBorder(rect, zero) with rect = (l, t, r, b) clipped to the image
template <typename PX>
static void set_border_zero(const Img<typename PX::T>& out, uint32_t l, uint32_t t, uint32_t r, uint32_t b) {
    const uint32_t il = std::min(l, out.cols), ir = std::min(r, out.cols), it = std::min(t, out.rows), ib = std::min(b, out.rows);
    if (il >= ir || it >= ib) {  // no intersection -> fill
        for (uint32_t y = 0; y < out.rows; ++y)
            for (uint32_t x = 0; x < out.cols; ++x) out.at(y, x) = PX::zero();
        return;
    }
    for (uint32_t y = 0; y < out.rows; ++y)
        for (uint32_t x = 0; x < out.cols; ++x)
            if (y < it || y >= ib || x < il || x >= ir) out.at(y, x) = PX::zero();
}

// transforms.zig:385-462.  kind: 1 = 0deg, 2 = 90 CCW, 3 = 180, 4 = 270 CCW.
template <typename PX>
static void rotate_orthogonal(const Img<typename PX::T>& self, const Img<typename PX::T>& out, int kind) {
    const bool swap = (kind == 2 || kind == 4);
    const uint32_t content_rows = swap ? self.cols : self.rows;
    const uint32_t content_cols = swap ? self.rows : self.cols;
    const uint32_t offset_r = (out.rows > content_rows ? out.rows - content_rows : 0) / 2;  // -| then /2
    const uint32_t offset_c = (out.cols > content_cols ? out.cols - content_cols : 0) / 2;
    for (uint32_t r = 0; r < self.rows; ++r)
        for (uint32_t c = 0; c < self.cols; ++c) {
            size_t new_r, new_c;
            switch (kind) {
                case 1: new_r = r; new_c = c; break;
                case 2: new_r = self.cols - 1 - c; new_c = r; break;
                case 3: new_r = self.rows - 1 - r; new_c = self.cols - 1 - c; break;
                default: new_r = c; new_c = self.rows - 1 - r; break;
            }
            new_r += offset_r;
            new_c += offset_c;
            if (new_r < out.rows && new_c < out.cols) out.at(new_r, new_c) = self.at(r, c);
        }
    if (offset_r != 0 || offset_c != 0) set_border_zero<PX>(out, offset_c, offset_r, offset_c + content_cols, offset_r + content_rows);
}

// transforms.zig:163-212
template <typename PX>
static void rotate_into(const zo_image* src, zo_image* dst, float angle, float cos_a, float sin_a, int method, float mb,
                        float mc, int border) {
    using T = typename PX::T;
    Img<T> self(src), out(dst);
    const int cls = rotate_class(angle);
    if (cls != 0) { rotate_orthogonal<PX>(self, out, cls); return; }
    const float cx = (float)self.cols / 2.0f, cy = (float)self.rows / 2.0f;  // image.zig:322-327
    const float offset_x = ((float)out.cols - (float)self.cols) / 2.0f;
    const float offset_y = ((float)out.rows - (float)self.rows) / 2.0f;
    const float rcx = cx + offset_x, rcy = cy + offset_y;
#pragma omp parallel for schedule(static) num_threads(g_threads) if (g_threads > 1)
    for (uint32_t r = 0; r < out.rows; ++r) {
        const float y = (float)r;
        for (uint32_t c = 0; c < out.cols; ++c) {
            const float x = (float)c;
            const float dx = x - rcx;
            const float dy = y - rcy;
            const float rotated_dx = cos_a * dx - sin_a * dy;
            const float rotated_dy = sin_a * dx + cos_a * dy;
            const float src_x = rotated_dx + cx;
            const float src_y = rotated_dy + cy;
            T val;
            if (!interpolate<PX>(self, src_x, src_y, method, mb, mc, border, &val)) val = PX::zero();
            out.at(r, c) = val;
        }
    }
}

// geometry/transforms.zig:39-42 / :147-150: matrix.dot(src).add(bias); SMatrix.gemm scalar tail
// (SMatrix.zig:554-560): acc = 0; acc += a_ik*b_k ...; result = 0 + 1*acc.
static inline void project_affine(const float* m, float x, float y, float* ox, float* oy) {
    float a0 = 0; a0 += m[0] * x; a0 += m[1] * y; a0 = 0.0f + 1.0f * a0;
    float a1 = 0; a1 += m[2] * x; a1 += m[3] * y; a1 = 0.0f + 1.0f * a1;
    *ox = a0 + m[4];
    *oy = a1 + m[5];
}
// geometry/transforms.zig:224-231
static inline void project_projective(const float* m, float x, float y, float* ox, float* oy) {
    float d[3];
    for (int i = 0; i < 3; ++i) {
        float a = 0; a += m[3 * i + 0] * x; a += m[3 * i + 1] * y; a += m[3 * i + 2] * 1.0f;
        d[i] = 0.0f + 1.0f * a;
    }
    if (d[2] != 0) {
        const float s = 1 / d[2];
        d[0] = d[0] * s; d[1] = d[1] * s;  // SMatrix.scale
    }
    *ox = d[0];
    *oy = d[1];
}

// transforms.zig:522-531
template <typename PX>
static void warp(const zo_image* src, zo_image* dst, int kind, const float* m, int method, float mb, float mc) {
    using T = typename PX::T;
    Img<T> self(src), out(dst);
#pragma omp parallel for schedule(static) num_threads(g_threads) if (g_threads > 1)
    for (uint32_t r = 0; r < out.rows; ++r)
        for (uint32_t c = 0; c < out.cols; ++c) {
            float sx, sy;
            if (kind == ZO_XFORM_PROJECTIVE) project_projective(m, (float)c, (float)r, &sx, &sy);
            else project_affine(m, (float)c, (float)r, &sx, &sy);
            T val;
            if (!interpolate<PX>(self, sx, sy, method, mb, mc, ZO_BORDER_MIRROR, &val)) val = PX::zero();
            out.at(r, c) = val;
        }
}

// transforms.zig:232-283 (extract) with :465-518 (copyRect, the axis-aligned no-resampling fast path)
template <typename PX>
static void extract(const zo_image* src, zo_image* dst, float rl, float rt, float rr, float rb, float angle, float cos_a, float sin_a,
                    int method, float mb, float mc, int border) {
    using T = typename PX::T;
    Img<T> self(src), out(dst);
    if (out.rows == 0 || out.cols == 0) return;
    const float frows = (float)out.rows, fcols = (float)out.cols;
    const float width = rl >= rr ? 0.0f : rr - rl, height = rt >= rb ? 0.0f : rb - rt;   // Rectangle(f32).width / height, Rectangle.zig:76-93
    const float epsilon = 1e-6f;
    if (std::fabs(angle) < epsilon && std::fabs(width - fcols) < epsilon && std::fabs(height - frows) < epsilon) {
        const int32_t rect_top = (int32_t)std::round(rt), rect_left = (int32_t)std::round(rl);
        for (uint32_t r = 0; r < out.rows; ++r)
            for (uint32_t c = 0; c < out.cols; ++c) {
                const int64_t sr = resolve_index((int64_t)r + rect_top, (int64_t)self.rows, border);
                const int64_t sc = resolve_index((int64_t)c + rect_left, (int64_t)self.cols, border);
                out.at(r, c) = (sr < 0 || sc < 0) ? PX::zero() : self.at((uint32_t)sr, (uint32_t)sc);
            }
        return;
    }
    const float cx = (rl + rr) * 0.5f, cy = (rt + rb) * 0.5f;
    for (uint32_t r = 0; r < out.rows; ++r) {
        const float ty = out.rows == 1 ? 0.5f : (float)r / (frows - 1);
        const float y_rect = rt + ty * height;
        for (uint32_t c = 0; c < out.cols; ++c) {
            const float tx = out.cols == 1 ? 0.5f : (float)c / (fcols - 1);
            const float x_rect = rl + tx * width;
            const float dx = x_rect - cx, dy = y_rect - cy;
            const float src_x = cx + cos_a * dx - sin_a * dy;
            const float src_y = cy + sin_a * dx + cos_a * dy;
            T val;
            if (!interpolate<PX>(self, src_x, src_y, method, mb, mc, border, &val)) val = PX::zero();
            out.at(r, c) = val;
        }
    }
}

// blending.zig:26-156 blendColors(u8, base, overlay, mode): Rgba(u8) in / out, f32 arithmetic in the reference's operation order
// (separately rounded; the Makefile's -ffp-contract=off keeps it that way).  mode follows the enum order of blending.zig:8-22.
static Rgba8 blend_rgba8(Rgba8 base, Rgba8 overlay, int mode) {
    if (mode == ZO_BLEND_NONE) return overlay;                                   // :27
    if (overlay.v[3] == 0) return base;                                          // :30
    if (base.v[3] == 0) return overlay;                                          // :33
    if (mode == ZO_BLEND_NORMAL && overlay.v[3] == 255) return overlay;          // :36
    float b[4], o[4], bl[3];
    for (int k = 0; k < 4; ++k) { b[k] = (float)base.v[k] / 255.0f; o[k] = (float)overlay.v[k] / 255.0f; }   // color.zig:484-494 as(f32)
    for (int k = 0; k < 3; ++k) {
        const float bv = b[k], ov = o[k];
        float r;
        switch (mode) {
            case ZO_BLEND_NORMAL: r = ov; break;
            case ZO_BLEND_MULTIPLY: r = bv * ov; break;
            case ZO_BLEND_SCREEN: r = 1.0f - (1.0f - bv) * (1.0f - ov); break;
            case ZO_BLEND_OVERLAY: r = bv < 0.5f ? 2.0f * bv * ov : 1.0f - 2.0f * (1.0f - bv) * (1.0f - ov); break;
            case ZO_BLEND_SOFT_LIGHT:
                r = ov <= 0.5f ? bv - (1.0f - 2.0f * ov) * bv * (1.0f - bv) : bv + (2.0f * ov - 1.0f) * (std::sqrt(bv) - bv);
                break;
            case ZO_BLEND_HARD_LIGHT: r = ov < 0.5f ? 2.0f * ov * bv : 1.0f - 2.0f * (1.0f - ov) * (1.0f - bv); break;
            case ZO_BLEND_COLOR_DODGE: r = bv == 0.0f ? 0.0f : (ov >= 1.0f ? 1.0f : std::fmin(1.0f, bv / (1.0f - ov))); break;
            case ZO_BLEND_COLOR_BURN: r = bv >= 1.0f ? 1.0f : (ov <= 0.0f ? 0.0f : std::fmax(0.0f, 1.0f - (1.0f - bv) / ov)); break;
            case ZO_BLEND_DARKEN: r = std::fmin(bv, ov); break;
            case ZO_BLEND_LIGHTEN: r = std::fmax(bv, ov); break;
            case ZO_BLEND_DIFFERENCE: r = std::fabs(bv - ov); break;
            default: r = bv + ov - 2.0f * bv * ov; break;                        // exclusion
        }
        bl[k] = r;
    }
    float out[4];
    if (overlay.v[3] == 255) {                                                   // :127-136
        out[0] = bl[0]; out[1] = bl[1]; out[2] = bl[2]; out[3] = 1.0f;
    } else {                                                                     // :137-154 Porter-Duff "over"
        const float result_a = o[3] + b[3] * (1.0f - o[3]);
        if (result_a <= 0) return Rgba8{{0, 0, 0, 0}};
        const float base_weight = b[3] * (1.0f - o[3]);
        const float inv_result_a = 1.0f / result_a;
        for (int k = 0; k < 3; ++k) out[k] = (bl[k] * o[3] + b[k] * base_weight) * inv_result_a;
        out[3] = result_a;
    }
    Rgba8 res;
    for (int k = 0; k < 4; ++k) {                                                // color.zig:496-501 as(u8): @round(255 * clamp(v, 0, 1))
        const float c = std::fmax(0.0f, std::fmin(out[k], 1.0f));
        res.v[k] = (uint8_t)std::round(255.0f * c);
    }
    return res;
}

// image.zig:67-95 assignPixel for source and destination of the same pixel type: only Rgba(u8) samples blend (:74).
template <typename T> static inline void assign_pixel(T& dest, const T& sample, int) { dest = sample; }
template <> inline void assign_pixel<Rgba8>(Rgba8& dest, const Rgba8& sample, int blend) {
    dest = blend == ZO_BLEND_NONE ? sample : blend_rgba8(dest, sample, blend);
}

// transforms.zig:293-376 (insert, source and destination of the same pixel type): a gather over the
// destination pixels of the rotated rectangle's bounding box; pixels outside the rectangle or with a null sample stay untouched.
template <typename PX>
static void insert(zo_image* self_img, const zo_image* source_img, float rl, float rt, float rr, float rb, float angle, float cos_a,
                   float sin_a, int method, float mb, float mc, int blend) {
    using T = typename PX::T;
    Img<T> self(self_img), source(source_img);
    if (source.rows == 0 || source.cols == 0) return;
    const float frows = (float)source.rows, fcols = (float)source.cols;
    const float rect_width = rl >= rr ? 0.0f : rr - rl, rect_height = rt >= rb ? 0.0f : rb - rt;
    const float epsilon = 1e-6f;
    if (std::fabs(angle) < epsilon && std::fabs(rect_width - fcols) < epsilon && std::fabs(rect_height - frows) < epsilon) {
        const int32_t dst_top = (int32_t)std::round(rt), dst_left = (int32_t)std::round(rl);
        for (uint32_t r = 0; r < source.rows; ++r) {
            const int64_t y = (int64_t)dst_top + r;
            for (uint32_t c = 0; c < source.cols; ++c) {
                const int64_t x = (int64_t)dst_left + c;
                if (y >= 0 && y < (int64_t)self.rows && x >= 0 && x < (int64_t)self.cols) assign_pixel(self.at((uint32_t)y, (uint32_t)x), source.at(r, c), blend);
            }
        }
        return;
    }
    const float cx = (rl + rr) * 0.5f, cy = (rt + rb) * 0.5f;
    const float inv_width = 1.0f / rect_width, inv_height = 1.0f / rect_height;
    const float half_width = rect_width * 0.5f, half_height = rect_height * 0.5f;
    const float abs_cos = std::fabs(cos_a), abs_sin = std::fabs(sin_a);
    const float bound_hw = half_width * abs_cos + half_height * abs_sin;
    const float bound_hh = half_width * abs_sin + half_height * abs_cos;
    auto to_u32 = [](float v) -> uint32_t { return v <= 0 ? 0u : (v >= 4294967040.0f ? 4294967295u : (uint32_t)v); };
    const uint32_t min_r = (cy - bound_hh < 0) ? 0u : to_u32(std::floor(cy - bound_hh));
    const uint32_t max_r = (uint32_t)std::min<uint64_t>(self.rows, (uint64_t)to_u32(std::ceil(cy + bound_hh)) + 1);
    const uint32_t min_c = (cx - bound_hw < 0) ? 0u : to_u32(std::floor(cx - bound_hw));
    const uint32_t max_c = (uint32_t)std::min<uint64_t>(self.cols, (uint64_t)to_u32(std::ceil(cx + bound_hw)) + 1);
    for (uint32_t r = min_r; r < max_r; ++r) {
        const float dy = (float)r - cy;
        for (uint32_t c = min_c; c < max_c; ++c) {
            const float dx = (float)c - cx;
            const float rect_x = cos_a * dx + sin_a * dy;
            const float rect_y = -sin_a * dx + cos_a * dy;
            if (std::fabs(rect_x) > half_width || std::fabs(rect_y) > half_height) continue;
            const float norm_x = (rect_x + half_width) * inv_width;
            const float norm_y = (rect_y + half_height) * inv_height;
            const float src_x = source.cols == 1 ? 0.0f : norm_x * (fcols - 1);
            const float src_y = source.rows == 1 ? 0.0f : norm_y * (frows - 1);
            T val;
            if (interpolate<PX>(source, src_x, src_y, method, mb, mc, ZO_BORDER_MIRROR, &val)) assign_pixel(self.at(r, c), val, blend);
        }
    }
}

}  // namespace zo

extern "C" {

void zo_blend_rgba8(const uint8_t* base, const uint8_t* overlay, int mode, uint8_t* out) {
    const zo::Rgba8 r = zo::blend_rgba8(zo::Rgba8{{base[0], base[1], base[2], base[3]}}, zo::Rgba8{{overlay[0], overlay[1], overlay[2], overlay[3]}}, mode);
    for (int k = 0; k < 4; ++k) out[k] = r.v[k];
}

int zo_insert(zo_image* self, const zo_image* source, int pixfmt, float rl, float rt, float rr, float rb, float angle, float cos_a, float sin_a,
              int method, float mb, float mc) {
    return zo_insert_blend(self, source, pixfmt, rl, rt, rr, rb, angle, cos_a, sin_a, method, mb, mc, ZO_BLEND_NONE);
}

int zo_insert_blend(zo_image* self, const zo_image* source, int pixfmt, float rl, float rt, float rr, float rb, float angle, float cos_a,
                    float sin_a, int method, float mb, float mc, int blend) {
    using namespace zo;
    if (blend < ZO_BLEND_NONE || blend > ZO_BLEND_EXCLUSION) return ZO_ERR_INVALID_ARGUMENT;
    switch (pixfmt) {
        case ZO_PIX_U8: insert<PxU8>(self, source, rl, rt, rr, rb, angle, cos_a, sin_a, method, mb, mc, blend); return ZO_OK;
        case ZO_PIX_F32: insert<PxF32>(self, source, rl, rt, rr, rb, angle, cos_a, sin_a, method, mb, mc, blend); return ZO_OK;
        case ZO_PIX_RGB8: insert<PxRgb8>(self, source, rl, rt, rr, rb, angle, cos_a, sin_a, method, mb, mc, blend); return ZO_OK;
        case ZO_PIX_RGBA8: insert<PxRgba8>(self, source, rl, rt, rr, rb, angle, cos_a, sin_a, method, mb, mc, blend); return ZO_OK;
        case ZO_PIX_RGBAF32: insert<PxRgbaF32>(self, source, rl, rt, rr, rb, angle, cos_a, sin_a, method, mb, mc, blend); return ZO_OK;
    }
    return ZO_ERR_UNSUPPORTED;
}

int zo_extract(const zo_image* src, zo_image* dst, int pixfmt, float rl, float rt, float rr, float rb, float angle, float cos_a,
               float sin_a, int method, float mb, float mc, int border) {
    using namespace zo;
    switch (pixfmt) {
        case ZO_PIX_U8: extract<PxU8>(src, dst, rl, rt, rr, rb, angle, cos_a, sin_a, method, mb, mc, border); return ZO_OK;
        case ZO_PIX_F32: extract<PxF32>(src, dst, rl, rt, rr, rb, angle, cos_a, sin_a, method, mb, mc, border); return ZO_OK;
        case ZO_PIX_RGB8: extract<PxRgb8>(src, dst, rl, rt, rr, rb, angle, cos_a, sin_a, method, mb, mc, border); return ZO_OK;
        case ZO_PIX_RGBA8: extract<PxRgba8>(src, dst, rl, rt, rr, rb, angle, cos_a, sin_a, method, mb, mc, border); return ZO_OK;
        case ZO_PIX_RGBAF32: extract<PxRgbaF32>(src, dst, rl, rt, rr, rb, angle, cos_a, sin_a, method, mb, mc, border); return ZO_OK;
    }
    return ZO_ERR_UNSUPPORTED;
}

int zo_rotate_class(float angle) { return zo::rotate_class(angle); }

void zo_rotate_bounds(uint32_t rows, uint32_t cols, float angle, uint32_t* out_rows, uint32_t* out_cols) {
    const int cls = zo::rotate_class(angle);
    if (cls == 1 || cls == 3) { *out_rows = rows; *out_cols = cols; return; }
    if (cls == 2 || cls == 4) { *out_rows = cols; *out_cols = rows; return; }
    const float cos_abs = std::fabs(std::cos(angle)), sin_abs = std::fabs(std::sin(angle));
    const float w = (float)cols, h = (float)rows;
    const float new_w = w * cos_abs + h * sin_abs;
    const float new_h = h * cos_abs + w * sin_abs;
    *out_cols = (uint32_t)std::ceil(new_w);
    *out_rows = (uint32_t)std::ceil(new_h);
}

int zo_rotate_into(const zo_image* src, zo_image* dst, int pixfmt, float angle, float cos_a, float sin_a, int method,
                   float mb, float mc, int border) {
    using namespace zo;
    switch (pixfmt) {
        case ZO_PIX_U8: rotate_into<PxU8>(src, dst, angle, cos_a, sin_a, method, mb, mc, border); return ZO_OK;
        case ZO_PIX_F32: rotate_into<PxF32>(src, dst, angle, cos_a, sin_a, method, mb, mc, border); return ZO_OK;
        case ZO_PIX_RGB8: rotate_into<PxRgb8>(src, dst, angle, cos_a, sin_a, method, mb, mc, border); return ZO_OK;
        case ZO_PIX_RGBA8: rotate_into<PxRgba8>(src, dst, angle, cos_a, sin_a, method, mb, mc, border); return ZO_OK;
        case ZO_PIX_RGBAF32: rotate_into<PxRgbaF32>(src, dst, angle, cos_a, sin_a, method, mb, mc, border); return ZO_OK;
    }
    return ZO_ERR_UNSUPPORTED;
}

int zo_warp(const zo_image* src, zo_image* dst, int pixfmt, int kind, const float* m, int method, float mb, float mc) {
    using namespace zo;
    switch (pixfmt) {
        case ZO_PIX_U8: warp<PxU8>(src, dst, kind, m, method, mb, mc); return ZO_OK;
        case ZO_PIX_F32: warp<PxF32>(src, dst, kind, m, method, mb, mc); return ZO_OK;
        case ZO_PIX_RGB8: warp<PxRgb8>(src, dst, kind, m, method, mb, mc); return ZO_OK;
        case ZO_PIX_RGBA8: warp<PxRgba8>(src, dst, kind, m, method, mb, mc); return ZO_OK;
        case ZO_PIX_RGBAF32: warp<PxRgbaF32>(src, dst, kind, m, method, mb, mc); return ZO_OK;
    }
    return ZO_ERR_UNSUPPORTED;
}

}  // extern "C"
