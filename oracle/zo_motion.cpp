// zo_motion.cpp -- CPU oracle (TEST INFRASTRUCTURE ONLY, see zignal_oracle.h) for the gather forms of Image.motionBlur
// (reference image/motion_blur.zig): the diagonal branch of MotionBlurOps.linear (:115-250) and MotionBlurOps.radial (:252-436).
// The axis-aligned branches of linear are convolveSeparable calls (:80-114) and are composed in tests/oracle_lib.py.
// f32 arithmetic in the reference's order; cos / sin of the motion angle cross the interface as data, like rotateInto's.
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

struct View {
    const zo_image* im;
    int ch;
    bool flt;
    float at(size_t r, size_t c, int k) const {
        const size_t i = (r * (size_t)im->stride + c) * ch + k;
        return flt ? ((const float*)im->data)[i] : (float)((const uint8_t*)im->data)[i];
    }
};

// The bilinear tap shared by both filters (motion_blur.zig:135-152, :316-333): x1 / y1 clamp to the last column / row.
inline float bilinear(const View& v, float sx, float sy, int k) {
    const size_t x0 = (size_t)std::floor(sx), y0 = (size_t)std::floor(sy);
    const size_t x1 = std::min(x0 + 1, (size_t)v.im->cols - 1), y1 = std::min(y0 + 1, (size_t)v.im->rows - 1);
    const float fx = sx - (float)x0, fy = sy - (float)y0;
    const float v00 = v.at(y0, x0, k), v10 = v.at(y0, x1, k), v01 = v.at(y1, x0, k), v11 = v.at(y1, x1, k);
    const float v0 = v00 * (1 - fx) + v10 * fx;
    const float v1 = v01 * (1 - fx) + v11 * fx;
    return v0 * (1 - fy) + v1 * fy;
}

inline void store(zo_image* dst, int ch, bool flt, size_t r, size_t c, int k, float result) {
    const size_t i = (r * (size_t)dst->stride + c) * ch + k;
    if (flt) ((float*)dst->data)[i] = result;
    else ((uint8_t*)dst->data)[i] = (uint8_t)std::trunc(std::fmax(0.0f, std::fmin(255.0f, std::round(result))));   // :160-161
}

}  // namespace
}  // namespace zo

using namespace zo;

extern "C" int zo_motion_blur_line(const zo_image* src, zo_image* dst, int pixfmt, float cos_angle, float sin_angle, uint32_t distance) {
    if (src->rows != dst->rows || src->cols != dst->cols) return ZO_ERR_DIMENSION_MISMATCH;
    const int ch = channels(pixfmt);
    if (!ch) return ZO_ERR_UNSUPPORTED;
    if (distance == 0) return ZO_ERR_INVALID_ARGUMENT;                 // :66-69 copies; the caller handles it
    const bool flt = is_float(pixfmt);
    const View v{src, ch, flt};
    const float half_dist = (float)distance / 2.0f;                    // :74
    const float fcols = (float)src->cols, frows = (float)src->rows;
    for (size_t r = 0; r < src->rows; ++r)
        for (size_t c = 0; c < src->cols; ++c)
            for (int k = 0; k < ch; ++k) {
                float sum = 0, count = 0, t = -half_dist;
                for (uint64_t it = 0; it < (uint64_t)distance + 2; ++it) {                   // :125-156
                    if (t > half_dist) break;
                    const float sx = (float)c + t * cos_angle, sy = (float)r + t * sin_angle;
                    if (sx >= 0 && sx < fcols && sy >= 0 && sy < frows) {
                        sum += bilinear(v, sx, sy, k);
                        count += 1;
                    }
                    t += 1.0f;
                }
                store(dst, ch, flt, r, c, k, count > 0 ? sum / count : v.at(r, c, k));        // :158
            }
    return ZO_OK;
}

extern "C" int zo_motion_blur_radial(const zo_image* src, zo_image* dst, int pixfmt, float center_x, float center_y, float strength, int spin) {
    if (src->rows != dst->rows || src->cols != dst->cols) return ZO_ERR_DIMENSION_MISMATCH;
    const int ch = channels(pixfmt);
    if (!ch) return ZO_ERR_UNSUPPORTED;
    const bool flt = is_float(pixfmt);
    if (src->rows == 0 || src->cols == 0) return ZO_OK;
    if (strength == 0) {                                                                      // :262-265
        const size_t pb = (size_t)ch * (flt ? 4 : 1);
        if (src->data != dst->data)
            for (uint32_t r = 0; r < src->rows; ++r)
                std::memcpy((uint8_t*)dst->data + r * dst->stride * pb, (const uint8_t*)src->data + r * src->stride * pb, (size_t)src->cols * pb);
        return ZO_OK;
    }
    const View v{src, ch, flt};
    const float cx = center_x * (float)(src->cols - 1), cy = center_y * (float)(src->rows - 1);          // :268-269
    const float clamped = std::fmax(0.0f, std::fmin(1.0f, strength));                                   // :272
    const size_t num_samples = 8 + (size_t)std::trunc(clamped * 24.0f);                                 // :275-277
    const float fcols = (float)src->cols, frows = (float)src->rows;
    for (size_t r = 0; r < src->rows; ++r)
        for (size_t c = 0; c < src->cols; ++c) {
            const float fx = (float)c, fy = (float)r;
            const float dx = fx - cx, dy = fy - cy;
            const float distance = std::sqrt(dx * dx + dy * dy);
            const float angle = std::atan2(dy, dx);
            const float max_distance = std::sqrt(cx * cx + cy * cy);
            const float blur_amount = spin ? clamped * 0.5f : (distance / max_distance) * clamped * 20.0f;   // :294-297
            for (int k = 0; k < ch; ++k) {
                float sum = 0;
                size_t count = 0;
                for (size_t s = 0; s < num_samples; ++s) {
                    const float t = ((float)s - (float)(num_samples - 1) / 2.0f) / (float)(num_samples - 1);   // :304
                    float sx, sy;
                    if (!spin) {
                        const float scale = 1.0f + t * blur_amount * 0.1f;                   // :310-313
                        sx = cx + dx * scale;
                        sy = cy + dy * scale;
                    } else {
                        const float new_angle = angle + t * blur_amount;                     // :315-319
                        sx = cx + distance * std::cos(new_angle);
                        sy = cy + distance * std::sin(new_angle);
                    }
                    if (sx >= 0 && sx < fcols && sy >= 0 && sy < frows) {
                        sum += bilinear(v, sx, sy, k);
                        count += 1;
                    }
                }
                store(dst, ch, flt, r, c, k, count > 0 ? sum / (float)count : v.at(r, c, k));   // :346
            }
        }
    return ZO_OK;
}
