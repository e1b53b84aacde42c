// zo_integral.cpp -- CPU oracle for Image.boxBlur / Image.sharpen (f32 summed-area tables).
// TEST INFRASTRUCTURE ONLY (see zignal_oracle.h).  Restates image/integral.zig:41-90 (plane, sum),
// :95-143 (compute), :148-191 (boxBlur), :194-269 (boxBlurPlane), :273-322 (sharpen),
// :325-422 (sharpenPlane) and image.zig:635-648, :785-799.
#include "zo_common.h"

namespace zo {

// integral.zig:41-78.  Row pass is a sequential f32 running sum; column pass adds row r-1 into row r
// (sequential per column).  Both orders are part of the result once sums exceed 2^24.
template <typename T>
static void integral_plane(const Img<T>& src, const Img<float>& dst) {
    const size_t rows = src.rows, cols = src.cols;
#pragma omp parallel for schedule(static) num_threads(g_threads) if (g_threads > 1)
    for (size_t r = 0; r < rows; ++r) {
        float tmp = 0;
        const T* s = src.data + r * src.stride;
        float* d = dst.data + r * dst.stride;
        for (size_t c = 0; c < cols; ++c) {
            tmp += (float)s[c];
            d[c] = tmp;
        }
    }
    for (size_t r = 1; r < rows; ++r) {
        const float* prev = dst.data + (r - 1) * dst.stride;
        float* cur = dst.data + r * dst.stride;
        for (size_t c = 0; c < cols; ++c) cur[c] = prev[c] + cur[c];  // :69 prev_vec + curr_vec
    }
}

// integral.zig:85-90: ((D - left) - top) + corner
static inline float sat_sum(const Img<float>& sat, size_t r1, size_t c1, size_t r2, size_t c2) {
    return sat.data[r2 * sat.stride + c2] - (c1 > 0 ? sat.data[r2 * sat.stride + (c1 - 1)] : 0.0f) -
           (r1 > 0 ? sat.data[(r1 - 1) * sat.stride + c2] : 0.0f) +
           ((r1 > 0 && c1 > 0) ? sat.data[(r1 - 1) * sat.stride + (c1 - 1)] : 0.0f);
}

template <typename P>
static inline P store_plane(float v);
template <> inline uint8_t store_plane<uint8_t>(float v) { return clamp_u8(v); }  // meta.clamp(u8, f32)
template <> inline float store_plane<float>(float v) { return v; }               // meta.clamp(f32, v) == as(f32, v)

// integral.zig:194-269.  The SIMD interior (:227-251) evaluates int22 - int21 - int12 + int11 which
// is the same left-to-right order as `sum` (:86-89), and the same division, so one formula covers
// every branch: windows are clipped to the image and the divisor is the clipped area.
template <typename P>
static void box_blur_plane(const Img<float>& sat, const Img<P>& dst, size_t radius) {
    const size_t rows = sat.rows, cols = sat.cols;
#pragma omp parallel for schedule(static) num_threads(g_threads) if (g_threads > 1)
    for (size_t r = 0; r < rows; ++r) {
        const size_t r1 = r >= radius ? r - radius : 0;  // r -| radius
        const size_t r2 = std::min(r + radius, rows - 1);
        for (size_t c = 0; c < cols; ++c) {
            const size_t c1 = c >= radius ? c - radius : 0;
            const size_t c2 = std::min(c + radius, cols - 1);
            const float area = (float)((r2 - r1 + 1) * (c2 - c1 + 1));
            const float s = sat_sum(sat, r1, c1, r2, c2);
            const float val = s / area;
            dst.data[r * dst.stride + c] = store_plane<P>(val);
        }
    }
}

// integral.zig:325-422: sharpened = 2*original - blurred
template <typename P>
static void sharpen_plane(const Img<P>& src, const Img<float>& sat, const Img<P>& dst, size_t radius) {
    const size_t rows = sat.rows, cols = sat.cols;
#pragma omp parallel for schedule(static) num_threads(g_threads) if (g_threads > 1)
    for (size_t r = 0; r < rows; ++r) {
        const size_t r1 = r >= radius ? r - radius : 0;
        const size_t r2 = std::min(r + radius, rows - 1);
        for (size_t c = 0; c < cols; ++c) {
            const size_t c1 = c >= radius ? c - radius : 0;
            const size_t c2 = std::min(c + radius, cols - 1);
            const float area = (float)((r2 - r1 + 1) * (c2 - c1 + 1));
            const float s = sat_sum(sat, r1, c1, r2, c2);
            const float blurred = s / area;
            const float original = (float)src.data[r * src.stride + c];
            const float sharpened = 2 * original - blurred;
            dst.data[r * dst.stride + c] = store_plane<P>(sharpened);
        }
    }
}

// integral.zig:112-139: struct images extract each channel into an f32 plane, then Integral(f32).plane.
template <typename CT>
static void integral_struct(const CT* base, int ch, uint32_t rows, uint32_t cols, size_t stride_px,
                            std::vector<Plane<float>>& sats) {
    Plane<float> src_plane(rows, cols);
    for (int k = 0; k < ch; ++k) {
        for (uint32_t r = 0; r < rows; ++r)
            for (uint32_t c = 0; c < cols; ++c) src_plane.buf[(size_t)r * cols + c] = (float)base[(r * stride_px + c) * ch + k];
        sats.emplace_back(rows, cols);
        integral_plane<float>(src_plane.img(), sats.back().img());
    }
}

static void copy_rows(const zo_image* src, zo_image* dst, int pixfmt) {
    if (src->data == dst->data) return;
    const size_t pb = pixel_bytes(pixfmt);
    for (uint32_t r = 0; r < src->rows; ++r)
        std::memcpy((uint8_t*)dst->data + r * dst->stride * pb, (const uint8_t*)src->data + r * src->stride * pb,
                    (size_t)src->cols * pb);
}

template <typename CT>
static void box_blur_struct(const zo_image* src, zo_image* dst, int ch, size_t radius) {
    const uint32_t rows = src->rows, cols = src->cols;
    std::vector<Plane<float>> sats;
    sats.reserve(ch);
    integral_struct<CT>((const CT*)src->data, ch, rows, cols, (size_t)src->stride, sats);
    Plane<float> scratch(rows, cols);  // integral.zig:169
    CT* d = (CT*)dst->data;
    for (int k = 0; k < ch; ++k) {
        box_blur_plane<float>(sats[k].img(), scratch.img(), radius);
        for (uint32_t r = 0; r < rows; ++r)
            for (uint32_t c = 0; c < cols; ++c) {
                const float v = scratch.buf[(size_t)r * cols + c];
                d[(r * dst->stride + c) * ch + k] = store_plane<CT>(v);  // :180-184
            }
    }
}

template <typename CT>
static void sharpen_struct(const zo_image* src, zo_image* dst, int ch, size_t radius) {
    const uint32_t rows = src->rows, cols = src->cols;
    std::vector<Plane<float>> sats;
    sats.reserve(ch);
    integral_struct<CT>((const CT*)src->data, ch, rows, cols, (size_t)src->stride, sats);
    const CT* s = (const CT*)src->data;
    CT* d = (CT*)dst->data;
    for (size_t r = 0; r < rows; ++r) {  // integral.zig:296-318
        const size_t r1 = r >= radius ? r - radius : 0;
        const size_t r2 = std::min(r + radius, (size_t)rows - 1);
        for (size_t c = 0; c < cols; ++c) {
            const size_t c1 = c >= radius ? c - radius : 0;
            const size_t c2 = std::min(c + radius, (size_t)cols - 1);
            const float area = (float)((r2 - r1 + 1) * (c2 - c1 + 1));
            for (int k = 0; k < ch; ++k) {
                const float channel_sum = sat_sum(sats[k].img(), r1, c1, r2, c2);
                const float blurred = channel_sum / area;
                const float original = (float)s[(r * src->stride + c) * ch + k];
                // :311 @trunc(@max(minInt, @min(maxInt, @round(v)))) == meta.clamp for finite v
                d[(r * dst->stride + c) * ch + k] = store_plane<CT>(2 * original - blurred);
            }
        }
    }
}

}  // namespace zo

extern "C" {

int zo_integral_plane(const zo_image* src, int pixfmt, float* sat) {
    zo::Img<float> d(sat, src->rows, src->cols, src->cols);
    if (pixfmt == ZO_PIX_U8) zo::integral_plane<uint8_t>(zo::Img<uint8_t>(src), d);
    else if (pixfmt == ZO_PIX_F32) zo::integral_plane<float>(zo::Img<float>(src), d);
    else return ZO_ERR_UNSUPPORTED;
    return ZO_OK;
}

int zo_box_blur(const zo_image* src, zo_image* dst, int pixfmt, uint32_t radius) {
    if (src->rows != dst->rows || src->cols != dst->cols) return ZO_ERR_DIMENSION_MISMATCH;  // image.zig:636
    if (radius == 0) { zo::copy_rows(src, dst, pixfmt); return ZO_OK; }                       // :639
    if ((size_t)src->rows * src->cols == 0) return ZO_OK;                                     // integral.zig:100
    const uint32_t rows = src->rows, cols = src->cols;
    switch (pixfmt) {
        case ZO_PIX_U8: {
            zo::Plane<float> sat(rows, cols);
            zo::integral_plane<uint8_t>(zo::Img<uint8_t>(src), sat.img());
            zo::box_blur_plane<uint8_t>(sat.img(), zo::Img<uint8_t>(dst), radius);
            return ZO_OK;
        }
        case ZO_PIX_F32: {
            zo::Plane<float> sat(rows, cols);
            zo::integral_plane<float>(zo::Img<float>(src), sat.img());
            zo::box_blur_plane<float>(sat.img(), zo::Img<float>(dst), radius);
            return ZO_OK;
        }
        case ZO_PIX_RGB8: zo::box_blur_struct<uint8_t>(src, dst, 3, radius); return ZO_OK;
        case ZO_PIX_RGBA8: zo::box_blur_struct<uint8_t>(src, dst, 4, radius); return ZO_OK;
        case ZO_PIX_RGBAF32: zo::box_blur_struct<float>(src, dst, 4, radius); return ZO_OK;
    }
    return ZO_ERR_UNSUPPORTED;
}

int zo_sharpen(const zo_image* src, zo_image* dst, int pixfmt, uint32_t radius) {
    if (src->rows != dst->rows || src->cols != dst->cols) return ZO_ERR_DIMENSION_MISMATCH;  // image.zig:786
    if (radius == 0) { zo::copy_rows(src, dst, pixfmt); return ZO_OK; }                       // :789
    if ((size_t)src->rows * src->cols == 0) return ZO_OK;
    const uint32_t rows = src->rows, cols = src->cols;
    switch (pixfmt) {
        case ZO_PIX_U8: {
            zo::Plane<float> sat(rows, cols);
            zo::integral_plane<uint8_t>(zo::Img<uint8_t>(src), sat.img());
            zo::sharpen_plane<uint8_t>(zo::Img<uint8_t>(src), sat.img(), zo::Img<uint8_t>(dst), radius);
            return ZO_OK;
        }
        case ZO_PIX_F32: {
            zo::Plane<float> sat(rows, cols);
            zo::integral_plane<float>(zo::Img<float>(src), sat.img());
            zo::sharpen_plane<float>(zo::Img<float>(src), sat.img(), zo::Img<float>(dst), radius);
            return ZO_OK;
        }
        case ZO_PIX_RGB8: zo::sharpen_struct<uint8_t>(src, dst, 3, radius); return ZO_OK;
        case ZO_PIX_RGBA8: zo::sharpen_struct<uint8_t>(src, dst, 4, radius); return ZO_OK;
        case ZO_PIX_RGBAF32: zo::sharpen_struct<float>(src, dst, 4, radius); return ZO_OK;
    }
    return ZO_ERR_UNSUPPORTED;
}

}  // extern "C"
