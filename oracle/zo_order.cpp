// zo_order.cpp -- CPU oracle (TEST INFRASTRUCTURE ONLY, see zignal_oracle.h) for the order-statistic filters:
// Image.medianBlur / percentileBlur / minBlur / maxBlur / midpointBlur / alphaTrimmedMeanBlur
// (reference image.zig:650-790 -> image/order_statistic_blur.zig:22-413, image/histogram.zig:586-612).
// Restates the reference's algorithm: per-column histograms of the (2r+1)-row band, a window histogram slid along each row,
// and a reducer applied to the 256-bin window histogram.  Interleaved 8-bit pixels are filtered channel by channel (:199-229).
#include "zo_common.h"

namespace zo {
namespace {

struct Hist {
    uint32_t v[256];
    Hist() { std::memset(v, 0, sizeof(v)); }
    void add(const Hist& o) { for (int i = 0; i < 256; ++i) v[i] += o.v[i]; }
    void sub(const Hist& o) { for (int i = 0; i < 256; ++i) v[i] -= o.v[i]; }
    size_t total() const { size_t t = 0; for (int i = 0; i < 256; ++i) t += v[i]; return t; }
};

// histogram.zig:586-612 stats.percentile
uint8_t percentile(const Hist& h, double p) {
    const size_t total = h.total();
    if (total == 0) return 0;
    const size_t total_minus_one = total - 1;
    const double rank_f = p * (double)total_minus_one;
    const double rank_floor = std::floor(rank_f + 1e-12);
    size_t rank = (size_t)std::trunc(rank_floor);
    rank = std::min(rank, total_minus_one);
    size_t cumulative = 0;
    for (int value = 0; value < 256; ++value) {
        if (h.v[value] == 0) continue;
        cumulative += h.v[value];
        if (cumulative > rank) return (uint8_t)value;
    }
    return 255;
}

// order_statistic_blur.zig:357-365 MidpointReducer
uint8_t midpoint(const Hist& h) {
    int mn = 0, mx;
    bool found = false;
    for (int i = 0; i < 256; ++i) if (h.v[i] > 0) { mn = i; found = true; break; }
    mx = mn;
    if (found) for (int i = 255; i >= 0; --i) if (h.v[i] > 0) { mx = i; break; }
    return (uint8_t)((mn + mx + 1) / 2);
}

// order_statistic_blur.zig:366-410 AlphaTrimmedMeanReducer (kept_count == 0 cannot happen for an odd window area).
uint8_t alpha_trimmed(const Hist& h, size_t window_area, double trim_fraction) {
    const double trimmed_total = std::floor(trim_fraction * (double)window_area);
    const size_t trimmed_each = (size_t)std::trunc(trimmed_total);
    const size_t trim_each = std::min(trimmed_each, window_area / 2);
    uint64_t total_sum = 0;
    for (int i = 0; i < 256; ++i) total_sum += (uint64_t)h.v[i] * (uint64_t)i;
    uint64_t low_sum = 0, high_sum = 0;
    size_t low_count = 0, high_count = 0, remaining = trim_each;
    for (int i = 0; i < 256 && remaining > 0; ++i) {
        const size_t take = std::min((size_t)h.v[i], remaining);
        low_sum += (uint64_t)take * (uint64_t)i;
        low_count += take;
        remaining -= take;
    }
    remaining = trim_each;
    for (int i = 255; i >= 0 && remaining > 0; --i) {
        if (h.v[i] == 0) continue;
        const size_t take = std::min((size_t)h.v[i], remaining);
        high_sum += (uint64_t)take * (uint64_t)i;
        high_count += take;
        remaining -= take;
    }
    const size_t kept_count = window_area - low_count - high_count;
    if (kept_count == 0) return 0;
    const uint64_t kept_sum = total_sum - low_sum - high_sum;
    const uint64_t rounded = (kept_sum + (uint64_t)kept_count / 2) / (uint64_t)kept_count;
    return (uint8_t)std::min<uint64_t>(255, rounded);
}

uint8_t reduce(const Hist& h, int mode, double param) {
    switch (mode) {
        case ZO_ORDER_PERCENTILE: return percentile(h, param);
        case ZO_ORDER_MIDPOINT: return midpoint(h);
        default: return alpha_trimmed(h, h.total(), param);
    }
}

// order_statistic_blur.zig:231-330 applyScalarOp on one contiguous plane.
void scalar_op(const uint8_t* img, uint32_t rows, uint32_t cols, uint8_t* out, int64_t radius, int border, int mode, double param) {
    const int64_t window = 2 * radius + 1;
    std::vector<Hist> column(cols);
    Hist zero_column;
    zero_column.v[0] = (uint32_t)window;
    auto pixel = [&](int64_t row, int64_t col) -> uint8_t {     // :338-347 getPixel
        const int64_t r = resolve_index(row, rows, border), c = resolve_index(col, cols, border);
        return (r >= 0 && c >= 0) ? img[(size_t)r * cols + (size_t)c] : 0;
    };
    for (uint32_t col = 0; col < cols; ++col)
        for (int64_t off = 0; off < window; ++off) column[col].v[pixel(off - radius, col)] += 1;
    for (uint32_t row = 0; row < rows; ++row) {
        Hist win;
        for (int64_t off = 0; off < window; ++off) {
            const int64_t c = resolve_index(off - radius, cols, border);
            win.add(c >= 0 ? column[(size_t)c] : zero_column);
        }
        out[(size_t)row * cols] = reduce(win, mode, param);
        for (uint32_t col = 1; col < cols; ++col) {
            const int64_t l = resolve_index((int64_t)col - radius - 1, cols, border);
            win.sub(l >= 0 ? column[(size_t)l] : zero_column);
            const int64_t r = resolve_index((int64_t)col + radius, cols, border);
            win.add(r >= 0 ? column[(size_t)r] : zero_column);
            out[(size_t)row * cols + col] = reduce(win, mode, param);
        }
        if (row + 1 == rows) break;
        const int64_t rem = resolve_index((int64_t)row - radius, rows, border), add = resolve_index((int64_t)row + radius + 1, rows, border);
        for (uint32_t col = 0; col < cols; ++col) {
            column[col].v[rem >= 0 ? img[(size_t)rem * cols + col] : 0] -= 1;
            column[col].v[add >= 0 ? img[(size_t)add * cols + col] : 0] += 1;
        }
    }
}

}  // namespace
}  // namespace zo

extern "C" int zo_order_blur(const zo_image* src, zo_image* dst, int pixfmt, uint32_t radius, int mode, double param, int border) {
    if (src->rows != dst->rows || src->cols != dst->cols) return ZO_ERR_DIMENSION_MISMATCH;      // image.zig:679 hasSameShape
    if (mode != ZO_ORDER_PERCENTILE && mode != ZO_ORDER_MIDPOINT && mode != ZO_ORDER_ALPHA_TRIMMED) return ZO_ERR_INVALID_ARGUMENT;
    if (src->rows == 0 || src->cols == 0) return ZO_OK;                                          // order_statistic_blur.zig:39,111,156
    if (mode == ZO_ORDER_ALPHA_TRIMMED && (!std::isfinite(param) || param < 0.0 || param >= 0.5)) return ZO_ERR_INVALID_TRIM;   // :160
    const int ch = pixfmt == ZO_PIX_U8 ? 1 : pixfmt == ZO_PIX_RGB8 ? 3 : pixfmt == ZO_PIX_RGBA8 ? 4 : 0;
    if (radius == 0) {                                                                           // :43-46 image.copy(out)
        const size_t pb = pixfmt == ZO_PIX_F32 ? 4 : pixfmt == ZO_PIX_RGBAF32 ? 16 : (size_t)ch;
        if (src->data != dst->data)
            for (uint32_t r = 0; r < src->rows; ++r)
                std::memcpy((uint8_t*)dst->data + r * dst->stride * pb, (const uint8_t*)src->data + r * src->stride * pb, (size_t)src->cols * pb);
        return ZO_OK;
    }
    if (mode == ZO_ORDER_PERCENTILE && !(param >= 0.0 && param <= 1.0)) return ZO_ERR_INVALID_PERCENTILE;   // :48-50 (NaN: the reference asserts)
    if (ch == 0) return ZO_ERR_UNSUPPORTED;                                                      // :66,74 UnsupportedPixelType
    const uint32_t rows = src->rows, cols = src->cols;
    const size_t n = (size_t)rows * cols;
    std::vector<uint8_t> plane(n), res(n), merged(n * ch);
    for (int k = 0; k < ch; ++k) {                                                               // :199-229 split, filter, merge
        for (uint32_t r = 0; r < rows; ++r)
            for (uint32_t c = 0; c < cols; ++c) plane[(size_t)r * cols + c] = ((const uint8_t*)src->data)[((size_t)r * src->stride + c) * ch + k];
        zo::scalar_op(plane.data(), rows, cols, res.data(), (int64_t)radius, border, mode, param);
        for (size_t i = 0; i < n; ++i) merged[i * ch + k] = res[i];
    }
    for (uint32_t r = 0; r < rows; ++r)
        std::memcpy((uint8_t*)dst->data + (size_t)r * dst->stride * ch, merged.data() + (size_t)r * cols * ch, (size_t)cols * ch);
    return ZO_OK;
}
