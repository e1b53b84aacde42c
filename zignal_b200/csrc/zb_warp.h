// zb_warp.h -- internal declarations shared by zb_resize.cu / zb_warp.cu.
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>

#include "zb_internal.h"

namespace zb {
// Device copy of the reference's 1025-entry Lanczos3 LUT (interpolation.zig:256-267), built once per device on the host.
int lanczos_lut_device(const float** out, cudaStream_t s);
int resize_dispatch(const zb_image* src, zb_image* dst, int pixfmt, int method, float mb, float mc, cudaStream_t s);

// Image.rotateInto (transforms.zig:189-211): the geometry of one call, shared by the gather kernel (zb_warp.cu) and the tile kernel
// for Rgba(u8) / bilinear / .zero (zb_rotate_tile.cu; returns ZB_ERR_UNSUPPORTED when it does not apply).
struct RotParams {
    float cos_a, sin_a, cx, cy, rcx, rcy;
    int method, border;
    float mb, mc;
};
int rotate_tile_rgba8(const zb_image* src, unsigned long long spitch, zb_image* dst, unsigned long long dpitch, uint32_t n, const RotParams& rp,
                      cudaStream_t s);
extern std::atomic<int> g_tune_rotate_tile;   // zb_tune("rotate_tile", 0) forces the gather kernel (tests cross-check the two)

// Image.insert (transforms.zig:293-376): the geometry of one call, shared by the same-type kernel (zb_warp.cu) and the mixed-type
// kernel (zb_insert_mixed.cu).
struct InsertParams {
    float cx, cy, cos_a, sin_a, half_w, half_h, inv_w, inv_h, fcols1, frows1;
    int copy_rect, dst_top, dst_left;      // fast path: source pixel (r, c) lands on (dst_top + r, dst_left + c)
    int min_r, min_c, n_r, n_c;            // destination window this launch covers
    int src_rows, src_cols, method;
    float mb, mc;
    int blend;                             // ZB_BLEND_*; only Rgba(u8) pixels composite
};

// Fills `p` (fast copy path :305-323 or the rotated rectangle's bounding box :341-349).  Returns false when no destination pixel
// can be written.
static inline bool insert_plan(const zb_image* self, const zb_image* source, float rl, float rt, float rr, float rb, float angle, float cos_a,
                               float sin_a, int method, float mb, float mc, int blend, InsertParams& p) {
    memset(&p, 0, sizeof(p));
    const float frows = (float)source->rows, fcols = (float)source->cols;
    const float rect_width = rl >= rr ? 0.0f : rr - rl, rect_height = rt >= rb ? 0.0f : rb - rt;
    const float epsilon = 1e-6f;
    p.src_rows = (int)source->rows; p.src_cols = (int)source->cols;
    p.method = method; p.mb = mb; p.mc = mc;
    p.blend = blend;
    long long r0, r1, c0, c1;   // destination window [r0, r1) x [c0, c1)
    if (std::fabs(angle) < epsilon && std::fabs(rect_width - fcols) < epsilon && std::fabs(rect_height - frows) < epsilon) {   // :305-323
        p.copy_rect = 1;
        p.dst_top = (int)std::round(rt);
        p.dst_left = (int)std::round(rl);
        r0 = std::max<long long>(0, p.dst_top);
        r1 = std::min<long long>(self->rows, (long long)p.dst_top + source->rows);
        c0 = std::max<long long>(0, p.dst_left);
        c1 = std::min<long long>(self->cols, (long long)p.dst_left + source->cols);
    } else {
        p.cx = (rl + rr) * 0.5f; p.cy = (rt + rb) * 0.5f;
        p.cos_a = cos_a; p.sin_a = sin_a;
        p.inv_w = 1.0f / rect_width; p.inv_h = 1.0f / rect_height;
        p.half_w = rect_width * 0.5f; p.half_h = rect_height * 0.5f;
        p.fcols1 = fcols - 1; p.frows1 = frows - 1;
        const float abs_cos = std::fabs(cos_a), abs_sin = std::fabs(sin_a);
        const float bound_hw = p.half_w * abs_cos + p.half_h * abs_sin;   // exact bounding box of the rotated rectangle, :341-349
        const float bound_hh = p.half_w * abs_sin + p.half_h * abs_cos;
        auto to_u32 = [](float v) -> long long { return v <= 0 ? 0ll : (v >= 4294967040.0f ? 4294967295ll : (long long)v); };
        r0 = (p.cy - bound_hh < 0) ? 0 : to_u32(std::floor(p.cy - bound_hh));
        r1 = std::min<long long>(self->rows, to_u32(std::ceil(p.cy + bound_hh)) + 1);
        c0 = (p.cx - bound_hw < 0) ? 0 : to_u32(std::floor(p.cx - bound_hw));
        c1 = std::min<long long>(self->cols, to_u32(std::ceil(p.cx + bound_hw)) + 1);
    }
    if (r1 <= r0 || c1 <= c0) return false;
    p.min_r = (int)r0; p.min_c = (int)c0; p.n_r = (int)(r1 - r0); p.n_c = (int)(c1 - c0);
    return true;
}
}  // namespace zb
