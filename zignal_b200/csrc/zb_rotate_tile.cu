// zb_rotate_tile.cu -- Image.rotateInto for Rgba(u8), bilinear, `.zero` border (BASELINE config 4) from shared-memory source tiles.
// Reference: transforms.zig:189-211 (the inverse-mapped loop; every coordinate is the reference's own unfused f32 sequence),
// interpolation.zig:313-368 (bilinear: out-of-range neighbours are zero pixels under .zero, integer lerp with
// fx = @round(frac * 256) and +32768 rounding).
//
// The gather kernel of zb_warp.cu spends ~100 instructions per interior pixel, most of them on 64-bit addressing, edge tests and
// four scattered global loads.  Here a CTA owns a 64 x 64 destination tile:
//   * the source coordinates are monotone in the column and in the row (every step of the formula is a monotone f32 operation), so
//     the four corners of the tile bound the source footprint exactly; TMA loads that bounding box (<= 92 x 92 pixels for any
//     angle) into shared memory as 16-row boxes of a {cols, rows, frames} tensor map.  TMA fills what lies outside the image with
//     zeros, which IS the `.zero` border: no edge test per sample, and a tile whose footprint misses the image is only zero stores;
//   * floor(x) and @round(256 * frac(x)) come out of ONE magic-number add per axis: round-half-up(256 x) = floor(256 x + 1/2)
//     = 256 * floor(x) + fx with fx = 256 carried into the integer part (left + 1 with weight 0 is the same sample as left with
//     weight 256), computed as fma.rm(x, 256, 0.5) (rounding down never crosses an integer) + 1.5 * 2^23 in round-down mode: the
//     low byte of the sum's bit pattern is the weight, the bits above it the tile-relative column.  No conversion instructions;
//   * a warp covers an 8 x 4 pixel patch per step, so its shared-memory footprint is a ~9 x 9 pixel patch for any angle; the tile
//     pitch of 104 words (== 8 mod 32) keeps those four rows in distinct banks; stores are 32-byte row segments.
// The terms of the coordinate that depend on the row only (sin * dy, cos * dy) are tabulated once per tile.
#include <cuda.h>

#include "zb_internal.h"
#include "zb_sample.cuh"
#include "zb_tma.cuh"
#include "zb_warp.h"

namespace zb {

namespace {

constexpr int RT_T = 64;                 // destination tile edge
constexpr int RT_P = 104;                // shared tile pitch in pixels (TMA box width)
constexpr int RT_BOXH = 16;              // rows per TMA box
constexpr int RT_MAXH = 96;              // >= 63 * sqrt(2) + 3, multiple of RT_BOXH
constexpr int RT_THREADS = 256;
constexpr uint32_t RT_TILE_BYTES = RT_MAXH * RT_P * 4;
constexpr uint32_t RT_SMEM = RT_TILE_BYTES + RT_T * 8 + 16;
constexpr float RT_MAGIC = 12582912.0f;  // 1.5 * 2^23
constexpr int RT_MAGIC_BITS = 0x4B400000;

// 256 * floor(x) + @round(256 * (x - floor(x))) as an integer (see the header); exact for |x| < 16384
__device__ __forceinline__ int q8_coord(float x) { return __float_as_int(__fadd_rd(__fmaf_rd(x, 256.0f, 0.5f), RT_MAGIC)) - RT_MAGIC_BITS; }

struct TileParams {
    uint32_t* dst;
    unsigned long long dst_stride, dst_image_pitch;   // pixels
    int dst_rows, dst_cols, src_rows, src_cols;
    float cos_a, sin_a, cx, cy, rcx, rcy;
};

__global__ void __launch_bounds__(RT_THREADS, 4) rotate_tile_rgba8_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TileParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t tile = smem_u32(smem_raw);
    float2* rowterm = reinterpret_cast<float2*>(smem_raw + RT_TILE_BYTES);
    const uint32_t bar = tile + RT_TILE_BYTES + RT_T * 8;

    const int c0 = blockIdx.x * RT_T, r0 = blockIdx.y * RT_T;
    const int c1 = min(c0 + RT_T - 1, p.dst_cols - 1), r1 = min(r0 + RT_T - 1, p.dst_rows - 1);
    // source footprint of the tile: the coordinates of its four corners (transforms.zig:199-209)
    int min_x, max_x, min_y, max_y;
    {
        const float dxa = (float)c0 - p.rcx, dxb = (float)c1 - p.rcx, dya = (float)r0 - p.rcy, dyb = (float)r1 - p.rcy;
        const float ca = p.cos_a * dxa, cb = p.cos_a * dxb, sa = p.sin_a * dxa, sb = p.sin_a * dxb;
        const float sya = p.sin_a * dya, syb = p.sin_a * dyb, cya = p.cos_a * dya, cyb = p.cos_a * dyb;
        const int x0 = q8_coord((ca - sya) + p.cx) >> 8, x1 = q8_coord((ca - syb) + p.cx) >> 8, x2 = q8_coord((cb - sya) + p.cx) >> 8,
                  x3 = q8_coord((cb - syb) + p.cx) >> 8;
        const int y0 = q8_coord((sa + cya) + p.cy) >> 8, y1 = q8_coord((sa + cyb) + p.cy) >> 8, y2 = q8_coord((sb + cya) + p.cy) >> 8,
                  y3 = q8_coord((sb + cyb) + p.cy) >> 8;
        min_x = min(min(x0, x1), min(x2, x3)); max_x = max(max(x0, x1), max(x2, x3));
        min_y = min(min(y0, y1), min(y2, y3)); max_y = max(max(y0, y1), max(y2, y3));
    }
    const int pc = (int)(threadIdx.x >> 5) * 8 + (int)(threadIdx.x & 7u);   // column of the tile this thread produces
    const int pr = (int)((threadIdx.x >> 3) & 3u);                          // its row within every 4-row step
    const int c = c0 + pc;
    uint32_t* out = p.dst + (size_t)blockIdx.z * p.dst_image_pitch + (size_t)(r0 + pr) * p.dst_stride + (size_t)c;
    const size_t step = 4 * p.dst_stride;

    if (max_x + 1 < 0 || min_x >= p.src_cols || max_y + 1 < 0 || min_y >= p.src_rows) {   // every neighbour of every sample is outside
        if (c <= c1)
            for (int r = r0 + pr; r <= r1; r += 4, out += step) *out = 0u;
        return;
    }
    const int n_boxes = (max_y - min_y + 2 + RT_BOXH - 1) / RT_BOXH;   // rows min_y .. max_y + 1
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    if (threadIdx.x < RT_T) {
        const float dy = (float)(r0 + (int)threadIdx.x) - p.rcy;
        rowterm[threadIdx.x] = make_float2(p.sin_a * dy, p.cos_a * dy);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, (uint32_t)n_boxes * RT_BOXH * RT_P * 4);
        for (int b = 0; b < n_boxes; ++b) tma_load_3d(tile + (uint32_t)b * RT_BOXH * RT_P * 4, &tmap, min_x, min_y + b * RT_BOXH, (int)blockIdx.z, bar);
    }
    if (c > c1) return;
    const float dx = (float)c - p.rcx;
    const float cos_dx = p.cos_a * dx, sin_dx = p.sin_a * dx;
    const int bias_x = min_x << 8, bias_y = min_y << 8;
    while (!mbar_try_wait(bar, 0)) {}
    const int nj = (r1 - r0 - pr) / 4 + 1;   // rows r0 + pr + 4 j <= r1
#pragma unroll 4
    for (int j = 0; j < nj; ++j) {
        const float2 rt = rowterm[4 * j + pr];
        const float src_x = (cos_dx - rt.x) + p.cx;
        const float src_y = (sin_dx + rt.y) + p.cy;
        const unsigned mx = (unsigned)(q8_coord(src_x) - bias_x), my = (unsigned)(q8_coord(src_y) - bias_y);
        const uint32_t a = tile + (my >> 8) * (RT_P * 4) + ((mx >> 8) << 2);
        *out = bilerp_rgba8(lds32(a), lds32(a + 4), lds32(a + RT_P * 4), lds32(a + RT_P * 4 + 4), mx & 255u, my & 255u);
        out += step;
    }
}

}  // namespace

// Returns ZB_ERR_UNSUPPORTED when the tile kernel does not apply (the caller then takes the gather kernel).
int rotate_tile_rgba8(const zb_image* src, unsigned long long spitch, zb_image* dst, unsigned long long dpitch, uint32_t n, const RotParams& rp,
                      cudaStream_t s) {
    if (rp.method != ZB_INTERP_BILINEAR || rp.border != ZB_BORDER_ZERO) return ZB_ERR_UNSUPPORTED;
    if (src->rows == 0 || src->cols == 0 || src->rows > 8192 || src->cols > 8192 || dst->rows > 16384 || dst->cols > 16384) return ZB_ERR_UNSUPPORTED;
    if (n > 65535 || ((uintptr_t)src->data & 15) || (src->stride & 3) || (n > 1 && (spitch & 3)) || ((uintptr_t)dst->data & 3)) return ZB_ERR_UNSUPPORTED;
    EncodeTiledFn encode = encode_tiled_fn();
    if (!encode) return ZB_ERR_UNSUPPORTED;
    CUtensorMap tmap;
    const cuuint64_t gdim[3] = {(cuuint64_t)src->cols, (cuuint64_t)src->rows, (cuuint64_t)n};
    const cuuint64_t frame_bytes = n > 1 ? (cuuint64_t)spitch * 4 : (cuuint64_t)src->rows * src->stride * 4;
    const cuuint64_t gstr[2] = {(cuuint64_t)src->stride * 4, frame_bytes};
    const cuuint32_t box[3] = {(cuuint32_t)RT_P, (cuuint32_t)RT_BOXH, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, src->data, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return ZB_ERR_UNSUPPORTED;
    TileParams p;
    p.dst = (uint32_t*)dst->data;
    p.dst_stride = dst->stride;
    p.dst_image_pitch = dpitch;
    p.dst_rows = (int)dst->rows; p.dst_cols = (int)dst->cols;
    p.src_rows = (int)src->rows; p.src_cols = (int)src->cols;
    p.cos_a = rp.cos_a; p.sin_a = rp.sin_a; p.cx = rp.cx; p.cy = rp.cy; p.rcx = rp.rcx; p.rcy = rp.rcy;
    static bool attr_set = false;   // (per process; the attribute is per function, every device of the process gets it on first use)
    if (!attr_set) {
        ZB_CUDA(cudaFuncSetAttribute(rotate_tile_rgba8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RT_SMEM));
        attr_set = true;
    }
    const dim3 grid(div_up(dst->cols, RT_T), div_up(dst->rows, RT_T), n);
    rotate_tile_rgba8_kernel<<<grid, RT_THREADS, RT_SMEM, s>>>(tmap, p);
    t_last_kernel = "rotate_tile_rgba8";
    ZB_LAUNCHED();
    return ZB_OK;
}

}  // namespace zb
