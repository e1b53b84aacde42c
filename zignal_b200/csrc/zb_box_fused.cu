// zb_box_fused.cu -- Image.boxBlur / sharpen for 8-bit images without materialising the summed-area table in HBM.
//
// Reference: integral.zig:41-78 (f32 SAT: row prefix, then column accumulation), :85-90 (((D - left) - top) + corner),
// :194-269 / :325-422 (clipped-window box mean, 2*orig - blur), image.zig:635-648, :785-799.
//
// The SAT is f32 and its sums pass 2^24, so the SUMMATION ORDER is part of the result.  Two facts make a fused,
// parallel evaluation possible without changing a single bit:
//   * the row prefix P[r][c] = sum_{x<=c} src[r][x] of an 8-bit row is an integer < 2^24 (cols <= 65793): exact in f32,
//     so it may be computed in ANY order (integer warp scans);
//   * the column pass S[r][c] = fl(S[r-1][c] + P[r][c]) is a sequential chain in r, but chains of different columns are
//     independent, and a chain can be resumed from any row whose S is known.
// Three kernels (traffic: the source three times, the destination once, ~2% of it in side tables -- the 3-kernel SAT
// path moves 17x the image):
//   box_row_offsets : O[r][s] = exact integer row prefix at the first column of strip s          (one warp per row)
//   box_chain<CKPT> : one CTA per strip walks all rows: P tile (parallel) -> chains (sequential) and stores S at the
//                     row above each band's first input row
//   box_chain<EVAL> : one CTA per (strip, band): resumes the chains from its checkpoint, keeps the last 2r+2 SAT rows
//                     of its strip in a shared-memory ring and evaluates the box means / sharpen from it.
// Geometry: a "unit" is 4 bytes = one Rgba pixel (4 chains) or 4 gray pixels (1 chain each); a strip is 128 units =
// 512 element columns of which the inner OU = 128 - 2*MU units are outputs (MU covers radius+1 columns of halo).
#include "zb_device.cuh"
#include "zb_internal.h"

namespace zb {

namespace {

constexpr int SU = 128;          // units per strip row
constexpr int SE = 4 * SU;       // element columns per strip
constexpr int GR = 8;            // rows per block (one warp per row in the P phase)
constexpr int BF_THREADS = 256;
constexpr int BAND = 64;         // output rows per band

struct BoxParams {
    const uint8_t* src;
    uint8_t* dst;
    size_t src_pitch, dst_pitch;   // bytes
    int rows, cols;
    int row_units;                 // ceil(cols * CH / 4)
    int row_bytes;                 // cols * CH
    int radius, mu, ou, n_strips, n_bands;
    int ring;                      // rows in the SAT ring (power of two)
    const int* offs;               // [rows][n_strips][4]
    float* ckpt;                   // [n_bands][n_strips][SE]
};

// 4 bytes of row `row` at unit `u` (0 outside the row; the last unit of a gray row may be partial)
__device__ __forceinline__ uint32_t load_unit(const uint8_t* __restrict__ rowp, int u, int row_units, int row_bytes) {
    if (u < 0 || u >= row_units) return 0u;
    if (4 * u + 4 <= row_bytes) return *reinterpret_cast<const uint32_t*>(rowp + 4 * (size_t)u);
    uint32_t v = 0;
    for (int b = 0; 4 * u + b < row_bytes; ++b) v |= (uint32_t)rowp[4 * (size_t)u + b] << (8 * b);
    return v;
}

// P tile index with a 16-byte-chunk swizzle: the P phase stores 4 consecutive chunks per lane (64-byte stride between
// lanes), the chain phase reads consecutive floats; XOR-ing the chunk position with the low bits of its 128-byte group
// makes both conflict-free.
__device__ __forceinline__ int pt_chunk(int chunk) { return (chunk & ~7) | ((chunk ^ (chunk >> 3)) & 3) | (chunk & 4); }
__device__ __forceinline__ int pt_index(int e) { return pt_chunk(e >> 2) * 4 + (e & 3); }

// ---- 1. exact row prefix at every strip start --------------------------------------------------------------------------
template <int CH>
__global__ void __launch_bounds__(256) box_row_offsets(const BoxParams p, int* __restrict__ offs) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= p.rows) return;
    const uint8_t* rowp = p.src + (size_t)r * p.src_pitch;
    int carry[4] = {0, 0, 0, 0};
    int s = 0;                       // next strip whose start has not been passed
    int next_start = -p.mu;          // unit index of strip s's first unit
    int* out = offs + (size_t)r * p.n_strips * 4;
    while (s < p.n_strips && next_start <= 0) {   // strips that start at or left of the row start: prefix 0
        if (lane < 4) out[s * 4 + lane] = 0;
        ++s;
        next_start += p.ou;
    }
    for (int u0 = 0; u0 < p.row_units && s < p.n_strips; u0 += 32) {
        const uint32_t v = load_unit(rowp, u0 + lane, p.row_units, p.row_bytes);
        int c[4];
        if constexpr (CH == 4) {
            c[0] = v & 0xFF; c[1] = (v >> 8) & 0xFF; c[2] = (v >> 16) & 0xFF; c[3] = v >> 24;
        } else {
            c[0] = (int)__dp4a(v, 0x01010101u, 0u); c[1] = c[2] = c[3] = 0;
        }
        while (s < p.n_strips && next_start < u0 + 32) {   // a strip starts inside this window (next_start > u0 - 32 holds)
            const int pos = next_start - u0;               // exclusive prefix over lanes < pos
#pragma unroll
            for (int k = 0; k < (CH == 4 ? 4 : 1); ++k) {
                const int part = __reduce_add_sync(0xffffffffu, lane < pos ? c[k] : 0);
                if (lane == 0) out[s * 4 + k] = carry[k] + part;
            }
            ++s;
            next_start += p.ou;
        }
#pragma unroll
        for (int k = 0; k < (CH == 4 ? 4 : 1); ++k) carry[k] += __reduce_add_sync(0xffffffffu, c[k]);
    }
    while (s < p.n_strips) {   // strips starting beyond the row end (cannot happen for strips that own outputs; keep the table defined)
        if (lane < 4) out[s * 4 + lane] = carry[lane];
        ++s;
    }
}

// ---- 2./3. chains, checkpoints, evaluation --------------------------------------------------------------------------------
// MODE 0: checkpoint pass (grid = n_strips).  MODE 1: box blur, MODE 2: sharpen (grid = n_strips x n_bands).
template <int CH, int MODE>
__global__ void __launch_bounds__(BF_THREADS) box_chain(const BoxParams p) {
    extern __shared__ float smem_f[];
    float* pt = smem_f;                 // [GR][SE]  P tile (swizzled)
    float* ring = smem_f + GR * SE;     // [ring][SE] SAT rows (MODE != 0)
    const int t = threadIdx.x, lane = t & 31, wrow = t >> 5;
    const int strip = blockIdx.x;
    const int band = MODE == 0 ? 0 : blockIdx.y;
    const int unit0 = strip * p.ou - p.mu;           // first unit of the strip (may be negative)
    const int elem0 = unit0 * 4;
    const int y0 = MODE == 0 ? 0 : band * BAND;
    const int y1 = MODE == 0 ? p.rows : min(p.rows, y0 + BAND);
    const int ys = max(0, y0 - p.radius - 1);         // first chain row
    const int y_last = MODE == 0 ? p.rows - 1 : y1 - 1 + p.radius;   // last (possibly virtual) chain row
    const int rmask = p.ring - 1;

    // chains: thread t owns element columns t and t + 256 of the strip
    float s0 = 0.0f, s1 = 0.0f;
    if (MODE != 0 && ys > 0) {
        const float* ck = p.ckpt + ((size_t)band * p.n_strips + strip) * SE;
        s0 = ck[t];
        s1 = ck[t + 256];
    }

    for (int yb = ys; yb <= y_last; yb += GR) {
        // ---- P phase: warp `wrow` owns row yb + wrow; lane l owns units 4l .. 4l+3 of the strip
        {
            const int y = yb + wrow;
            if (y < p.rows) {
                const uint8_t* rowp = p.src + (size_t)y * p.src_pitch;
                uint32_t u[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) u[j] = load_unit(rowp, unit0 + 4 * lane + j, p.row_units, p.row_bytes);
                const int* off = p.offs + ((size_t)y * p.n_strips + strip) * 4;
                float* dstp = pt + wrow * SE;
                if constexpr (CH == 4) {
                    int loc[4][4];   // [unit][channel] inclusive local prefix
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        int run = 0;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            run += (int)((u[j] >> (8 * k)) & 0xFFu);
                            loc[j][k] = run;
                        }
                    }
                    int base[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        int incl = loc[3][k];
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            const int o = __shfl_up_sync(0xffffffffu, incl, d);
                            if (lane >= d) incl += o;
                        }
                        base[k] = off[k] + incl - loc[3][k];
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float4 f;
                        f.x = (float)(base[0] + loc[j][0]);
                        f.y = (float)(base[1] + loc[j][1]);
                        f.z = (float)(base[2] + loc[j][2]);
                        f.w = (float)(base[3] + loc[j][3]);
                        *reinterpret_cast<float4*>(dstp + pt_chunk(4 * lane + j) * 4) = f;
                    }
                } else {
                    int loc[16];
                    int run = 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            run += (int)((u[j] >> (8 * b)) & 0xFFu);
                            loc[4 * j + b] = run;
                        }
                    int incl = run;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int o = __shfl_up_sync(0xffffffffu, incl, d);
                        if (lane >= d) incl += o;
                    }
                    const int base = off[0] + incl - run;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float4 f;
                        f.x = (float)(base + loc[4 * j + 0]);
                        f.y = (float)(base + loc[4 * j + 1]);
                        f.z = (float)(base + loc[4 * j + 2]);
                        f.w = (float)(base + loc[4 * j + 3]);
                        *reinterpret_cast<float4*>(dstp + pt_chunk(4 * lane + j) * 4) = f;
                    }
                }
            }
        }
        __syncthreads();
        // ---- chain phase
        {
            const int i0 = pt_index(t), i1 = pt_index(t + 256);
#pragma unroll
            for (int g = 0; g < GR; ++g) {
                const int y = yb + g;
                if (y < p.rows) {
                    s0 = __fadd_rn(s0, pt[g * SE + i0]);
                    s1 = __fadd_rn(s1, pt[g * SE + i1]);
                    if constexpr (MODE == 0) {
                        const int q = y + p.radius + 2;          // S of this row resumes band q / BAND
                        if (q % BAND == 0 && q / BAND < p.n_bands) {
                            float* ck = p.ckpt + ((size_t)(q / BAND) * p.n_strips + strip) * SE;
                            ck[t] = s0;
                            ck[t + 256] = s1;
                        }
                    } else {
                        float* rr = ring + (size_t)(y & rmask) * SE;
                        rr[t] = s0;
                        rr[t + 256] = s1;
                    }
                }
            }
        }
        __syncthreads();
        // ---- evaluation of the output rows this block completed: yo = y - radius
        if constexpr (MODE != 0) {
            const int out_u0 = strip * p.ou;                          // first output unit of the strip
            const int out_un = min(p.ou, p.row_units - out_u0);      // output units in this strip
            for (int idx = t; idx < GR * out_un; idx += BF_THREADS) {
                const int g = idx / out_un, uu = idx - g * out_un;
                const int yo = yb + g - p.radius;
                if (yo < y0 || yo >= y1) continue;
                const int r1 = max(yo - p.radius, 0), r2 = min(yo + p.radius, p.rows - 1);
                const float* row_d = ring + (size_t)(r2 & rmask) * SE;
                const float* row_t = ring + (size_t)((r1 - 1) & rmask) * SE;
                const int unit = out_u0 + uu;
                uint32_t orig = 0;
                if constexpr (MODE == 2) orig = load_unit(p.src + (size_t)yo * p.src_pitch, unit, p.row_units, p.row_bytes);
                uint32_t packed = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c = CH == 4 ? unit : 4 * unit + j;      // pixel column
                    if (CH == 1 && c >= p.cols) break;
                    const int c1 = max(c - p.radius, 0), c2 = min(c + p.radius, p.cols - 1);
                    const int ed = (CH == 4 ? 4 * c2 + j : c2) - elem0;           // local element of (.., c2)
                    const int el = (CH == 4 ? 4 * (c1 - 1) + j : c1 - 1) - elem0;  // local element of (.., c1 - 1)
                    const float area = (float)((r2 - r1 + 1) * (c2 - c1 + 1));
                    const float D = row_d[ed];
                    const float left = c1 > 0 ? row_d[el] : 0.0f;
                    const float top = r1 > 0 ? row_t[ed] : 0.0f;
                    const float corner = (r1 > 0 && c1 > 0) ? row_t[el] : 0.0f;
                    const float s = __fadd_rn(__fsub_rn(__fsub_rn(D, left), top), corner);   // integral.zig:86-89
                    float val = __fdiv_rn(s, area);
                    if constexpr (MODE == 2) val = __fsub_rn(__fmul_rn(2.0f, (float)((orig >> (8 * j)) & 0xFFu)), val);   // integral.zig:357
                    packed |= (uint32_t)clamp_u8_from_float(val) << (8 * j);
                }
                uint8_t* drow = p.dst + (size_t)yo * p.dst_pitch;
                if (4 * unit + 4 <= p.row_bytes) {
                    *reinterpret_cast<uint32_t*>(drow + 4 * (size_t)unit) = packed;
                } else {
                    for (int b = 0; 4 * unit + b < p.row_bytes; ++b) drow[4 * (size_t)unit + b] = (uint8_t)(packed >> (8 * b));
                }
            }
            // no barrier needed here: the ring holds 2*GR more rows than the evaluation reaches back
        }
    }
}

template <int CH>
int launch_all(const BoxParams& p, bool sharpen, int* offs, cudaStream_t s) {
    box_row_offsets<CH><<<div_up(p.rows, 8), 256, 0, s>>>(p, offs);
    ZB_LAUNCHED();
    const int smem_ck = GR * SE * (int)sizeof(float);
    const int smem_ev = (GR + p.ring) * SE * (int)sizeof(float);
    if (p.n_bands > 1) {
        box_chain<CH, 0><<<p.n_strips, BF_THREADS, smem_ck, s>>>(p);
        ZB_LAUNCHED();
    }
    dim3 grid(p.n_strips, p.n_bands);
    if (sharpen) {
        ZB_CUDA(cudaFuncSetAttribute(box_chain<CH, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_ev));
        box_chain<CH, 2><<<grid, BF_THREADS, smem_ev, s>>>(p);
    } else {
        ZB_CUDA(cudaFuncSetAttribute(box_chain<CH, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_ev));
        box_chain<CH, 1><<<grid, BF_THREADS, smem_ev, s>>>(p);
    }
    ZB_LAUNCHED();
    return ZB_OK;
}

}  // namespace

// Returns ZB_ERR_UNSUPPORTED outside the fused envelope (the caller then builds the SAT in HBM).
int box_fused_u8(const zb_image* src, zb_image* dst, int channels, uint32_t radius, bool sharpen, cudaStream_t s) {
    if (channels != 1 && channels != 4) return ZB_ERR_UNSUPPORTED;
    if (radius == 0 || radius > 15) return ZB_ERR_UNSUPPORTED;                 // ring <= 64 rows (128 KB)
    if (src->data == dst->data) return ZB_ERR_UNSUPPORTED;                     // bands would race with in-place rows
    if ((uint64_t)src->cols * 255u >= (1u << 24)) return ZB_ERR_UNSUPPORTED;   // row prefixes must be exact in f32
    if (src->rows > (1u << 30) || src->cols > (1u << 28)) return ZB_ERR_UNSUPPORTED;
    const size_t sp = (size_t)src->stride * channels, dp = (size_t)dst->stride * channels;
    if (((uintptr_t)src->data & 3u) || ((uintptr_t)dst->data & 3u) || (sp & 3u) || (dp & 3u)) return ZB_ERR_UNSUPPORTED;
    BoxParams p;
    p.src = (const uint8_t*)src->data;
    p.dst = (uint8_t*)dst->data;
    p.src_pitch = sp;
    p.dst_pitch = dp;
    p.rows = (int)src->rows;
    p.cols = (int)src->cols;
    p.row_bytes = p.cols * channels;
    p.row_units = (p.row_bytes + 3) / 4;
    p.radius = (int)radius;
    p.mu = channels == 4 ? p.radius + 1 : (p.radius + 1 + 3) / 4;
    p.ou = SU - 2 * p.mu;
    p.n_strips = (p.row_units + p.ou - 1) / p.ou;
    p.n_bands = (p.rows + BAND - 1) / BAND;
    int ring = 32;
    while (ring < 2 * p.radius + 2 + 2 * GR) ring <<= 1;
    p.ring = ring;
    Scratch offs, ckpt;
    int rc = offs.alloc((size_t)p.rows * p.n_strips * 4 * sizeof(int), s);
    if (rc) return rc;
    if ((rc = ckpt.alloc((size_t)p.n_bands * p.n_strips * SE * sizeof(float), s))) return rc;
    p.offs = offs.as<int>();
    p.ckpt = ckpt.as<float>();
    t_last_kernel = sharpen ? "box_fused_sharpen" : "box_fused_blur";
    return channels == 4 ? launch_all<4>(p, sharpen, offs.as<int>(), s) : launch_all<1>(p, sharpen, offs.as<int>(), s);
}

}  // namespace zb
