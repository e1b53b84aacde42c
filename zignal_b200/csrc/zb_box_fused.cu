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
//   box_checkpoints : the only sequential pass, kept to the bare chain: one WARP per (strip, channel) walks all rows with
//                     no block-level synchronisation (unit loads run ahead, one warp scan per row) and stores S at the
//                     row above each band's first input row
//   box_chain       : one CTA per (strip, band): resumes the chains from its checkpoint, keeps the last 2r+2 SAT rows
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
    const int* offs;               // [rows][n_strips][4]   row prefix at each strip's first unit
    const int* offs32;             // [rows][n_win][4]      row prefix at every 32nd unit (checkpoint pass)
    int n_win;                     // ceil(row_units / 32)
    int ck_pitch;                  // floats per checkpoint row: n_win * 128
    float* ckpt;                   // [n_bands][ck_pitch]   S at row (band * BAND - radius - 2), indexed by global element column
};

// 4 bytes of a row at unit `u` (0 outside the row; the last unit of a gray row may be partial, an Rgba unit never is)
template <int CH>
__device__ __forceinline__ uint32_t load_unit(const uint8_t* __restrict__ rowp, int u, int row_units, int row_bytes) {
    if ((unsigned)u >= (unsigned)row_units) return 0u;
    if (CH == 4 || 4 * u + 4 <= row_bytes) return __ldg(reinterpret_cast<const uint32_t*>(rowp) + u);
    uint32_t v = 0;
    for (int b = 0; 4 * u + b < row_bytes; ++b) v |= (uint32_t)rowp[4 * (size_t)u + b] << (8 * b);
    return v;
}

// P tile index with a 16-byte-chunk swizzle: the P phase stores 4 consecutive chunks per lane (64-byte stride between
// lanes), the chain phase reads consecutive floats; XOR-ing the chunk position with the low bits of its 128-byte group
// makes both conflict-free.
__device__ __forceinline__ int pt_chunk(int chunk) { return (chunk & ~7) | ((chunk ^ (chunk >> 3)) & 3) | (chunk & 4); }
__device__ __forceinline__ int pt_index(int e) { return pt_chunk(e >> 2) * 4 + (e & 3); }

// ---- 1. exact row prefixes: at every 32nd unit (for the checkpoint pass) and at every strip start ------------------------------
template <int CH>
__global__ void __launch_bounds__(256) box_row_offsets(const BoxParams p, int* __restrict__ offs, int* __restrict__ offs32) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= p.rows) return;
    const uint8_t* rowp = p.src + (size_t)r * p.src_pitch;
    int4 carry = make_int4(0, 0, 0, 0);
    int s = 0;                       // next strip whose start has not been passed
    int next_start = -p.mu;          // unit index of strip s's first unit
    int4* out = reinterpret_cast<int4*>(offs) + (size_t)r * p.n_strips;
    int4* out32 = reinterpret_cast<int4*>(offs32) + (size_t)r * p.n_win;
    while (s < p.n_strips && next_start <= 0) {   // strips that start at or left of the row start: prefix 0
        if (lane == 0) out[s] = carry;
        ++s;
        next_start += p.ou;
    }
    for (int w = 0; w < p.n_win; ++w) {
        const int u0 = w * 32;
        const uint32_t v = load_unit<CH>(rowp, u0 + lane, p.row_units, p.row_bytes);
        int c[4];
        if constexpr (CH == 4) {
            c[0] = v & 0xFF; c[1] = (v >> 8) & 0xFF; c[2] = (v >> 16) & 0xFF; c[3] = v >> 24;
        } else {
            c[0] = (int)__dp4a(v, 0x01010101u, 0u); c[1] = c[2] = c[3] = 0;
        }
        if (lane == 0) out32[w] = carry;
        if (s < p.n_strips && next_start < u0 + 32) {   // a strip starts inside this window (at most one: OU >= 96)
            const int pos = next_start - u0;              // exclusive prefix over lanes < pos
            int4 part = carry;
            part.x += __reduce_add_sync(0xffffffffu, lane < pos ? c[0] : 0);
            if constexpr (CH == 4) {
                part.y += __reduce_add_sync(0xffffffffu, lane < pos ? c[1] : 0);
                part.z += __reduce_add_sync(0xffffffffu, lane < pos ? c[2] : 0);
                part.w += __reduce_add_sync(0xffffffffu, lane < pos ? c[3] : 0);
            }
            if (lane == 0) out[s] = part;
            ++s;
            next_start += p.ou;
        }
        carry.x += __reduce_add_sync(0xffffffffu, c[0]);
        if constexpr (CH == 4) {
            carry.y += __reduce_add_sync(0xffffffffu, c[1]);
            carry.z += __reduce_add_sync(0xffffffffu, c[2]);
            carry.w += __reduce_add_sync(0xffffffffu, c[3]);
        }
    }
}

// exclusive prefix over the lanes of a warp
__device__ __forceinline__ int warp_exclusive(int v, int lane) {
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    return incl - v;
}

// ---- 2. checkpoints: S at row (j * BAND - radius - 2) for every band j >= 1 ------------------------------------------------
// The only sequential pass, so it carries nothing but the chain: one WARP per (32-unit window, channel) for Rgba -- lane = one
// pixel, one chain -- and per window for gray (lane = 4 pixels, 4 chains); no block-level synchronisation.  Rows are taken
// PF at a time: the loads of the next PF rows are in flight while the PF independent warp scans of the current ones run.
template <int CH>
__global__ void __launch_bounds__(128) box_checkpoints(const BoxParams p) {
    constexpr int PF = 8;
    constexpr int NC = CH == 4 ? 1 : 4;   // chains per lane
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int win = CH == 4 ? gw >> 2 : gw;
    const int k = CH == 4 ? gw & 3 : 0;
    if (win >= p.n_win) return;
    const int unit = win * 32 + lane;
    const bool in_row = unit < p.row_units;
    const bool full = CH == 4 || 4 * unit + 4 <= p.row_bytes;
    const uint8_t* colp = p.src + 4 * (size_t)unit;
    const int* offp = p.offs32 + (size_t)win * 4 + k;
    const size_t off_row = (size_t)p.n_win * 4;
    float S[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) S[i] = 0.0f;
    const int y_end = (p.n_bands - 1) * BAND - p.radius - 2;   // last checkpoint row

    uint32_t u[PF], un[PF];
    int off[PF], offn[PF];
    auto fetch = [&](int y0) {
#pragma unroll
        for (int i = 0; i < PF; ++i) {
            const int y = min(y0 + i, p.rows - 1);
            const uint8_t* q = colp + (size_t)y * p.src_pitch;
            uint32_t v = 0;
            if (in_row) {
                if (full) v = __ldg(reinterpret_cast<const uint32_t*>(q));
                else for (int b = 0; 4 * unit + b < p.row_bytes; ++b) v |= (uint32_t)q[b] << (8 * b);
            }
            un[i] = v;
            offn[i] = __ldg(offp + (size_t)y * off_row);
        }
    };
    fetch(0);
    for (int yb = 0; yb <= y_end; yb += PF) {
#pragma unroll
        for (int i = 0; i < PF; ++i) { u[i] = un[i]; off[i] = offn[i]; }
        if (yb + PF <= y_end) fetch(yb + PF);
        float pv[PF][NC];
#pragma unroll
        for (int i = 0; i < PF; ++i) {   // PF independent scans
            int loc[NC];
            int run = 0;
            if constexpr (CH == 4) {
                run = (int)((u[i] >> (8 * k)) & 0xFFu);
                loc[0] = run;
            } else {
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    run += (int)((u[i] >> (8 * b)) & 0xFFu);
                    loc[b] = run;
                }
            }
            const int base = off[i] + warp_exclusive(run, lane);
#pragma unroll
            for (int c = 0; c < NC; ++c) pv[i][c] = (float)(base + loc[c]);
        }
#pragma unroll
        for (int i = 0; i < PF; ++i) {   // the chains, row by row
            const int y = yb + i;
            if (y <= y_end) {
#pragma unroll
                for (int c = 0; c < NC; ++c) S[c] = __fadd_rn(S[c], pv[i][c]);
                const int q = y + p.radius + 2;
                if (q % BAND == 0) {
                    float* ck = p.ckpt + (size_t)(q / BAND) * p.ck_pitch + (size_t)unit * 4;
#pragma unroll
                    for (int c = 0; c < NC; ++c) ck[CH == 4 ? k : c] = S[c];
                }
            }
        }
    }
}

// ---- 3. chains from the checkpoint, SAT ring, evaluation ---------------------------------------------------------------------
// correctly rounded s / area from a correctly rounded reciprocal: two Markstein corrections (q1 is faithful, q2 exact)
__device__ __forceinline__ float div_exact(float s, float area, float rcp) {
    const float q0 = __fmul_rn(s, rcp);
    const float q1 = __fmaf_rn(__fmaf_rn(-q0, area, s), rcp, q0);
    return __fmaf_rn(__fmaf_rn(-q1, area, s), rcp, q1);
}
// meta.clamp(u8, f32) for the values this path produces: v = m / area or 2*orig - m / area with integer m, so v is either an
// exact tie k + 0.5 or at least 1/(2*961) away from one -- trunc(v + 0.5) equals round-half-away for v >= 0, and v < 0 clamps to 0.
__device__ __forceinline__ uint32_t clamp_u8_fast(float v) { return (uint32_t)min(max(__float2int_rz(__fadd_rn(v, 0.5f)), 0), 255); }

// MODE 1: box blur, MODE 2: sharpen (grid = n_strips x n_bands).
template <int CH, int MODE>
__global__ void __launch_bounds__(BF_THREADS) box_chain(const BoxParams p) {
    extern __shared__ float smem_f[];
    float* pt = smem_f;                 // [GR][SE]  P tile (swizzled)
    float* ring = smem_f + GR * SE;     // [ring][SE] SAT rows (MODE != 0)
    const int t = threadIdx.x, lane = t & 31, wrow = t >> 5;
    const int strip = blockIdx.x;
    const int band = blockIdx.y;
    const int unit0 = strip * p.ou - p.mu;           // first unit of the strip (may be negative)
    const int elem0 = unit0 * 4;
    const int y0 = band * BAND;
    const int y1 = min(p.rows, y0 + BAND);
    const int ys = max(0, y0 - p.radius - 1);         // first chain row
    const int y_last = y1 - 1 + p.radius;             // last (possibly virtual) chain row
    const int rmask = p.ring - 1;

    // chains: thread t owns element columns t and t + 256 of the strip
    float s0 = 0.0f, s1 = 0.0f;
    if (ys > 0) {   // resume from the checkpoint (columns outside the image are never read by an output)
        const float* ck = p.ckpt + (size_t)band * p.ck_pitch;
        const int e0 = elem0 + t, e1 = elem0 + t + 256;
        s0 = (e0 >= 0 && e0 < p.ck_pitch) ? ck[e0] : 0.0f;
        s1 = (e1 >= 0 && e1 < p.ck_pitch) ? ck[e1] : 0.0f;
    }

    // the loads of a block are issued one block ahead of their use
    uint32_t un[4] = {0u, 0u, 0u, 0u};
    int4 offn = make_int4(0, 0, 0, 0);
    auto fetch = [&](int y) {
        if (y < p.rows) {
            const uint8_t* rowp = p.src + (size_t)y * p.src_pitch;
#pragma unroll
            for (int j = 0; j < 4; ++j) un[j] = load_unit<CH>(rowp, unit0 + 4 * lane + j, p.row_units, p.row_bytes);
            offn = __ldg(reinterpret_cast<const int4*>(p.offs + ((size_t)y * p.n_strips + strip) * 4));
        }
    };
    fetch(ys + wrow);

    for (int yb = ys; yb <= y_last; yb += GR) {
        // ---- P phase: warp `wrow` owns row yb + wrow; lane l owns units 4l .. 4l+3 of the strip
        {
            const int y = yb + wrow;
            const uint32_t u[4] = {un[0], un[1], un[2], un[3]};
            const int off[4] = {offn.x, offn.y, offn.z, offn.w};
            if (yb + GR <= y_last) fetch(y + GR);
            if (y < p.rows) {
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
                    for (int k = 0; k < 4; ++k) base[k] = off[k] + warp_exclusive(loc[3][k], lane);
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
                    const int base = off[0] + warp_exclusive(run, lane);
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
                    float* rr = ring + (y & rmask) * SE;
                    rr[t] = s0;
                    rr[t + 256] = s1;
                }
            }
        }
        __syncthreads();
        // ---- evaluation of the output rows this block completed: yo = y - radius
        {
            const int out_u0 = strip * p.ou;                          // first output unit of the strip
            const int out_un = min(p.ou, p.row_units - out_u0);      // output units in this strip
            const int uu = t & (SU - 1);
            for (int g = t >> 7; g < GR; g += BF_THREADS / SU) {
                const int yo = yb + g - p.radius;
                if (uu >= out_un || yo < y0 || yo >= y1) continue;
                const int r1 = max(yo - p.radius, 0), r2 = min(yo + p.radius, p.rows - 1);
                const float* row_d = ring + (r2 & rmask) * SE;
                const float* row_t = ring + ((r1 - 1) & rmask) * SE;
                const int unit = out_u0 + uu;
                uint32_t orig = 0;
                if constexpr (MODE == 2) orig = load_unit<CH>(p.src + (size_t)yo * p.src_pitch, unit, p.row_units, p.row_bytes);
                uint32_t packed = 0;
                if constexpr (CH == 4) {
                    const int c = unit;
                    const int c1 = max(c - p.radius, 0), c2 = min(c + p.radius, p.cols - 1);
                    const float area = (float)((r2 - r1 + 1) * (c2 - c1 + 1));
                    const float rcp = __frcp_rn(area);
                    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                    const float4 D = *reinterpret_cast<const float4*>(row_d + (4 * c2 - elem0));
                    const float4 Lf = c1 > 0 ? *reinterpret_cast<const float4*>(row_d + (4 * (c1 - 1) - elem0)) : z;
                    const float4 T = r1 > 0 ? *reinterpret_cast<const float4*>(row_t + (4 * c2 - elem0)) : z;
                    const float4 Cn = (r1 > 0 && c1 > 0) ? *reinterpret_cast<const float4*>(row_t + (4 * (c1 - 1) - elem0)) : z;
                    const float dv[4] = {D.x, D.y, D.z, D.w}, lv[4] = {Lf.x, Lf.y, Lf.z, Lf.w};
                    const float tv[4] = {T.x, T.y, T.z, T.w}, cv[4] = {Cn.x, Cn.y, Cn.z, Cn.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float sum = __fadd_rn(__fsub_rn(__fsub_rn(dv[j], lv[j]), tv[j]), cv[j]);   // integral.zig:86-89
                        float val = div_exact(sum, area, rcp);
                        if constexpr (MODE == 2) val = __fsub_rn(__fmul_rn(2.0f, (float)((orig >> (8 * j)) & 0xFFu)), val);   // integral.zig:357
                        packed |= clamp_u8_fast(val) << (8 * j);
                    }
                } else {
                    const float rows_f = (float)(r2 - r1 + 1);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int c = 4 * unit + j;                   // pixel column
                        if (c >= p.cols) break;
                        const int c1 = max(c - p.radius, 0), c2 = min(c + p.radius, p.cols - 1);
                        const int ed = c2 - elem0, el = c1 - 1 - elem0;
                        const float area = __fmul_rn(rows_f, (float)(c2 - c1 + 1));
                        const float D = row_d[ed];
                        const float left = c1 > 0 ? row_d[el] : 0.0f;
                        const float top = r1 > 0 ? row_t[ed] : 0.0f;
                        const float corner = (r1 > 0 && c1 > 0) ? row_t[el] : 0.0f;
                        const float sum = __fadd_rn(__fsub_rn(__fsub_rn(D, left), top), corner);
                        float val = div_exact(sum, area, __frcp_rn(area));
                        if constexpr (MODE == 2) val = __fsub_rn(__fmul_rn(2.0f, (float)((orig >> (8 * j)) & 0xFFu)), val);
                        packed |= clamp_u8_fast(val) << (8 * j);
                    }
                }
                uint8_t* drow = p.dst + (size_t)yo * p.dst_pitch;
                if (4 * unit + 4 <= p.row_bytes) {
                    *reinterpret_cast<uint32_t*>(drow + 4 * (size_t)unit) = packed;
                } else {
                    for (int b = 0; 4 * unit + b < p.row_bytes; ++b) drow[4 * (size_t)unit + b] = (uint8_t)(packed >> (8 * b));
                }
            }
            // no barrier needed here: a thread only reaches the next chain phase (the next ring write) through the barrier after
            // the next P phase, which every thread enters after finishing this loop
        }
    }
}

template <int CH>
int launch_all(const BoxParams& p, bool sharpen, int* offs, int* offs32, cudaStream_t s) {
    box_row_offsets<CH><<<div_up(p.rows, 8), 256, 0, s>>>(p, offs, offs32);
    ZB_LAUNCHED();
    const int smem_ev = (GR + p.ring) * SE * (int)sizeof(float);
    if (p.n_bands > 1) {
        const int warps = CH == 4 ? 4 * p.n_win : p.n_win;
        box_checkpoints<CH><<<div_up(warps, 4), 128, 0, s>>>(p);
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
    if (radius == 0 || radius > 15) return ZB_ERR_UNSUPPORTED;                 // ring <= 64 rows (128 KB), margins <= 16 units
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
    int ring = 16;
    while (ring < 2 * p.radius + 2 + GR) ring <<= 1;
    p.ring = ring;
    p.n_win = (p.row_units + 31) / 32;
    p.ck_pitch = p.n_win * 128;
    Scratch offs, offs32, ckpt;
    int rc = offs.alloc((size_t)p.rows * p.n_strips * 4 * sizeof(int), s);
    if (rc) return rc;
    if ((rc = offs32.alloc((size_t)p.rows * p.n_win * 4 * sizeof(int), s))) return rc;
    if ((rc = ckpt.alloc((size_t)p.n_bands * p.ck_pitch * sizeof(float), s))) return rc;
    p.offs = offs.as<int>();
    p.offs32 = offs32.as<int>();
    p.ckpt = ckpt.as<float>();
    t_last_kernel = sharpen ? "box_fused_sharpen" : "box_fused_blur";
    return channels == 4 ? launch_all<4>(p, sharpen, offs.as<int>(), offs32.as<int>(), s)
                         : launch_all<1>(p, sharpen, offs.as<int>(), offs32.as<int>(), s);
}

}  // namespace zb
