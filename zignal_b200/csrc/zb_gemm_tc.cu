// zb_gemm_tc.cu -- the dense contraction behind Pca.fit on the 5th-generation tensor cores:
//     C = X^T X      (X: n x dim f32 row-major, n >> dim; reference pca.zig:338 -> Matrix.gemm, Matrix.zig:763-787)
// with FP32-level accuracy from TF32 MMAs ("3xTF32"): x = big + small with big = the upper 19 bits of x
// (exactly what kind::tf32 reads from a 32-bit operand) and small = x - big, and
//     x*y ~= big*big' + big*small' + small*big'        (relative error ~2^-21 per product).
//
// Work decomposition: C is symmetric, so only the 128x128 tiles (ti <= tj) of its upper triangle are computed
// (3 of 4 tiles at dim = 256) and the finish kernel mirrors them.  The grid is tiles x K-slices ~ one CTA per SM.
//
// Accuracy: the tensor core adds into its FP32 accumulator with truncation, which biases long coherent sums
// (measured ~3e-8 of the running sum per accumulate).  So an accumulator in tensor memory only ever holds ONE
// 32-row chunk (12 MMAs); four accumulators rotate, and while the MMAs of the next chunk run the drain warps
// pull the finished one into registers and add it with round-to-nearest.  K-slice partials meet in f64.
//
// Data flow per CTA:
//   warp 0   : TMA producer -- each 32-row chunk of the tile's columns lands as boxes {32 floats x 32 rows},
//              SWIZZLE_128B_ATOM_32B.  That is the canonical MN-major SWIZZLE_128B_BASE32B UMMA layout (the only
//              MN-major layout tf32 operands have; with the 16-byte-base SW128 the MMA silently yields zeros), and
//              because A = X^T and B = X are both MN-major views of the same rows, a diagonal tile feeds both
//              operands from ONE buffer.
//   warps 2-5: compute the `small` tile next to it (x - trunc19(x)), fence to the async proxy, arrive.
//   warp 1   : one thread issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=128, K=8): per 8-row group three
//              MMAs (big.big, big.small, small.big); tcgen05.commit frees the stage and publishes the accumulator.
//   warps 6-9: tcgen05.ld the finished accumulator (32 lanes per warp), add into 128 registers per thread.
#include <cstdlib>

#include "zb_internal.h"
#include "zb_linalg.h"
#include "zb_tma.cuh"

namespace zb {

namespace {

constexpr int BK = 32;                       // rows of X per chunk
constexpr int BLK_BYTES = BK * 128;          // one {32 floats x 32 rows} box: 4096 B
constexpr int TC_SMEM_TILES = 196608;        // big + small stages: 6 x (16+16) KB (diagonal tile) or 3 x (32+32) KB
constexpr int TC_MAX_STAGES = 6;
constexpr int TC_ACCS = 4;                   // rotating 128-column TMEM accumulators
constexpr int TC_THREADS = 320;              // warp0 TMA + TMEM alloc, warp1 MMA, warps 2..5 split, warps 6..9 drain
constexpr int TC_SMEM = TC_SMEM_TILES + 1024 + 256;

struct TcParams {
    float* partial;          // [slice][tile][128][128]
    unsigned n;
    unsigned rows_per_slice; // multiple of BK
    int n_tiles;             // nb (nb + 1) / 2 upper-triangle tiles
    int nb;                  // dim / 128
};

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr) {
    // UMMA shared-memory descriptor (sm_100): start addr [0,14) >>4, LBO [16,30) >>4, SBO [32,46) >>4, version=1 [46,48),
    // base_offset 0, layout_type [61,64) = 1 (SWIZZLE_128B_BASE32B).  Canonical MN-major layout in 16-byte units:
    // ((8,n),(4,k)):((1,LBO),(8,SBO)) -- 32 floats of MN contiguous, MN blocks LBO = 4096 B apart (one TMA box), a K atom
    // is 4 rows of 128 B, and the two atoms of a K=8 MMA are SBO = 512 B apart.
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFFu);
    d |= (uint64_t)((BLK_BYTES >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((512u >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;
    return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive1(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait1(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}

__global__ void __launch_bounds__(TC_THREADS, 1) xtx_tf32x3_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TcParams p) {
    extern __shared__ unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // which tile of the upper triangle, which K slice
    const int tile = blockIdx.x % p.n_tiles, slice = blockIdx.x / p.n_tiles;
    int ti = 0, tj = tile;
    while (tj >= p.nb - ti) { tj -= p.nb - ti; ++ti; }
    tj += ti;
    const bool diag = ti == tj;
    const int nblk = diag ? 4 : 8;                    // 32-column boxes per stage: A columns, then (off-diagonal) B columns
    const int stage_bytes = nblk * BLK_BYTES;
    const int n_stages = diag ? 6 : 3;
    const uint32_t b_off = diag ? 0u : 4u * BLK_BYTES;

    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t big0 = smem0;
    const uint32_t small0 = smem0 + TC_SMEM_TILES / 2;
    const uint32_t bars = smem0 + TC_SMEM_TILES;
    const uint32_t full_bar = bars;                          // TMA landed                   [stage]
    const uint32_t ready_bar = bars + 8 * TC_MAX_STAGES;      // small tile written           [stage]
    const uint32_t empty_bar = bars + 16 * TC_MAX_STAGES;     // MMAs of the stage retired    [stage]
    const uint32_t accf_bar = bars + 24 * TC_MAX_STAGES;      // accumulator complete         [acc]
    const uint32_t acce_bar = accf_bar + 8 * TC_ACCS;         // accumulator drained          [acc]
    const uint32_t tmem_slot = acce_bar + 8 * TC_ACCS;

    const unsigned row0 = (unsigned)slice * p.rows_per_slice;
    const unsigned row1 = min(p.n, row0 + p.rows_per_slice);
    const int n_chunks = row0 < row1 ? (int)((row1 - row0 + BK - 1) / BK) : 0;

    if (tid == 0) {
        for (int s = 0; s < TC_MAX_STAGES; ++s) {
            mbar_init(full_bar + 8 * s, 1);
            mbar_init(ready_bar + 8 * s, 128);
            mbar_init(empty_bar + 8 * s, 1);
        }
        for (int a = 0; a < TC_ACCS; ++a) {
            mbar_init(accf_bar + 8 * a, 1);
            mbar_init(acce_bar + 8 * a, 128);
        }
        fence_barrier_init();
    }
    if (warp == 0) {  // all 512 TMEM columns: four 128-column accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = lds32(tmem_slot);

    if (warp == 0) {
        // ------------------------------------------ TMA producer ------------------------------------------
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
            int s = 0;
            uint32_t ph = 0;
            for (int c = 0; c < n_chunks; ++c) {
                if (c >= n_stages) mbar_wait1(empty_bar + 8 * s, ph ^ 1u);
                mbar_arrive_expect_tx(full_bar + 8 * s, stage_bytes);
                const uint32_t dst = big0 + s * stage_bytes;
                const int r = (int)(row0 + c * BK);
                for (int b = 0; b < 4; ++b) tma_load_2d(dst + b * BLK_BYTES, &tmap, 128 * ti + 32 * b, r, full_bar + 8 * s);
                if (!diag)
                    for (int b = 0; b < 4; ++b) tma_load_2d(dst + (4 + b) * BLK_BYTES, &tmap, 128 * tj + 32 * b, r, full_bar + 8 * s);
                if (++s == n_stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------ MMA issuer ------------------------------------------
        if (lane == 0) {
            // instruction descriptor: D=F32 (1<<4), A=B=TF32 (2<<7, 2<<10), A and B MN-major (bits 15,16), N>>3 at [17,23), M>>4 at [24,29)
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
            int s = 0;
            uint32_t ph = 0;
            for (int c = 0; c < n_chunks; ++c) {
                const int a = c % TC_ACCS;
                mbar_wait1(ready_bar + 8 * s, ph);   // implies the TMA data landed (the split warps waited for it)
                if (c >= TC_ACCS) mbar_wait1(acce_bar + 8 * a, ((c / TC_ACCS) - 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t big = big0 + s * stage_bytes, sml = small0 + s * stage_bytes;
                const uint32_t d = tmem_base + (uint32_t)(a * 128);
#pragma unroll
                for (int kg = 0; kg < BK / 8; ++kg) {
                    const uint64_t a_big = make_smem_desc(big + kg * 1024);
                    const uint64_t a_sml = make_smem_desc(sml + kg * 1024);
                    const uint64_t b_big = make_smem_desc(big + b_off + kg * 1024);
                    const uint64_t b_sml = make_smem_desc(sml + b_off + kg * 1024);
                    umma_tf32(d, a_big, b_big, idesc, kg != 0 ? 1u : 0u);
                    umma_tf32(d, a_big, b_sml, idesc, 1u);
                    umma_tf32(d, a_sml, b_big, idesc, 1u);
                }
                umma_commit(empty_bar + 8 * s);   // arrives when the MMAs above have finished reading this stage ...
                umma_commit(accf_bar + 8 * a);    // ... and writing this accumulator
                if (++s == n_stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp < 6) {
        // ------------------------------ split warps: small = x - trunc19(x) ------------------------------
        const int t = tid - 64;  // 0..127
        int s = 0;
        uint32_t ph = 0;
        const int n16 = stage_bytes / 16;
        for (int c = 0; c < n_chunks; ++c) {
            mbar_wait1(full_bar + 8 * s, ph);
            const uint32_t big = big0 + s * stage_bytes, sml = small0 + s * stage_bytes;
#pragma unroll 4
            for (int i = t; i < n16; i += 128) {
                const float4 v = lds128(big + i * 16);
                float4 r;
                r.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
                r.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
                r.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
                r.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
                sts128(sml + i * 16, r);
            }
            fence_proxy_async();               // make the generic-proxy writes visible to the tensor core (async proxy)
            mbar_arrive1(ready_bar + 8 * s);
            if (++s == n_stages) { s = 0; ph ^= 1u; }
        }
    } else {
        // ------------------------------ drain warps: TMEM accumulator -> registers (round-to-nearest adds) ------------------------------
        const int sub = warp & 3;  // TMEM sub-partition this warp may access: lanes [32*sub, 32*sub+32)
        float acc[128];
#pragma unroll
        for (int i = 0; i < 128; ++i) acc[i] = 0.0f;
        for (int c = 0; c < n_chunks; ++c) {
            const int a = c % TC_ACCS;
            mbar_wait1(accf_bar + 8 * a, (c / TC_ACCS) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(a * 128 + q * 32);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                      "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                      "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                      "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[q * 32 + i] += __uint_as_float(r[i]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive1(acce_bar + 8 * a);
        }
        // this CTA's partial tile, row m of 128 floats
        float4* dst = reinterpret_cast<float4*>(p.partial + ((size_t)blockIdx.x * 128 + (size_t)(sub * 32 + lane)) * 128);
#pragma unroll
        for (int q = 0; q < 32; ++q) dst[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// out = beta*c + alpha * sum_slices partial, mirrored across the diagonal   (f64 sum in a fixed order)
__global__ void __launch_bounds__(256) xtx_finish(const float* __restrict__ partial, int slices, int n_tiles, int nb, unsigned dim, float alpha,
                                                  float beta, const float* __restrict__ c, float* __restrict__ out) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)dim * dim) return;
    unsigned i = (unsigned)(idx / dim), j = (unsigned)(idx % dim);
    if (i > j) { const unsigned t = i; i = j; j = t; }   // lower triangle: read the mirror image (C comes out exactly symmetric)
    const int ti = (int)(i / 128), tj = (int)(j / 128);
    const int tile = ti * nb - ti * (ti - 1) / 2 + (tj - ti);
    const size_t off = (size_t)tile * 16384 + (size_t)(i % 128) * 128 + (j % 128);
    double s = 0.0;
    for (int z = 0; z < slices; ++z) s += (double)partial[(size_t)z * n_tiles * 16384 + off];
    const float prod = (float)((double)alpha * s);
    out[idx] = (c != nullptr && beta != 0.0f) ? (beta * c[idx] + prod) : prod;
}

}  // namespace

// C = alpha * X^T X + beta * C on the tensor cores.  Returns ZB_ERR_UNSUPPORTED outside its envelope.
int gemm_xtx_tensorcore(const float* x, uint32_t n, uint32_t dim, float alpha, float beta, const float* c, float* out, cudaStream_t s) {
    if (dim == 0 || dim % 128 != 0 || dim > 1024) return ZB_ERR_UNSUPPORTED;
    if (n < 4096) return ZB_ERR_UNSUPPORTED;
    if ((uintptr_t)x & 15u) return ZB_ERR_UNSUPPORTED;
    EncodeTiledFn encode = encode_tiled_fn();
    if (!encode) return ZB_ERR_UNSUPPORTED;
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    if (di.smem_optin < (size_t)TC_SMEM) return ZB_ERR_UNSUPPORTED;
    const int nb = (int)(dim / 128), n_tiles = nb * (nb + 1) / 2;
    int slices = std::max(1, di.sm_count / n_tiles);
    unsigned rows_per_slice = (n + slices - 1) / slices;
    rows_per_slice = ((rows_per_slice + BK - 1) / BK) * BK;
    slices = (int)((n + rows_per_slice - 1) / rows_per_slice);
    const int grid = slices * n_tiles;
    Scratch part;
    if ((rc = part.alloc((size_t)grid * 16384 * sizeof(float), s))) return rc;
    TcParams p{part.as<float>(), n, rows_per_slice, n_tiles, nb};
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {dim, n};
    const cuuint64_t gstr[1] = {(cuuint64_t)dim * 4};
    const cuuint32_t box[2] = {32, (cuuint32_t)BK};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)x, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return ZB_ERR_UNSUPPORTED;
    ZB_CUDA(cudaFuncSetAttribute(xtx_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    xtx_tf32x3_kernel<<<grid, TC_THREADS, TC_SMEM, s>>>(tmap, p);
    ZB_LAUNCHED();
    xtx_finish<<<div_up((size_t)dim * dim, 256), 256, 0, s>>>(part.as<float>(), slices, n_tiles, nb, dim, alpha, beta, c, out);
    ZB_LAUNCHED();
    t_last_kernel = "gemm_xtx_tf32x3_tcgen05";
    return ZB_OK;
}

}  // namespace zb
