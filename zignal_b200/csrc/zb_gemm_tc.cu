// zb_gemm_tc.cu -- the dense contraction behind Pca.fit on the 5th-generation tensor cores:
//     C = X^T X      (X: n x dim f32 row-major, n >> dim; reference pca.zig:338 -> Matrix.gemm, Matrix.zig:763-787)
// with FP32-level accuracy from TF32 MMAs ("3xTF32"): x = big + small with big = the upper 19 bits of x
// (exactly what kind::tf32 reads from a 32-bit operand) and small = x - big, and
//     x*y ~= big*big' + big*small' + small*big'        (relative error ~2^-21 per product)
// accumulated in FP32 in tensor memory.
//
// Data flow per CTA (one per SM, persistent over its slice of the n rows):
//   warp 0   : TMA producer -- each 32-row chunk of X lands in shared memory as 8 boxes {32 floats x 32 rows},
//              SWIZZLE_128B.  That is precisely the canonical MN-major SW128 UMMA layout, and because
//              A = X^T and B = X are both "MN-major" views of the same rows, ONE tile feeds both operands.
//   warps 2-5: compute the `small` tile next to it (x - trunc19(x)), fence to the async proxy, arrive.
//   warp 1   : one thread issues tcgen05.mma.cta_group::1.kind::tf32  (M=128, N=dim, K=8): per 8-row group and
//              per 128-row half of C three MMAs (big.big, big.small, small.big); tcgen05.commit frees the stage.
//   warps 2-5: at the end read the accumulators back with tcgen05.ld (32 lanes per warp) and store this CTA's
//              partial C; a small finish kernel sums the partials in f64 in a fixed order (deterministic),
//              applies alpha / beta.
// Accumulators: 2 x (128 lanes x 256 columns) f32 = all 512 TMEM columns.
#include "zb_internal.h"
#include "zb_linalg.h"
#include "zb_tma.cuh"

namespace zb {

namespace {

constexpr int BK = 32;                       // rows of X per stage
constexpr int BLK_BYTES = BK * 128;          // one {32 floats x 32 rows} box: 4096 B
constexpr int MAX_BLKS = 8;                  // dim <= 256
constexpr int TC_STAGES = 3;
constexpr int TC_THREADS = 192;              // warp0 TMA+TMEM alloc, warp1 MMA, warps 2..5 split + epilogue

struct TcParams {
    float* partial;        // [grid][dim][dim]
    unsigned n, dim;
    unsigned rows_per_cta; // multiple of BK
};

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // UMMA shared-memory descriptor (sm_100): start addr [0,14) >>4, LBO [16,30) >>4, SBO [32,46) >>4, version=1 [46,48),
    // base_offset 0, layout_type [61,64) = 2 (SWIZZLE_128B)
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
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
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"l"((uint64_t)bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive1(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait1(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}

template <int DIM>
__global__ void __launch_bounds__(TC_THREADS, 1) xtx_tf32x3_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TcParams p) {
    constexpr int NBLK = DIM / 32;                 // 32-column blocks of X
    constexpr int TILE_BYTES = NBLK * BLK_BYTES;   // one stage of `big` (and of `small`)
    constexpr int MHALVES = DIM / 128;             // 128-row halves of C
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t big0 = smem0;
    const uint32_t small0 = smem0 + TC_STAGES * TILE_BYTES;
    const uint32_t bars = small0 + TC_STAGES * TILE_BYTES;
    const uint32_t full_bar = bars;                 // TMA landed          [stage]
    const uint32_t ready_bar = bars + 8 * TC_STAGES;   // small tile written  [stage]
    const uint32_t empty_bar = bars + 16 * TC_STAGES;  // MMAs of the stage retired [stage]
    const uint32_t done_bar = bars + 24 * TC_STAGES;   // all MMAs retired
    const uint32_t tmem_slot = done_bar + 8;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned row0 = blockIdx.x * p.rows_per_cta;
    const unsigned row1 = min(p.n, row0 + p.rows_per_cta);
    const int n_chunks = row0 < row1 ? (int)((row1 - row0 + BK - 1) / BK) : 0;

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(full_bar + 8 * s, 1);
            mbar_init(ready_bar + 8 * s, 128);
            mbar_init(empty_bar + 8 * s, 1);
        }
        mbar_init(done_bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) {  // allocate all 512 TMEM columns (2 accumulators of 256 columns)
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
            for (int c = 0; c < n_chunks; ++c) {
                const int s = c % TC_STAGES;
                if (c >= TC_STAGES) mbar_wait1(empty_bar + 8 * s, ((c / TC_STAGES) - 1) & 1u);
                mbar_arrive_expect_tx(full_bar + 8 * s, TILE_BYTES);
                for (int b = 0; b < NBLK; ++b)
                    tma_load_2d(big0 + s * TILE_BYTES + b * BLK_BYTES, &tmap, 32 * b, (int)(row0 + c * BK), full_bar + 8 * s);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------ MMA issuer ------------------------------------------
        if (lane == 0) {
            // instruction descriptor: D=F32 (1<<4), A=B=TF32 (2<<7, 2<<10), A and B MN-major (bits 15,16), N>>3 at [17,23), M>>4 at [24,29)
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(DIM >> 3) << 17) | ((128u >> 4) << 24);
            for (int c = 0; c < n_chunks; ++c) {
                const int s = c % TC_STAGES;
                mbar_wait1(ready_bar + 8 * s, (c / TC_STAGES) & 1u);   // implies the TMA data landed (the split warps waited for it)
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t big = big0 + s * TILE_BYTES, sml = small0 + s * TILE_BYTES;
#pragma unroll
                for (int kg = 0; kg < BK / 8; ++kg) {
                    const uint64_t b_big = make_smem_desc(big + kg * 1024, BLK_BYTES, 1024);
                    const uint64_t b_sml = make_smem_desc(sml + kg * 1024, BLK_BYTES, 1024);
#pragma unroll
                    for (int h = 0; h < MHALVES; ++h) {
                        const uint64_t a_big = make_smem_desc(big + kg * 1024 + h * 4 * BLK_BYTES, BLK_BYTES, 1024);
                        const uint64_t a_sml = make_smem_desc(sml + kg * 1024 + h * 4 * BLK_BYTES, BLK_BYTES, 1024);
                        const uint32_t d = tmem_base + (uint32_t)(h * 256);
                        umma_tf32(d, a_big, b_big, idesc, (c | kg) != 0 ? 1u : 0u);
                        umma_tf32(d, a_big, b_sml, idesc, 1u);
                        umma_tf32(d, a_sml, b_big, idesc, 1u);
                    }
                }
                umma_commit(empty_bar + 8 * s);   // arrives when the MMAs above have finished reading this stage
            }
            umma_commit(done_bar);
        }
    } else {
        // ------------------------------ split warps (then epilogue) ------------------------------
        const int t = tid - 64;  // 0..127
        for (int c = 0; c < n_chunks; ++c) {
            const int s = c % TC_STAGES;
            mbar_wait1(full_bar + 8 * s, (c / TC_STAGES) & 1u);
            const uint32_t big = big0 + s * TILE_BYTES, sml = small0 + s * TILE_BYTES;
#pragma unroll 4
            for (int i = t; i < TILE_BYTES / 16; i += 128) {
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
        }
        // epilogue: accumulators -> this CTA's partial C
        mbar_wait1(done_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int sub = warp & 3;  // TMEM sub-partition this warp may access: lanes [32*sub, 32*sub+32)
        float* out = p.partial + (size_t)blockIdx.x * DIM * DIM;
#pragma unroll 1
        for (int h = 0; h < MHALVES; ++h) {
            const int m = h * 128 + sub * 32 + lane;
#pragma unroll 1
            for (int c0 = 0; c0 < DIM; c0 += 32) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(h * 256 + c0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                      "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                      "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                      "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (m < DIM) {
                    float4* dst = reinterpret_cast<float4*>(out + (size_t)m * DIM + c0);
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        dst[q] = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                                             __uint_as_float(r[4 * q + 3]));
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// out = beta*c + alpha * sum_cta partial   (f64 sum in a fixed order)
__global__ void __launch_bounds__(256) xtx_finish(const float* __restrict__ partial, int parts, size_t mn, float alpha, float beta,
                                                  const float* __restrict__ c, float* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= mn) return;
    double s = 0.0;
    for (int z = 0; z < parts; ++z) s += (double)partial[(size_t)z * mn + i];
    const float prod = (float)((double)alpha * s);
    out[i] = (c != nullptr && beta != 0.0f) ? (beta * c[i] + prod) : prod;
}

template <int DIM>
int launch_xtx(const CUtensorMap& tmap, const TcParams& p, int grid, cudaStream_t s) {
    constexpr int smem = 2 * TC_STAGES * (DIM / 32) * BLK_BYTES + 256 + 1024;
    auto k = xtx_tf32x3_kernel<DIM>;
    ZB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k<<<grid, TC_THREADS, smem, s>>>(tmap, p);
    ZB_LAUNCHED();
    return ZB_OK;
}

}  // namespace

// C = alpha * X^T X + beta * C on the tensor cores.  Returns ZB_ERR_UNSUPPORTED outside its envelope.
int gemm_xtx_tensorcore(const float* x, uint32_t n, uint32_t dim, float alpha, float beta, const float* c, float* out, cudaStream_t s) {
    if (dim != 128 && dim != 256) return ZB_ERR_UNSUPPORTED;
    if (n < 4096) return ZB_ERR_UNSUPPORTED;
    if ((uintptr_t)x & 15u) return ZB_ERR_UNSUPPORTED;
    EncodeTiledFn encode = encode_tiled_fn();
    if (!encode) return ZB_ERR_UNSUPPORTED;
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    const int smem_need = 2 * TC_STAGES * (int)(dim / 32) * BLK_BYTES + 256 + 1024;
    if (di.smem_optin < (size_t)smem_need) return ZB_ERR_UNSUPPORTED;
    int grid = di.sm_count;
    unsigned rows_per_cta = (n + grid - 1) / grid;
    rows_per_cta = ((rows_per_cta + BK - 1) / BK) * BK;
    grid = (int)((n + rows_per_cta - 1) / rows_per_cta);
    Scratch part;
    if ((rc = part.alloc((size_t)grid * dim * dim * sizeof(float), s))) return rc;
    TcParams p{part.as<float>(), n, dim, rows_per_cta};
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {dim, n};
    const cuuint64_t gstr[1] = {(cuuint64_t)dim * 4};
    const cuuint32_t box[2] = {32, (cuuint32_t)BK};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)x, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return ZB_ERR_UNSUPPORTED;
    rc = dim == 256 ? launch_xtx<256>(tmap, p, grid, s) : launch_xtx<128>(tmap, p, grid, s);
    if (rc) return rc;
    const size_t mn = (size_t)dim * dim;
    xtx_finish<<<div_up(mn, 256), 256, 0, s>>>(part.as<float>(), grid, mn, alpha, beta, c, out);
    ZB_LAUNCHED();
    t_last_kernel = "gemm_xtx_tf32x3_tcgen05";
    return ZB_OK;
}

}  // namespace zb
