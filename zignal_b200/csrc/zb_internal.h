// zb_internal.h -- shared internals of libzignal_b200.so (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/zignal_b200.h"

namespace zb {

extern std::atomic<uint64_t> g_launches;
extern thread_local char t_last_error[512];
extern thread_local const char* t_last_kernel;
extern std::atomic<int> g_exact_f32;
extern std::atomic<int> g_force_generic;
extern std::atomic<int> g_tune_stages;     // fused conv: TMA pipeline depth (2 or 3)
extern std::atomic<int> g_tune_f2;         // fused conv: packed f32x2 FMA
extern std::atomic<int> g_tune_band_rows;
extern std::atomic<int> g_tune_host_band_rows;  // host-pointer pipeline: rows per PCIe band (0 disables the pipeline)
extern std::atomic<int> g_tune_edge_fast;  // fused RGBA f32 conv: x borders of .replicate / .mirror as in-stage copies (default on; 0 = generic fixup pass)
extern std::atomic<int> g_tune_sobel_tile; // Image.sobel on gray u8: byte-tile kernel (default on; 0 = per-pixel kernel)
extern std::atomic<int> g_tune_jacobi_cluster;  // Jacobi SVD in one cluster's distributed shared memory when it fits (default on)
extern std::atomic<int> g_tune_u8_dp;      // fused RGBA8 conv: dp4a / dp2a variant when every tap is a byte (default on)
extern std::atomic<int> g_tune_u8_fmath;   // fused RGBA8 conv: run the exact-integer pipeline on FFMA when provably exact
extern std::atomic<int> g_tune_variant;    // fused conv: -1 auto, 0 = phase-synchronous kernel, 1 = warp-specialised kernel  // fused conv: target rows per work unit

int set_cuda_error(cudaError_t e, const char* what, const char* file, int line);

#define ZB_CUDA(expr)                                                                    \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) return ::zb::set_cuda_error(_e, #expr, __FILE__, __LINE__); \
    } while (0)

// Count + check a kernel launch.
#define ZB_LAUNCHED()                                                                      \
    do {                                                                                   \
        ::zb::g_launches.fetch_add(1, std::memory_order_relaxed);                          \
        cudaError_t _e = cudaGetLastError();                                               \
        if (_e != cudaSuccess) return ::zb::set_cuda_error(_e, "kernel launch", __FILE__, __LINE__); \
    } while (0)

struct DeviceInfo {
    int ordinal = -1;
    int sm_count = 0;
    size_t smem_optin = 0;
};
// Lazily initialised per-device state (mempool threshold, SM count, driver entry points).
int device_info(DeviceInfo* out);
// cuTensorMapEncodeTiled resolved through cudaGetDriverEntryPoint (no link-time libcuda dependency).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn();

static inline int channels_of(int pixfmt) {
    switch (pixfmt) {
        case ZB_PIX_U8: case ZB_PIX_F32: return 1;
        case ZB_PIX_RGB8: return 3;
        case ZB_PIX_RGBA8: case ZB_PIX_RGBAF32: return 4;
    }
    return 0;
}
static inline size_t channel_bytes(int pixfmt) { return (pixfmt == ZB_PIX_F32 || pixfmt == ZB_PIX_RGBAF32) ? 4 : 1; }
static inline size_t pixel_bytes(int pixfmt) { return (size_t)channels_of(pixfmt) * channel_bytes(pixfmt); }
static inline bool is_float_fmt(int pixfmt) { return pixfmt == ZB_PIX_F32 || pixfmt == ZB_PIX_RGBAF32; }

// RAII stream-ordered scratch (cudaMallocAsync / cudaFreeAsync on the op's stream).
struct Scratch {
    void* p = nullptr;
    cudaStream_t s = nullptr;
    ~Scratch() { if (p) cudaFreeAsync(p, s); }
    int alloc(size_t bytes, cudaStream_t stream);
    template <typename T> T* as() const { return (T*)p; }
};

// Do the byte ranges two image views span intersect?  (Two views of one buffer with different base pointers may still
// overlap; single-pass kernels read neighbourhoods other blocks are writing, so any overlap must take the snapshot path.)
static inline bool images_overlap(const zb_image* a, const zb_image* b, size_t pixel_bytes_) {
    if (!a->data || !b->data || a->rows == 0 || a->cols == 0 || b->rows == 0 || b->cols == 0) return false;
    const uintptr_t a0 = (uintptr_t)a->data, b0 = (uintptr_t)b->data;
    const uintptr_t a1 = a0 + ((size_t)(a->rows - 1) * a->stride + a->cols) * pixel_bytes_;
    const uintptr_t b1 = b0 + ((size_t)(b->rows - 1) * b->stride + b->cols) * pixel_bytes_;
    return a0 < b1 && b0 < a1;
}

static inline unsigned div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// One grid row per image row without the 65,535 limit of gridDim.y: rows are spread over (y, z); kernels read the row with
// ZB_GRID_ROW() and return when it is >= rows (the last z-slice may be partly empty).
#ifdef __CUDACC__
#define ZB_GRID_ROW() ((int)(blockIdx.y + blockIdx.z * gridDim.y))
static inline dim3 row_grid(unsigned x_blocks, size_t rows) {
    const size_t gz = (rows + 65534) / 65535;
    const size_t gy = gz ? (rows + gz - 1) / gz : 0;
    return dim3(x_blocks, (unsigned)gy, (unsigned)(gz ? gz : 1));
}
#endif

}  // namespace zb
