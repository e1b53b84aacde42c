// zb_linalg.h -- host SVD shared with zb_fdm.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace zb {
// Golub-Reinsch SVD (reference svd.zig:149-496).  a: m x n row-major (m >= n); u: m x ucols; q: n; v: n x n.
// Returns 0 or the index of the singular value whose QR iteration failed to converge.
template <typename T>
uint64_t svd_golub_reinsch(const T* a, uint32_t m, uint32_t n, int mode, bool with_v, T* u, uint32_t ucols, T* q, T* v);
// zb_gemm_tc.cu: C = alpha * X^T X + beta * C with 3xTF32 tcgen05 MMAs (dim in {128, 256}, n >= 4096); ZB_ERR_UNSUPPORTED otherwise.
int gemm_xtx_tensorcore(const float* x, uint32_t n, uint32_t dim, float alpha, float beta, const float* c, float* out, cudaStream_t s);
}  // namespace zb
