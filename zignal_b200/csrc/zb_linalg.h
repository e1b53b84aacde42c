// zb_linalg.h -- linear-algebra internals shared between translation units.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace zb {
// zb_gemm_tc.cu: C = alpha * X^T X + beta * C with 3xTF32 tcgen05 MMAs (dim in {128, 256}, n >= 4096); ZB_ERR_UNSUPPORTED otherwise.
int gemm_xtx_tensorcore(const float* x, uint32_t n, uint32_t dim, float alpha, float beta, const float* c, float* out, cudaStream_t s);
}  // namespace zb
