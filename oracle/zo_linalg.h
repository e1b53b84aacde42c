// zo_linalg.h -- internal declarations shared by zo_linalg.cpp and zo_fdm.cpp (oracle, TEST INFRASTRUCTURE ONLY).
#pragma once
#include "zo_common.h"

namespace zo {
// svd.zig:149-496.  u is m x u_cols row-major (u_cols = m for full_u, n otherwise), q has n entries, v is n x n.
template <typename T>
int64_t svd_kernel(const T* a, uint32_t m, uint32_t n, int mode, bool with_v, T* u, uint32_t u_cols, T* q, T* v);
template <typename T>
int64_t svd(const T* a, uint32_t m, uint32_t n, int mode, bool with_v, T* u, T* s, T* v);
// Matrix.zig:696-822
template <typename T>
int gemm(const T* a, uint32_t ar, uint32_t ac, bool trans_a, const T* b, uint32_t br, uint32_t bc, bool trans_b, T alpha,
         T beta, const T* c, T* out, int vec_len);
}  // namespace zo
