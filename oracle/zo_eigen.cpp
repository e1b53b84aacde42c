// zo_eigen.cpp -- CPU oracle (TEST INFRASTRUCTURE ONLY, see zignal_oracle.h) for Matrix.eigh (reference matrix/eigen.zig:34-136):
// symmetric eigendecomposition by cyclic Jacobi rotations, eigenvalues ascending, eigenvectors as columns.  Row-major n x n.
#include "zo_common.h"

namespace zo {

// returns ZO_OK, ZO_ERR_NOT_FINITE or ZO_ERR_NOT_SYMMETRIC
template <typename T>
static int eigh(const T* a, uint32_t n, T* values, T* vectors) {
    const size_t nn = (size_t)n * n;
    T max_abs = 0;                                                                 // eigen.zig:42-54
    for (size_t i = 0; i < nn; ++i) {
        if (!std::isfinite(a[i])) return ZO_ERR_NOT_FINITE;
        max_abs = std::max(max_abs, std::fabs(a[i]));
    }
    const T eps = std::numeric_limits<T>::epsilon();
    const T tol = max_abs * std::sqrt(eps);
    for (uint32_t i = 0; i < n; ++i)
        for (uint32_t j = i + 1; j < n; ++j)
            if (std::fabs(a[(size_t)i * n + j] - a[(size_t)j * n + i]) > tol) return ZO_ERR_NOT_SYMMETRIC;
    std::vector<T> work(a, a + nn);                                                // :57-60
    for (size_t i = 0; i < nn; ++i) vectors[i] = 0;
    for (uint32_t i = 0; i < n; ++i) vectors[(size_t)i * n + i] = 1;
    T frob_sq = 0;                                                                 // :64-67
    for (size_t i = 0; i < nn; ++i) frob_sq += work[i] * work[i];
    const T off_tol = frob_sq * eps * eps;
    auto W = [&](uint32_t r, uint32_t c) -> T& { return work[(size_t)r * n + c]; };
    auto V = [&](uint32_t r, uint32_t c) -> T& { return vectors[(size_t)r * n + c]; };
    for (int sweep = 0; sweep < 100; ++sweep) {                                    // :69-112
        T off = 0;
        for (uint32_t p = 0; p < n; ++p)
            for (uint32_t q = p + 1; q < n; ++q) off += W(p, q) * W(p, q);
        if (off <= off_tol) break;
        for (uint32_t p = 0; p < n; ++p)
            for (uint32_t q = p + 1; q < n; ++q) {
                const T apq = W(p, q);
                if (apq == 0) continue;
                const T theta = (T)0.5 * (W(q, q) - W(p, p)) / apq;
                T t;
                if (std::fabs(theta) > (T)1.0 / std::sqrt(eps)) {
                    t = (T)0.5 / theta;
                } else {
                    const T sign = theta < 0 ? (T)-1 : (T)1;
                    t = sign / (std::fabs(theta) + std::sqrt(theta * theta + 1));
                }
                const T c = (T)1.0 / std::sqrt(t * t + 1);
                const T s = t * c;
                for (uint32_t k = 0; k < n; ++k) {
                    const T akp = W(k, p), akq = W(k, q);
                    W(k, p) = c * akp - s * akq;
                    W(k, q) = s * akp + c * akq;
                }
                for (uint32_t k = 0; k < n; ++k) {
                    const T apk = W(p, k), aqk = W(q, k);
                    W(p, k) = c * apk - s * aqk;
                    W(q, k) = s * apk + c * aqk;
                }
                for (uint32_t k = 0; k < n; ++k) {
                    const T vkp = V(k, p), vkq = V(k, q);
                    V(k, p) = c * vkp - s * vkq;
                    V(k, q) = s * vkp + c * vkq;
                }
            }
    }
    for (uint32_t i = 0; i < n; ++i) values[i] = W(i, i);                          // :114-116
    for (uint32_t i = 0; i < n; ++i) {                                             // :118-129 selection sort, columns follow
        uint32_t min_idx = i;
        for (uint32_t j = i + 1; j < n; ++j)
            if (values[j] < values[min_idx]) min_idx = j;
        if (min_idx != i) {
            std::swap(values[i], values[min_idx]);
            for (uint32_t k = 0; k < n; ++k) std::swap(V(k, i), V(k, min_idx));
        }
    }
    return ZO_OK;
}

}  // namespace zo

extern "C" int zo_eigh_f64(const double* a, uint32_t rows, uint32_t cols, double* values, double* vectors) {
    if (rows != cols) return ZO_ERR_NOT_SQUARE;                                    // eigen.zig:36
    return zo::eigh<double>(a, rows, values, vectors);
}
extern "C" int zo_eigh_f32(const float* a, uint32_t rows, uint32_t cols, float* values, float* vectors) {
    if (rows != cols) return ZO_ERR_NOT_SQUARE;
    return zo::eigh<float>(a, rows, values, vectors);
}
