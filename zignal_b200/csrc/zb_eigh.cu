// zb_eigh.cu -- Matrix.eigh (reference matrix/eigen.zig:34-136): symmetric eigendecomposition A = V diag(lambda) V^T by
// cyclic Jacobi rotations.  Like the SVD it is a small sequential host routine ("replicas only"): the matrices it is
// meant for (covariances, Hessians) are a few dozen rows.  Row-major n x n in, eigenvalues ascending, eigenvectors as columns.
#include <cmath>
#include <limits>
#include <utility>
#include <vector>

#include "zb_internal.h"

namespace zb {
namespace {

template <typename T>
struct JacobiRotation {
    T c, s;
    // eigen.zig:79-92: t = tan of the rotation angle that annihilates a(p, q); large |theta| uses 0.5 / theta (no theta^2 overflow).
    static JacobiRotation annihilating(T app, T aqq, T apq, T eps) {
        const T theta = (T)0.5 * (aqq - app) / apq;
        T t;
        if (std::fabs(theta) > (T)1.0 / std::sqrt(eps)) {
            t = (T)0.5 / theta;
        } else {
            const T sign = theta < 0 ? (T)-1 : (T)1;
            t = sign / (std::fabs(theta) + std::sqrt(theta * theta + 1));
        }
        const T c = (T)1.0 / std::sqrt(t * t + 1);
        return {c, t * c};
    }
    // (x, y) <- (c x - s y, s x + c y) on two strided vectors of length n
    void apply(T* x, size_t x_step, T* y, size_t y_step, uint32_t n) const {
        for (uint32_t k = 0; k < n; ++k, x += x_step, y += y_step) {
            const T xv = *x, yv = *y;
            *x = c * xv - s * yv;
            *y = s * xv + c * yv;
        }
    }
};

template <typename T>
int eigh_host(const T* a, uint32_t rows, uint32_t cols, T* values, T* vectors) {
    if (!values || !vectors || (!a && rows)) return ZB_ERR_INVALID_ARGUMENT;
    if (rows != cols) return ZB_ERR_NOT_SQUARE;                                    // :36
    const uint32_t n = rows;
    const size_t nn = (size_t)n * n;
    const T eps = std::numeric_limits<T>::epsilon();
    T max_abs = 0;
    for (size_t i = 0; i < nn; ++i) {                                              // :45-50
        if (!std::isfinite(a[i])) return ZB_ERR_NOT_FINITE;
        max_abs = std::max(max_abs, std::fabs(a[i]));
    }
    const T sym_tol = max_abs * std::sqrt(eps);                                    // :51-54
    for (uint32_t i = 0; i < n; ++i)
        for (uint32_t j = i + 1; j < n; ++j)
            if (std::fabs(a[(size_t)i * n + j] - a[(size_t)j * n + i]) > sym_tol) return ZB_ERR_NOT_SYMMETRIC;

    std::vector<T> w(a, a + nn);
    std::fill(vectors, vectors + nn, (T)0);
    for (uint32_t i = 0; i < n; ++i) vectors[(size_t)i * n + i] = 1;
    T frob_sq = 0;
    for (T x : w) frob_sq += x * x;                                                // :64-67
    const T off_tol = frob_sq * eps * eps;

    for (int sweep = 0; sweep < 100; ++sweep) {                                    // :69-112
        T off = 0;
        for (uint32_t p = 0; p < n; ++p)
            for (uint32_t q = p + 1; q < n; ++q) off += w[(size_t)p * n + q] * w[(size_t)p * n + q];
        if (off <= off_tol) break;
        for (uint32_t p = 0; p < n; ++p)
            for (uint32_t q = p + 1; q < n; ++q) {
                const T apq = w[(size_t)p * n + q];
                if (apq == 0) continue;
                const auto rot = JacobiRotation<T>::annihilating(w[(size_t)p * n + p], w[(size_t)q * n + q], apq, eps);
                rot.apply(&w[p], n, &w[q], n, n);                                  // columns p, q of A
                rot.apply(&w[(size_t)p * n], 1, &w[(size_t)q * n], 1, n);          // rows p, q of A
                rot.apply(&vectors[p], n, &vectors[q], n, n);                      // V <- V J
            }
    }
    for (uint32_t i = 0; i < n; ++i) values[i] = w[(size_t)i * n + i];
    for (uint32_t i = 0; i < n; ++i) {                                             // :118-129
        uint32_t lo = i;
        for (uint32_t j = i + 1; j < n; ++j)
            if (values[j] < values[lo]) lo = j;
        if (lo == i) continue;
        std::swap(values[i], values[lo]);
        for (uint32_t k = 0; k < n; ++k) std::swap(vectors[(size_t)k * n + i], vectors[(size_t)k * n + lo]);
    }
    return ZB_OK;
}

}  // namespace
}  // namespace zb

extern "C" int zb_eigh_f64(const double* a, uint32_t rows, uint32_t cols, double* values, double* vectors) {
    return zb::eigh_host<double>(a, rows, cols, values, vectors);
}
extern "C" int zb_eigh_f32(const float* a, uint32_t rows, uint32_t cols, float* values, float* vectors) {
    return zb::eigh_host<float>(a, rows, cols, values, vectors);
}
