// zb_svd_core.h -- Golub-Reinsch SVD (reference matrix/svd.zig:149-496), one body for the host entry points and for the
// device-side 3x3 solves of the FDM path.  Only + - * / sqrt on T: IEEE-exact on both sides (the library is built without
// FMA contraction), so host and device results are bit-identical.
// u: m x ucols row-major; q: n singular values; v: n x n; e: n values of workspace.  Returns 0 or the index that failed.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>

#include "../../include/zignal_b200.h"

#ifdef __CUDACC__
#define ZB_HD __host__ __device__
#else
#define ZB_HD
#endif

namespace zb {

template <typename T> struct SvdLimits;
template <> struct SvdLimits<float> {
    ZB_HD static float epsilon() { return 1.1920928955078125e-07f; }
    ZB_HD static float min_normal() { return 1.1754943508222875e-38f; }
};
template <> struct SvdLimits<double> {
    ZB_HD static double epsilon() { return 2.220446049250313e-16; }
    ZB_HD static double min_normal() { return 2.2250738585072014e-308; }
};

template <typename T>
ZB_HD uint64_t svd_gr_core(const T* a, uint32_t m_, uint32_t n_, int mode, bool with_v, T* u, uint32_t ucols_, T* q, T* v, T* e) {
    const size_t m = m_, n = n_, ucols = ucols_;
    auto U = [&](size_t i, size_t j) -> T& { return u[i * ucols + j]; };
    auto V = [&](size_t i, size_t j) -> T& { return v[i * n + j]; };
    for (size_t i = 0; i < n; ++i) e[i] = (T)0;
    T eps = SvdLimits<T>::epsilon();
    const T tol = SvdLimits<T>::min_normal() / eps;
    const size_t max_iterations = 300;  // svd.zig:178
    uint64_t failed = 0;
    T c = 0, f = 0, g = 0, h = 0, s = 0, x = 0, y = 0, z = 0;
    size_t l = 0;

    for (size_t i = 0; i < m; ++i)
        for (size_t j = 0; j < n; ++j) U(i, j) = a[i * n + j];

    // 1. Householder reduction to bidiagonal form: q holds the diagonal, e the super-diagonal
    for (size_t i = 0; i < n; ++i) {
        e[i] = g;
        l = i + 1;
        s = 0;
        for (size_t j = i; j < m; ++j) s += U(j, i) * U(j, i);
        if (s < tol) {
            g = 0;
        } else {
            f = U(i, i);
            g = f < 0 ? sqrt(s) : -sqrt(s);
            h = f * g - s;
            U(i, i) = f - g;
            for (size_t j = l; j < n; ++j) {
                s = 0;
                for (size_t k = i; k < m; ++k) s += U(k, i) * U(k, j);
                f = s / h;
                for (size_t k = i; k < m; ++k) U(k, j) += f * U(k, i);
            }
        }
        q[i] = g;
        s = 0;
        for (size_t j = l; j < n; ++j) s += U(i, j) * U(i, j);
        if (s < tol) {
            g = 0;
        } else {
            f = U(i, i + 1);
            g = f < 0 ? sqrt(s) : -sqrt(s);
            h = f * g - s;
            U(i, i + 1) = f - g;
            for (size_t j = l; j < n; ++j) e[j] = U(i, j) / h;
            for (size_t j = l; j < m; ++j) {
                s = 0;
                for (size_t k = l; k < n; ++k) s += U(j, k) * U(i, k);
                for (size_t k = l; k < n; ++k) U(j, k) += s * e[k];
            }
        }
        y = fabs(q[i]) + fabs(e[i]);
        x = x > y ? x : y;
    }
    // 2. accumulate the right-hand transformations
    if (with_v) {
        for (size_t i = n; i-- > 0;) {
            if (g != 0) {
                h = U(i, i + 1) * g;
                for (size_t j = l; j < n; ++j) V(j, i) = U(i, j) / h;
                for (size_t j = l; j < n; ++j) {
                    s = 0;
                    for (size_t k = l; k < n; ++k) s += U(i, k) * V(k, j);
                    for (size_t k = l; k < n; ++k) V(k, j) += s * V(k, i);
                }
            }
            for (size_t j = l; j < n; ++j) V(i, j) = V(j, i) = 0;
            V(i, i) = 1;
            g = e[i];
            l = i;
        }
    }
    // 3. accumulate the left-hand transformations
    if (mode != ZB_SVD_NO_U) {
        for (size_t i = n; i < m; ++i) {
            for (size_t j = n; j < ucols; ++j) U(i, j) = 0;
            if (i < ucols) U(i, i) = 1;
        }
        for (size_t i = n; i-- > 0;) {
            l = i + 1;
            g = q[i];
            for (size_t j = l; j < ucols; ++j) U(i, j) = 0;
            if (g != 0) {
                h = U(i, i) * g;
                for (size_t j = l; j < ucols; ++j) {
                    s = 0;
                    for (size_t k = l; k < m; ++k) s += U(k, i) * U(k, j);
                    f = s / h;
                    for (size_t k = i; k < m; ++k) U(k, j) += f * U(k, i);
                }
                for (size_t j = i; j < m; ++j) U(j, i) /= g;
            } else {
                for (size_t j = i; j < m; ++j) U(j, i) = 0;
            }
            U(i, i) += 1;
        }
    }
    // 4. diagonalise the bidiagonal form with implicitly shifted QR sweeps
    eps *= x;
    for (size_t k = n; k-- > 0;) {
        size_t iter = 0;
        for (;;) {
            // test for splitting
            bool cancel = false;
            for (l = k;; --l) {
                if (fabs(e[l]) <= eps) break;           // e[0] == 0, so l never underflows
                if (fabs(q[l - 1]) <= eps) { cancel = true; break; }
            }
            if (cancel) {  // cancellation of e[l], l > 0
                c = 0;
                s = 1;
                const size_t l1 = l - 1;
                for (size_t i = l; i <= k; ++i) {
                    f = s * e[i];
                    e[i] *= c;
                    if (fabs(f) <= eps) break;
                    g = q[i];
                    h = sqrt(f * f + g * g);
                    q[i] = h;
                    c = g / h;
                    s = -f / h;
                    if (mode != ZB_SVD_NO_U)
                        for (size_t j = 0; j < m; ++j) {
                            y = U(j, l1);
                            z = U(j, i);
                            U(j, l1) = y * c + z * s;
                            U(j, i) = -y * s + z * c;
                        }
                }
            }
            // test for convergence
            z = q[k];
            if (l == k) {
                if (z < 0) {  // make the singular value non-negative
                    q[k] = -z;
                    if (with_v)
                        for (size_t j = 0; j < n; ++j) V(j, k) = -V(j, k);
                }
                break;
            }
            if (++iter > max_iterations) { failed = k; break; }
            // shift from the bottom 2x2 minor
            x = q[l];
            y = q[k - 1];
            g = e[k - 1];
            h = e[k];
            f = ((y - z) * (y + z) + (g - h) * (g + h)) / (2 * h * y);
            g = sqrt(f * f + (T)1);
            f = ((x - z) * (x + z) + h * (y / (f < 0 ? f - g : f + g) - h)) / x;
            // next QR transformation
            c = s = 1;
            for (size_t i = l + 1; i <= k; ++i) {
                g = e[i];
                y = q[i];
                h = s * g;
                g *= c;
                z = sqrt(f * f + h * h);
                e[i - 1] = z;
                c = f / z;
                s = h / z;
                f = x * c + g * s;
                g = -x * s + g * c;
                h = y * s;
                y *= c;
                if (with_v)
                    for (size_t j = 0; j < n; ++j) {
                        x = V(j, i - 1);
                        z = V(j, i);
                        V(j, i - 1) = x * c + z * s;
                        V(j, i) = -x * s + z * c;
                    }
                z = sqrt(f * f + h * h);
                q[i - 1] = z;
                if (z != 0) {
                    c = f / z;
                    s = h / z;
                }
                f = c * g + s * y;
                x = -s * g + c * y;
                if (mode != ZB_SVD_NO_U)
                    for (size_t j = 0; j < m; ++j) {
                        y = U(j, i - 1);
                        z = U(j, i);
                        U(j, i - 1) = y * c + z * s;
                        U(j, i) = -y * s + z * c;
                    }
            }
            e[l] = 0;
            e[k] = f;
            q[k] = x;
        }
    }
    // 5. selection-sort the singular values into descending order, permuting the vectors with them
    for (size_t i = 0; i < n; ++i) {
        size_t best = i;
        for (size_t j = i + 1; j < n; ++j)
            if (q[j] > q[best]) best = j;
        if (best == i) continue;
        { const T t_ = q[i]; q[i] = q[best]; q[best] = t_; }
        if (mode != ZB_SVD_NO_U)
            for (size_t r = 0; r < m; ++r) { const T t_ = U(r, i); U(r, i) = U(r, best); U(r, best) = t_; }
        if (with_v)
            for (size_t r = 0; r < n; ++r) { const T t_ = V(r, i); V(r, i) = V(r, best); V(r, best) = t_; }
    }
    return failed;
}


}  // namespace zb
