// zo_linalg.cpp -- CPU oracle for Matrix/SMatrix SVD and Matrix.gemm.
// TEST INFRASTRUCTURE ONLY (see zignal_oracle.h).  Restates matrix/svd.zig:149-496 (the shared
// Golub-Reinsch kernel: Householder bidiagonalisation :200-264, V accumulation :267-293,
// U accumulation :296-338, implicit-shift QR sweeps :341-468, descending sort :470-493) and
// matrix/Matrix.zig:631-684 (simdGemmKernel), :696-822 (gemm).
#include "zo_linalg.h"

namespace zo {

template <typename T>
int64_t svd_kernel(const T* a_in, uint32_t m_, uint32_t n_, int mode, bool with_v, T* u_, uint32_t u_cols_, T* q, T* v_) {
    const size_t m = m_, n = n_, ucols = u_cols_;
    const size_t urows = m;
    auto A = [&](size_t i, size_t j) -> const T& { return a_in[i * n + j]; };
    auto U = [&](size_t i, size_t j) -> T& { return u_[i * ucols + j]; };
    auto V = [&](size_t i, size_t j) -> T& { return v_[i * n + j]; };
    std::vector<T> e(n, (T)0);
    const size_t max_iterations = 300;
    T eps = std::numeric_limits<T>::epsilon();
    const T tol = std::numeric_limits<T>::min() / eps;

    size_t l = 0;
    int64_t retval = 0;
    T c = 0, f = 0, g = 0, h = 0, s = 0, x = 0, y = 0, z = 0;

    for (size_t i = 0; i < m; ++i)
        for (size_t j = 0; j < n; ++j) U(i, j) = A(i, j);

    // Householder's reduction to bidiagonal form.  svd.zig:200-264
    g = 0;
    x = 0;
    for (size_t i = 0; i < n; ++i) {
        e[i] = g;
        s = 0;
        l = i + 1;
        for (size_t j = i; j < m; ++j) s += U(j, i) * U(j, i);
        if (s < tol) {
            g = 0;
        } else {
            f = U(i, i);
            g = (f < 0) ? std::sqrt(s) : -std::sqrt(s);
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
            g = (f < 0) ? std::sqrt(s) : -std::sqrt(s);
            h = f * g - s;
            U(i, i + 1) = f - g;
            for (size_t j = l; j < n; ++j) e[j] = U(i, j) / h;
            for (size_t j = l; j < m; ++j) {
                s = 0;
                for (size_t k = l; k < n; ++k) s += U(j, k) * U(i, k);
                for (size_t k = l; k < n; ++k) U(j, k) += s * e[k];
            }
        }
        y = std::fabs(q[i]) + std::fabs(e[i]);
        x = std::max(x, y);
    }

    // Accumulation of right-hand transformations.  :267-293
    if (with_v) {
        for (size_t ri = 0; ri < n; ++ri) {
            const size_t i = n - 1 - ri;
            if (g != 0) {
                h = U(i, i + 1) * g;
                for (size_t j = l; j < n; ++j) V(j, i) = U(i, j) / h;
                for (size_t j = l; j < n; ++j) {
                    s = 0;
                    for (size_t k = l; k < n; ++k) s += U(i, k) * V(k, j);
                    for (size_t k = l; k < n; ++k) V(k, j) += s * V(k, i);
                }
            }
            for (size_t j = l; j < n; ++j) { V(i, j) = 0; V(j, i) = 0; }
            V(i, i) = 1;
            g = e[i];
            l = i;
        }
    }

    // Accumulation of left-hand transformations.  :296-338
    if (mode != ZO_SVD_NO_U) {
        for (size_t i = n; i < urows; ++i) {
            for (size_t j = n; j < ucols; ++j) U(i, j) = 0;
            if (i < ucols) U(i, i) = 1;
        }
        for (size_t ri = 0; ri < n; ++ri) {
            const size_t i = n - 1 - ri;
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

    // Diagonalisation of the bidiagonal form.  :341-468 (labeled-switch state machine -> gotos)
    eps *= x;
    for (size_t rk = 0; rk < n; ++rk) {
        const size_t k = n - 1 - rk;
        size_t iter = 0;
    test_splitting:
        for (size_t rl = 0; rl < k + 1; ++rl) {
            l = k - rl;
            if (std::fabs(e[l]) <= eps) goto test_convergence;
            if (std::fabs(q[l - 1]) <= eps) goto cancellation;  // l >= 1 here: e[0] == 0 always
        }
        goto test_convergence;
    cancellation : {
        c = 0;
        s = 1;
        const size_t l1 = l - 1;
        for (size_t i = l; i < k + 1; ++i) {
            f = s * e[i];
            e[i] *= c;
            if (std::fabs(f) <= eps) goto test_convergence;
            g = q[i];
            h = std::sqrt(f * f + g * g);
            q[i] = h;
            c = g / h;
            s = -f / h;
            if (mode != ZO_SVD_NO_U) {
                for (size_t j = 0; j < m; ++j) {
                    y = U(j, l1);
                    z = U(j, i);
                    U(j, l1) = y * c + z * s;
                    U(j, i) = -y * s + z * c;
                }
            }
        }
    }
    test_convergence:
        z = q[k];
        if (l == k) goto convergence_check;
        iter += 1;
        if (iter > max_iterations) { retval = (int64_t)k; continue; }
        x = q[l];
        y = q[k - 1];
        g = e[k - 1];
        h = e[k];
        f = ((y - z) * (y + z) + (g - h) * (g + h)) / (2 * h * y);
        g = std::sqrt(f * f + (T)1.0);
        f = ((x - z) * (x + z) + h * (y / ((f < 0) ? (f - g) : (f + g)) - h)) / x;
        c = 1;
        s = 1;
        for (size_t i = l + 1; i < k + 1; ++i) {
            g = e[i];
            y = q[i];
            h = s * g;
            g *= c;
            z = std::sqrt(f * f + h * h);
            e[i - 1] = z;
            c = f / z;
            s = h / z;
            f = x * c + g * s;
            g = -x * s + g * c;
            h = y * s;
            y *= c;
            if (with_v) {
                for (size_t j = 0; j < n; ++j) {
                    x = V(j, i - 1);
                    z = V(j, i);
                    V(j, i - 1) = x * c + z * s;
                    V(j, i) = -x * s + z * c;
                }
            }
            z = std::sqrt(f * f + h * h);
            q[i - 1] = z;
            if (z != 0) {
                c = f / z;
                s = h / z;
            }
            f = c * g + s * y;
            x = -s * g + c * y;
            if (mode != ZO_SVD_NO_U) {
                for (size_t j = 0; j < m; ++j) {
                    y = U(j, i - 1);
                    z = U(j, i);
                    U(j, i - 1) = y * c + z * s;
                    U(j, i) = -y * s + z * c;
                }
            }
        }
        e[l] = 0;
        e[k] = f;
        q[k] = x;
        goto test_splitting;
    convergence_check:
        if (z < 0) {
            q[k] = -z;
            if (with_v)
                for (size_t j = 0; j < n; ++j) V(j, k) = -V(j, k);
        }
    }

    // Sort singular values in descending order.  :470-493
    for (size_t i = 0; i < n; ++i) {
        size_t max_idx = i;
        T max_val = q[i];
        for (size_t j = i + 1; j < n; ++j)
            if (q[j] > max_val) { max_idx = j; max_val = q[j]; }
        if (max_idx != i) {
            std::swap(q[i], q[max_idx]);
            if (mode != ZO_SVD_NO_U)
                for (size_t row = 0; row < m; ++row) std::swap(U(row, i), U(row, max_idx));
            if (with_v)
                for (size_t row = 0; row < n; ++row) std::swap(V(row, i), V(row, max_idx));
        }
    }
    return retval;
}

template int64_t svd_kernel<float>(const float*, uint32_t, uint32_t, int, bool, float*, uint32_t, float*, float*);
template int64_t svd_kernel<double>(const double*, uint32_t, uint32_t, int, bool, double*, uint32_t, double*, double*);

// svd.zig:80-115 (dynamic wrapper): allocates u as m x m (full) or m x n, runs the kernel.
template <typename T>
int64_t svd(const T* a, uint32_t m, uint32_t n, int mode, bool with_v, T* u, T* s, T* v) {
    const uint32_t ucols = (mode == ZO_SVD_FULL_U) ? m : n;
    std::vector<T> ubuf((size_t)m * ucols, (T)0), vbuf(with_v ? (size_t)n * n : 1, (T)0);
    for (uint32_t i = 0; i < n; ++i) s[i] = 0;
    const int64_t rc = svd_kernel<T>(a, m, n, mode, with_v, ubuf.data(), ucols, s, vbuf.data());
    if (mode != ZO_SVD_NO_U && u) std::memcpy(u, ubuf.data(), ubuf.size() * sizeof(T));
    if (with_v && v) std::memcpy(v, vbuf.data(), (size_t)n * n * sizeof(T));
    return rc;
}
template int64_t svd<float>(const float*, uint32_t, uint32_t, int, bool, float*, float*, float*);
template int64_t svd<double>(const double*, uint32_t, uint32_t, int, bool, double*, double*, double*);

// Matrix.zig:631-684 simdGemmKernel: both operands row-contiguous with row length a_cols.
// @reduce(.Add) without reassociation is an in-order sum of the lanes.
template <typename T>
static void simd_gemm_kernel(T* result, const T* ma, const T* mb, T alpha, size_t a_rows, size_t a_cols, size_t b_cols, size_t vec_len) {
#pragma omp parallel for schedule(static) num_threads(g_threads) if (g_threads > 1)
    for (size_t i = 0; i < a_rows; ++i) {
        const T* arow = ma + i * a_cols;
        for (size_t j = 0; j < b_cols; ++j) {
            const T* brow = mb + j * a_cols;
            T accumulator = 0;
            size_t k = 0;
            for (; k + vec_len <= a_cols; k += vec_len) {
                T red = arow[k] * brow[k];
                for (size_t v = 1; v < vec_len; ++v) red += arow[k + v] * brow[k + v];
                accumulator += red;
            }
            for (; k < a_cols; ++k) accumulator += arow[k] * brow[k];
            result[i * b_cols + j] += alpha * accumulator;
        }
    }
}

// Matrix.zig:696-822
template <typename T>
int gemm(const T* a, uint32_t ar, uint32_t ac, bool trans_a, const T* b, uint32_t br, uint32_t bc, bool trans_b, T alpha,
         T beta, const T* c, T* out, int vec_len_in) {
    const size_t a_rows = trans_a ? ac : ar, a_cols = trans_a ? ar : ac;
    const size_t b_rows = trans_b ? bc : br, b_cols = trans_b ? br : bc;
    if (a_cols != b_rows) return ZO_ERR_DIMENSION_MISMATCH;
    auto A = [&](size_t i, size_t j) { return a[i * ac + j]; };
    auto B = [&](size_t i, size_t j) { return b[i * bc + j]; };
    if (c != nullptr && beta != 0) {
        for (size_t i = 0; i < a_rows * b_cols; ++i) out[i] = beta * c[i];
    } else {
        for (size_t i = 0; i < a_rows * b_cols; ++i) out[i] = 0;
    }
    if (alpha == 0) return ZO_OK;
    const size_t vec_len = vec_len_in > 0 ? (size_t)vec_len_in : (sizeof(T) == 4 ? 8 : 4);
    const size_t total_ops = a_rows * a_cols * b_cols;
    if (vec_len > 1 && total_ops >= 512) {
        std::vector<T> at, bt;
        const T* pa = a;
        const T* pb = b;
        if (trans_a) {  // a_transposed[i,k] = A[k,i]   (:765-772, :793-799)
            at.resize(a_rows * a_cols);
            for (size_t k = 0; k < a_cols; ++k)
                for (size_t i = 0; i < a_rows; ++i) at[i * a_cols + k] = A(k, i);
            pa = at.data();
        }
        if (!trans_b) {  // b_transposed[j,k] = B[k,j]  (:755-761, :779-785)
            if (trans_a && a == b) {
                pb = at.data();  // covariance special case :774-776
            } else {
                bt.resize(b_cols * a_cols);
                for (size_t k = 0; k < a_cols; ++k)
                    for (size_t j = 0; j < b_cols; ++j) bt[j * a_cols + k] = B(k, j);
                pb = bt.data();
            }
        }
        simd_gemm_kernel<T>(out, pa, pb, alpha, a_rows, a_cols, b_cols, vec_len);
    } else {  // :806-817
        for (size_t i = 0; i < a_rows; ++i)
            for (size_t j = 0; j < b_cols; ++j) {
                T accumulator = 0;
                for (size_t k = 0; k < a_cols; ++k) {
                    const T av = trans_a ? A(k, i) : A(i, k);
                    const T bv = trans_b ? B(j, k) : B(k, j);
                    accumulator += av * bv;
                }
                out[i * b_cols + j] += alpha * accumulator;
            }
    }
    return ZO_OK;
}
template int gemm<float>(const float*, uint32_t, uint32_t, bool, const float*, uint32_t, uint32_t, bool, float, float, const float*, float*, int);
template int gemm<double>(const double*, uint32_t, uint32_t, bool, const double*, uint32_t, uint32_t, bool, double, double, const double*, double*, int);

}  // namespace zo

extern "C" {

int64_t zo_svd_f64(const double* a, uint32_t m, uint32_t n, int mode, int with_v, double* u, double* s, double* v) {
    if (m < n) return -1;  // svd.zig:86 assert(rows >= cols)
    return zo::svd<double>(a, m, n, mode, with_v != 0, u, s, v);
}
int64_t zo_svd_f32(const float* a, uint32_t m, uint32_t n, int mode, int with_v, float* u, float* s, float* v) {
    if (m < n) return -1;
    return zo::svd<float>(a, m, n, mode, with_v != 0, u, s, v);
}
int zo_gemm_f32(const float* a, uint32_t ar, uint32_t ac, int ta, const float* b, uint32_t br, uint32_t bc, int tb,
                float alpha, float beta, const float* c, float* out, int vec_len) {
    return zo::gemm<float>(a, ar, ac, ta != 0, b, br, bc, tb != 0, alpha, beta, c, out, vec_len);
}
int zo_gemm_f64(const double* a, uint32_t ar, uint32_t ac, int ta, const double* b, uint32_t br, uint32_t bc, int tb,
                double alpha, double beta, const double* c, double* out, int vec_len) {
    return zo::gemm<double>(a, ar, ac, ta != 0, b, br, bc, tb != 0, alpha, beta, c, out, vec_len);
}

}  // extern "C"
