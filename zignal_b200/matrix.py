"""Host-side mirror of zignal's Matrix(T) operations on the hot path: `gemm` (reference
src/matrix/Matrix.zig:696-822) runs on the device through the C ABI; `svd` (Matrix.zig:1570, svd.zig:80-496) and `eigh`
(matrix/eigen.zig:34) are parallel-ordering Jacobi methods (csrc/zb_jacobi.cu: a persistent cooperative kernel for 24 columns
and more, the same algorithm on the host below that)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._ffi import ZignalError, check, lib


def _torch():
    import torch
    return torch


def _ptr(t, ctype):
    return C.cast(t.data_ptr(), C.POINTER(ctype))


def gemm_device(a, b, trans_a=False, trans_b=False, alpha=1.0, beta=0.0, c=None):
    """C = alpha * op(A) * op(B) + beta * C on CUDA tensors (row-major, contiguous); returns a new CUDA tensor."""
    import torch
    assert a.is_cuda and b.is_cuda and a.dtype == b.dtype and a.is_contiguous() and b.is_contiguous()
    ar, ac = a.shape
    br, bc = b.shape
    m = ac if trans_a else ar
    n = br if trans_b else bc
    out = torch.empty((m, n), dtype=a.dtype, device=a.device)
    stream = torch.cuda.current_stream().cuda_stream
    if a.dtype == torch.float32:
        cp = _ptr(c, C.c_float) if c is not None else None
        check(lib().zb_gemm_f32(_ptr(a, C.c_float), ar, ac, int(trans_a), _ptr(b, C.c_float), br, bc, int(trans_b), C.c_float(alpha),
                                C.c_float(beta), cp, _ptr(out, C.c_float), stream))
    elif a.dtype == torch.float64:
        cp = _ptr(c, C.c_double) if c is not None else None
        check(lib().zb_gemm_f64(_ptr(a, C.c_double), ar, ac, int(trans_a), _ptr(b, C.c_double), br, bc, int(trans_b), C.c_double(alpha),
                                C.c_double(beta), cp, _ptr(out, C.c_double), stream))
    else:
        raise TypeError("gemm supports f32 and f64")
    return out


def center_columns(x, mean, compute_mean: bool, centered):
    """Pca centering on CUDA tensors: (optionally) mean[j] = column mean of x, centered = x - mean."""
    import torch
    n, dim = x.shape
    stream = torch.cuda.current_stream().cuda_stream
    if x.dtype == torch.float32:
        check(lib().zb_center_columns_f32(_ptr(x, C.c_float), n, dim, _ptr(mean, C.c_float), int(compute_mean), _ptr(centered, C.c_float), stream))
    else:
        check(lib().zb_center_columns_f64(_ptr(x, C.c_double), n, dim, _ptr(mean, C.c_double), int(compute_mean), _ptr(centered, C.c_double), stream))


def gemm(a: np.ndarray, b: np.ndarray, trans_a=False, trans_b=False, alpha=1.0, beta=0.0, c=None) -> np.ndarray:
    """numpy in / numpy out convenience wrapper around `gemm_device`."""
    import torch
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b, dtype=a.dtype)
    ta = torch.from_numpy(a).cuda()
    tb = torch.from_numpy(b).cuda()
    tc = torch.from_numpy(np.ascontiguousarray(c, dtype=a.dtype)).cuda() if c is not None else None
    return gemm_device(ta, tb, trans_a, trans_b, alpha, beta, tc).cpu().numpy()


_MODES = {"no_u": 0, "skinny_u": 1, "full_u": 2}


def svd(a: np.ndarray, mode: str = "full_u", with_v: bool = False):
    """Matrix.svd(allocator, .{ .with_v, .mode }) -> (u, s, v, converged).  rows >= cols required (svd.zig:86)."""
    a = np.ascontiguousarray(a)
    m, n = a.shape
    if m < n:
        raise ZignalError(1, "DimensionMismatch")
    u = np.zeros((m, m if mode == "full_u" else n), a.dtype) if mode != "no_u" else np.zeros((0, 0), a.dtype)
    s = np.zeros(n, a.dtype)
    v = np.zeros((n, n), a.dtype) if with_v else np.zeros((0, 0), a.dtype)
    conv = C.c_uint64(0)
    if a.dtype == np.float64:
        P = lambda x: x.ctypes.data_as(C.POINTER(C.c_double)) if x.size else None
        check(lib().zb_svd_f64(P(a), m, n, _MODES[mode], int(with_v), P(u), P(s), P(v), C.byref(conv)))
    elif a.dtype == np.float32:
        P = lambda x: x.ctypes.data_as(C.POINTER(C.c_float)) if x.size else None
        check(lib().zb_svd_f32(P(a), m, n, _MODES[mode], int(with_v), P(u), P(s), P(v), C.byref(conv)))
    else:
        raise TypeError("svd supports f32 and f64")
    return u, s, v, conv.value


def svd_device(a, with_u: bool = True, with_v: bool = False):
    """The same decomposition for a CUDA tensor (m x n, m >= n, f32 / f64) that stays on the device (zb_svd_dev_*):
    -> (u: m x n or None, s: n descending, v: n x n or None, converged) as CUDA tensors."""
    torch = _torch()
    assert a.is_cuda and a.is_contiguous() and a.dim() == 2
    m, n = a.shape
    if m < n:
        raise ZignalError(1, "DimensionMismatch")
    u = torch.empty((m, n), dtype=a.dtype, device=a.device) if with_u else None
    s = torch.empty(n, dtype=a.dtype, device=a.device)
    v = torch.empty((n, n), dtype=a.dtype, device=a.device) if with_v else None
    conv = C.c_uint64(0)
    ct = C.c_double if a.dtype == torch.float64 else C.c_float
    fn = lib().zb_svd_dev_f64 if a.dtype == torch.float64 else lib().zb_svd_dev_f32
    with torch.cuda.device(a.device):
        check(fn(_ptr(a, ct), m, n, _ptr(u, ct) if u is not None else None, _ptr(s, ct), _ptr(v, ct) if v is not None else None,
                 C.byref(conv), torch.cuda.current_stream(a.device).cuda_stream))
    return u, s, v, conv.value


def eigh(a: np.ndarray):
    """Matrix.eigh(allocator) (matrix/eigen.zig:34-136) -> (values ascending, vectors with eigenvectors as columns)."""
    a = np.ascontiguousarray(a)
    if a.ndim != 2:
        raise ZignalError(18, "NotSquare")
    rows, cols = a.shape
    values = np.zeros(rows, a.dtype)
    vectors = np.zeros((rows, rows), a.dtype)
    if a.dtype == np.float64:
        P = lambda x: x.ctypes.data_as(C.POINTER(C.c_double))
        check(lib().zb_eigh_f64(P(a), rows, cols, P(values), P(vectors)))
    elif a.dtype == np.float32:
        P = lambda x: x.ctypes.data_as(C.POINTER(C.c_float))
        check(lib().zb_eigh_f32(P(a), rows, cols, P(values), P(vectors)))
    else:
        raise TypeError("eigh supports f32 and f64")
    return values, vectors
