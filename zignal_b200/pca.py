"""Host-side mirror of zignal's Pca(T) (reference src/pca.zig:54-427).  The dense step -- the scaled
covariance X^T X / (n-1) (or Gram X X^T / (n-1)) and the batch projection -- runs on the GPU through
zb_gemm, and the SVD of the dim x dim (or n x n) matrix follows on the device (zb_svd_dev: one-sided Jacobi in a persistent
cooperative kernel) -- the covariance never leaves the GPU."""
from __future__ import annotations

import numpy as np

from ._ffi import ZignalError
from . import matrix


class Pca:
    def __init__(self, dtype=np.float64):  # Pca(T).init, pca.zig:78
        self.dtype = np.dtype(dtype)
        self.mean = np.zeros(0, self.dtype)
        self.components = np.zeros((0, 0), self.dtype)
        self.eigenvalues = np.zeros(0, self.dtype)
        self.num_components = 0
        self.dim = 0

    def fit(self, data: np.ndarray, num_components=None):  # pca.zig:104-164
        import torch
        data = np.ascontiguousarray(data, dtype=self.dtype)
        n, dim = data.shape
        if n == 0 or n == 1:
            raise ZignalError(12, "InsufficientData")
        if num_components is not None and num_components == 0:
            raise ZignalError(13, "InvalidComponents")
        max_components = min(n - 1, dim)
        k = min(num_components if num_components is not None else max_components, max_components)
        self.dim = dim
        x = torch.from_numpy(data).cuda()
        mean = torch.empty(dim, dtype=x.dtype, device=x.device)
        centered = torch.empty_like(x)
        matrix.center_columns(x, mean, True, centered)  # pca.zig:135-154 (f64-accumulated column means; tolerance-based)
        scale = 1.0 / float(n - 1)
        self.mean = mean.cpu().numpy()
        if n <= dim:  # Gram path, pca.zig:380-425
            g = matrix.gemm_device(centered, centered, False, True, scale, 0.0, None)
            ud, sd, _, conv = matrix.svd_device(g, True, False)
            if conv != 0:
                raise ZignalError(4, "SvdFailed")
            u, s = ud.cpu().numpy(), sd.cpu().numpy()
            comps = np.zeros((dim, k), self.dtype)
            cen = centered.cpu().numpy()
            for i in range(k):
                if s[i] > 1e-12:
                    comps[:, i] = (cen.T @ u[:, i]) / np.sqrt(s[i] * self.dtype.type(n - 1))
            self.components, self.eigenvalues = comps, s[:k].copy()
        else:  # covariance path, pca.zig:331-362
            cov = matrix.gemm_device(centered, centered, True, False, scale, 0.0, None)
            ud, sd, _, conv = matrix.svd_device(cov, True, False)
            if conv != 0:
                raise ZignalError(4, "SvdFailed")
            self.components, self.eigenvalues = np.ascontiguousarray(ud[:, :k].cpu().numpy()), sd[:k].cpu().numpy().copy()
        self.num_components = k

    def transform(self, data: np.ndarray) -> np.ndarray:  # pca.zig:291-312
        import torch
        if self.num_components == 0:
            raise ZignalError(5, "NotFitted")
        data = np.ascontiguousarray(data, dtype=self.dtype)
        if data.shape[1] != self.dim:
            raise ZignalError(1, "DimensionMismatch")
        x = torch.from_numpy(data).cuda()
        centered = torch.empty_like(x)
        matrix.center_columns(x, torch.from_numpy(self.mean).cuda(), False, centered)  # pca.zig:300-308
        comps = torch.from_numpy(np.ascontiguousarray(self.components)).cuda()
        return matrix.gemm_device(centered, comps, False, False, 1.0, 0.0, None).cpu().numpy()

    def project(self, vector: np.ndarray) -> np.ndarray:  # pca.zig:169-234
        if self.num_components == 0:
            raise ZignalError(5, "NotFitted")
        v = np.asarray(vector, self.dtype)
        if v.size != self.dim:
            raise ZignalError(1, "DimensionMismatch")
        return self.components.T @ (v - self.mean)

    def reconstruct(self, coefficients: np.ndarray) -> np.ndarray:  # pca.zig:238-287
        c = np.asarray(coefficients, self.dtype)
        if c.size != self.num_components:
            raise ZignalError(5, "InvalidCoefficients")
        return self.mean + self.components @ c
