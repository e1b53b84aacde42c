"""Matrix.svd / Matrix.eigh on the GPU (csrc/zb_jacobi.cu: one-sided / two-sided Jacobi with a round-robin pair ordering in a
persistent cooperative kernel) against the oracle's restatement of the reference's sequential routines (svd.zig:149-496,
eigen.zig:34-136), by the reference's own criterion (test_svd_comparison.zig:51-72): singular / eigen values to sqrt(eps) -- in
fact to a few ulps of |A| --, orthonormal factors, reconstruction, and vectors up to sign where the spectrum is separated."""
import numpy as np
import pytest

import oracle_lib as zo
from test_abi import _check_svd_against_oracle, _svd_prod

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def zb():
    import torch
    assert torch.cuda.is_available()
    import zignal_b200 as zb
    return zb


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(24, 24), (40, 33), (64, 64), (300, 47), (256, 256), (3000, 40)])
@pytest.mark.parametrize("mode", ["no_u", "skinny_u", "full_u"])
def test_svd_device_kernel_vs_oracle(zb, dtype, shape, mode):
    if mode == "full_u" and shape[0] > 100:
        pytest.skip("full U of a tall matrix: host completion, covered at (40, 33)")
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    a = (rng.standard_normal(shape) @ np.diag(np.logspace(0, -3, shape[1]))).astype(dtype)    # graded spectrum
    u, s, v, conv = _svd_prod(a, mode, True)
    # from 32 columns on, a problem that fits one cluster's shared memory runs there (hardware cluster barrier per round)
    assert conv == 0 and zb.lib().zb_last_kernel().decode() == ("jacobi_svd_cluster" if shape[1] >= 32 else "jacobi_svd_onesided")
    assert zb.lib().zb_last_sweeps() <= 20, zb.lib().zb_last_sweeps()          # quadratic convergence, not the sweep limit
    _check_svd_against_oracle(a, u, s, v, mode)
    if shape[1] >= 32 and mode == "skinny_u":   # the cooperative kernel (global-memory barrier) runs the same arithmetic in the same order
        zb.lib().zb_tune(b"jacobi.cluster", 0)
        try:
            u2, s2, v2, conv2 = _svd_prod(a, mode, True)
            assert conv2 == 0 and zb.lib().zb_last_kernel().decode() == "jacobi_svd_onesided"
        finally:
            zb.lib().zb_tune(b"jacobi.cluster", 1)
        if shape[0] <= 2048:     # both are warp-per-pair kernels: bit for bit
            assert np.array_equal(s, s2) and np.array_equal(v, v2) and np.array_equal(u, u2)
        else:                    # the cooperative path takes a CTA per pair there (another summation order)
            assert np.allclose(s, s2, rtol=0, atol=64 * np.finfo(dtype).eps * s[0])


def test_svd_rank_deficient_and_clustered(zb):
    rng = np.random.default_rng(1)
    b = rng.standard_normal((60, 5))
    a = b @ rng.standard_normal((5, 40))                                    # rank 5 of 40
    u, s, v, conv = _svd_prod(a, "skinny_u", True)
    assert conv == 0
    assert np.count_nonzero(s > 1e-10 * s[0]) == 5
    assert np.allclose(u @ np.diag(s) @ v.T, a, atol=1e-11 * s[0])
    assert np.allclose(u.T @ u, np.eye(40), atol=1e-10) and np.allclose(v.T @ v, np.eye(40), atol=1e-10)
    q, _ = np.linalg.qr(rng.standard_normal((32, 32)))                      # all singular values equal: any basis is valid
    u, s, v, conv = _svd_prod(q, "skinny_u", True)
    assert conv == 0 and np.allclose(s, 1.0, atol=1e-13) and np.allclose(u @ np.diag(s) @ v.T, q, atol=1e-12)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_svd_device_pointer_entry(zb, dtype):
    """zb_svd_dev_*: the matrix and the factors stay on the device (Pca.fit's covariance path)."""
    import torch
    from zignal_b200 import matrix
    rng = np.random.default_rng(3)
    x = rng.standard_normal((500, 96)).astype(dtype)
    cov = (x.T @ x / 499).astype(dtype)
    u, s, v, conv = matrix.svd_device(torch.from_numpy(cov).cuda(), True, True)
    assert conv == 0
    u, s, v = u.cpu().numpy(), s.cpu().numpy(), v.cpu().numpy()
    _check_svd_against_oracle(cov, u, s, v, "skinny_u")
    ev = np.linalg.eigvalsh(cov.astype(np.float64))[::-1]
    assert np.allclose(s, ev, atol=(1e-12 if dtype == np.float64 else 2e-5) * ev[0])


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [24, 33, 64, 200])
def test_eigh_device_kernel_vs_oracle(zb, dtype, n):
    rng = np.random.default_rng(n)
    m = rng.normal(size=(n, n))
    a = ((m + m.T) * 0.5).astype(dtype)
    a = ((a + a.T) * dtype(0.5)).astype(dtype)
    vals, vecs = zb.matrix.eigh(a)
    assert zb.lib().zb_last_kernel().decode() == "jacobi_eigh_twosided"
    assert zb.lib().zb_last_sweeps() <= 20, zb.lib().zb_last_sweeps()
    ovals, ovecs = zo.eigh(a)
    eps = np.finfo(dtype).eps
    norm = float(np.linalg.norm(a.astype(np.float64), 2))
    # two Jacobi variants in the same precision: each is backward stable to O(n eps |A|), so that is how far apart they may be
    assert np.max(np.abs(vals.astype(np.float64) - ovals.astype(np.float64))) <= 8 * n * eps * norm
    v64 = vecs.astype(np.float64)
    assert np.allclose(v64.T @ v64, np.eye(n), atol=8 * n * eps)
    assert np.allclose(v64 @ np.diag(vals.astype(np.float64)) @ v64.T, a, atol=8 * n * eps * norm)
    assert np.all(np.diff(vals) >= 0)
    for i in range(n):
        gap = min(abs(ovals[i] - ovals[j]) for j in range(n) if j != i)
        if gap > 0.05 * norm / n:
            assert abs(abs(float(v64[:, i] @ ovecs[:, i].astype(np.float64))) - 1.0) <= (1e-9 if dtype == np.float64 else 2e-2)


def test_pca_fit_end_to_end_at_baseline_size(zb):
    """Pca.fit at the PCA-GEMM configuration (n = 1,048,576 x dim 256 f32): centring, tcgen05 X^T X, device SVD.  The eigenvalues
    must match an f64 eigendecomposition of the f64 covariance; the leading component must span the planted direction."""
    import time

    import torch
    from zignal_b200 import matrix
    g = torch.Generator(device="cuda").manual_seed(7)
    n, dim = 1 << 20, 256
    x = torch.randn(n, dim, device="cuda", generator=g)
    x[:, 0] *= 3.0                                                     # a planted dominant axis
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mean = torch.empty(dim, device="cuda")
    centered = torch.empty_like(x)
    matrix.center_columns(x, mean, True, centered)
    cov = matrix.gemm_device(centered, centered, True, False, 1.0 / (n - 1), 0.0, None)
    u, s, _, conv = matrix.svd_device(cov, True, False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert conv == 0
    ref = torch.zeros(dim, dim, dtype=torch.float64, device="cuda")
    m64 = x.double().mean(0)
    for i in range(0, n, 131072):
        blk = x[i:i + 131072].double() - m64
        ref += blk.T @ blk
    ref /= (n - 1)
    ev = torch.linalg.eigvalsh(ref).flip(0)
    assert ((s.double() - ev).abs().max() / ev[0]).item() <= 1e-5
    assert abs(abs(float(u[0, 0])) - 1.0) < 1e-3
    print(f"Pca.fit core (center + X^T X + SVD) at 1M x 256 f32: {dt * 1e3:.2f} ms")
