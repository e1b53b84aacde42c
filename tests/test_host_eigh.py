"""Matrix.eigh (reference matrix/eigen.zig): the reference's own tests (:138-296) against the oracle and against the
product's host routine, and bit-identity of the two on random symmetric matrices.  Host code only: runs without a GPU."""
import numpy as np
import pytest

import oracle_lib as zo


@pytest.fixture(scope="module")
def zb():
    import zignal_b200 as zb
    return zb


def _impls(zb):
    def prod(a):
        return zb.matrix.eigh(a)
    return [("oracle", zo.eigh), ("product", prod)]


def test_eigh_reference_cases(zb):
    for name, eigh in _impls(zb):
        vals, vecs = eigh(np.array([[2.0, 1.0], [1.0, 2.0]]))
        assert abs(vals[0] - 1) < 1e-9 and abs(vals[1] - 3) < 1e-9, name
        assert np.allclose((vecs * vecs).sum(axis=0), 1.0, atol=1e-9)
        vals, _ = eigh(np.array([[0.0, 1.0], [1.0, 0.0]]))                       # indefinite: signs survive
        assert abs(vals[0] + 1) < 1e-9 and abs(vals[1] - 1) < 1e-9
        vals, _ = eigh(np.array([[2e8, 1e8], [1e8, 2e8]]))                       # scale invariant
        assert abs(vals[0] - 1e8) < 1.0 and abs(vals[1] - 3e8) < 1.0
        vals, vecs = eigh(np.array([[-2.0]]))
        assert abs(vals[0] + 2) < 1e-12 and abs(vecs[0, 0] - 1) < 1e-12
        a = np.array([[4.0, 1.0, -2.0], [1.0, 2.0, 0.0], [-2.0, 0.0, 3.0]])
        vals, vecs = eigh(a)
        assert np.allclose(a @ vecs, vecs * vals[None, :], atol=1e-9)             # every column is a matching eigenpair
        assert np.allclose(vecs @ np.diag(vals) @ vecs.T, a, atol=1e-9)           # A = V diag(lambda) V^T
        vals, vecs = eigh(np.diag([5.0, -3.0, 2.0]))                              # no rotations: only the sort permutes the basis
        assert np.allclose(vals, [-3, 2, 5], atol=1e-12)
        assert abs(abs(vecs[1, 0]) - 1) < 1e-12 and abs(abs(vecs[2, 1]) - 1) < 1e-12 and abs(abs(vecs[0, 2]) - 1) < 1e-12
        off = np.float32(1e-20)                                                   # huge Jacobi theta in f32 stays finite
        a32 = np.array([[1, off, 0.5], [off, 2, 0], [0.5, 0, 3]], np.float32)
        vals, _ = eigh(a32)
        assert np.all(np.isfinite(vals)) and np.allclose(vals, [0.881966, 2.0, 3.118034], atol=1e-4)


def test_eigh_errors(zb):
    with pytest.raises(zo.OracleStatus) as ei:
        zo.eigh(np.array([[0.0, 1.0], [2.0, 0.0]]))
    assert ei.value.status == 19
    with pytest.raises(zo.OracleStatus) as ei:
        zo.eigh(np.zeros((2, 3)))
    assert ei.value.status == 18
    for bad in (np.array([[np.nan, 1.0], [1.0, 2.0]]), np.array([[np.inf, 0.0], [0.0, 1.0]])):
        with pytest.raises(zo.OracleStatus) as ei:
            zo.eigh(bad)
        assert ei.value.status == 20
    for bad, name in [(np.array([[0.0, 1.0], [2.0, 0.0]]), "NotSymmetric"), (np.zeros((2, 3)), "NotSquare"),
                      (np.array([[np.nan, 1.0], [1.0, 2.0]]), "NotFinite"), (np.array([[np.inf, 0.0], [0.0, 1.0]]), "NotFinite")]:
        with pytest.raises(zb.ZignalError) as ei:
            zb.matrix.eigh(bad)
        assert ei.value.name == name
    a = np.array([[1.0, 2.0], [2.0 + 1e-9, 1.0]])                                 # rounding-level asymmetry passes (:39-41)
    zo.eigh(a)
    zb.matrix.eigh(a)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [1, 2, 3, 7, 16, 23])
def test_eigh_product_agrees_with_the_oracle(zb, dtype, n):
    """The product's parallel-ordering Jacobi against the oracle's cyclic Jacobi (eigen.zig:34-136): the same eigenvalues to
    eps * |A|, orthonormal vectors that diagonalise A, and -- for separated eigenvalues -- the same vectors up to sign.
    (n < 24 runs on the host; larger matrices take the GPU kernel, tests/test_gpu_linalg.py.)"""
    rng = np.random.default_rng(n)
    m = rng.normal(size=(n, n))
    a = ((m + m.T) * 0.5).astype(dtype)
    a = ((a + a.T) * dtype(0.5)).astype(dtype)
    vals, vecs = zb.matrix.eigh(a)
    ovals, ovecs = zo.eigh(a)
    eps = np.finfo(dtype).eps
    norm = max(float(np.abs(a).max()) * n, 1e-300)
    assert np.max(np.abs(vals.astype(np.float64) - ovals.astype(np.float64))) <= 64 * eps * norm
    v64 = vecs.astype(np.float64)
    assert np.allclose(v64.T @ v64, np.eye(n), atol=200 * eps * n)
    assert np.allclose(v64 @ np.diag(vals.astype(np.float64)) @ v64.T, a, atol=200 * eps * norm)
    assert np.allclose(np.sort(np.linalg.eigvalsh(a.astype(np.float64))), vals, atol=200 * eps * norm)
    assert np.all(np.diff(vals) >= 0)
    for i in range(n):
        gap = min(abs(ovals[i] - ovals[j]) for j in range(n) if j != i) if n > 1 else 1.0
        if gap > 1e-2 * norm / n:
            assert abs(abs(float(v64[:, i] @ ovecs[:, i].astype(np.float64))) - 1.0) <= (1e-9 if dtype == np.float64 else 5e-3)
