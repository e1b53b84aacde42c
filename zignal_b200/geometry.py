"""Host-side mirror of zignal's geometric transforms (reference src/geometry/transforms.zig:10-292).

`project()` of these transforms is what `Image.warp` evaluates per pixel on the GPU (zb_warp); the fits
(`find`) stay on the host, in f64, and use the Golub-Reinsch SVD through the C ABI (zb_svd_f64) exactly
where the reference does (similarity: 2x2 SVD :75; affine: SVD-based pseudo-inverse Matrix.zig:447-509;
projective: 9x9 SVD :283, or the exact 8x8 solve for four correspondences :248-268)."""
from __future__ import annotations

import numpy as np

from . import matrix
from ._ffi import ZignalError


class RankDeficient(ZignalError):
    def __init__(self):
        super().__init__(5, "RankDeficient")


class NotConverged(ZignalError):
    def __init__(self):
        super().__init__(4, "NotConverged")


def _pts(p):
    a = np.asarray(p, dtype=np.float64)
    assert a.ndim == 2 and a.shape[1] == 2
    return a


class _Linear:
    kind = 1

    def __init__(self, matrix2x2=None, bias=None):
        self.matrix = np.eye(2) if matrix2x2 is None else np.asarray(matrix2x2, np.float64).reshape(2, 2)
        self.bias = np.zeros(2) if bias is None else np.asarray(bias, np.float64).reshape(2)

    def project(self, point):  # transforms.zig:39-42 / :147-150
        p = np.asarray(point, np.float64)
        return self.matrix @ p + self.bias

    def as_f32(self):
        """(zb_xform kind, 6 floats {m00, m01, m10, m11, b0, b1}) -- `transform.as(f32)` of the reference's front-ends."""
        return self.kind, np.array([self.matrix[0, 0], self.matrix[0, 1], self.matrix[1, 0], self.matrix[1, 1], self.bias[0], self.bias[1]],
                                   dtype=np.float32)


class SimilarityTransform(_Linear):
    kind = 0

    def __init__(self, from_points=None, to_points=None):
        super().__init__()
        if from_points is not None:
            self.find(from_points, to_points)

    def find(self, from_points, to_points):  # transforms.zig:47-112
        f, t = _pts(from_points), _pts(to_points)
        assert len(f) >= 2 and len(f) == len(t)
        n = float(len(f))
        mean_from, mean_to = f.sum(0) * (1.0 / n), t.sum(0) * (1.0 / n)
        fc, tc = f - mean_from, t - mean_to
        sigma_from = float((fc * fc).sum()) / n
        cov = (tc.T @ fc) * (1.0 / n)
        det_cov = cov[0, 0] * cov[1, 1] - cov[0, 1] * cov[1, 0]
        u, s, v, conv = matrix.svd(cov, "skinny_u", True)
        if conv != 0:
            raise NotConverged()
        tol = s[0] * np.finfo(np.float64).eps * 2.0
        if int(np.count_nonzero(s > tol)) == 0:
            raise RankDeficient()
        det_u = u[0, 0] * u[1, 1] - u[0, 1] * u[1, 0]
        det_v = v[0, 0] * v[1, 1] - v[0, 1] * v[1, 0]
        smat = np.eye(2)
        if det_cov < 0 or (det_cov == 0 and det_u * det_v < 0):
            if s[1] < s[0]:
                smat[1, 1] = -1
            else:
                smat[0, 0] = -1
        r = u @ (smat @ v.T)
        c = 1.0
        if sigma_from != 0:
            c = 1.0 / sigma_from * float(np.trace(np.diag(s) @ smat))
        self.matrix = r * c
        self.bias = mean_to + (r @ mean_from) * (-c)


def pinv(a: np.ndarray, tolerance=None):
    """Matrix.pinv (Matrix.zig:447-509): SVD-based Moore-Penrose pseudo-inverse; returns (pinv, effective_rank)."""
    a = np.asarray(a, np.float64)
    if a.shape[0] == 0 or a.shape[1] == 0:
        raise ZignalError(1, "DimensionMismatch")
    if a.shape[0] < a.shape[1]:
        p, rank = pinv(np.ascontiguousarray(a.T), tolerance)
        return np.ascontiguousarray(p.T), rank
    u, s, v, conv = matrix.svd(a, "skinny_u", True)
    if conv != 0:
        raise NotConverged()
    sigma_max = s[0] if s.size else 0.0
    if sigma_max == 0:
        return np.zeros((a.shape[1], a.shape[0])), 0
    tol = sigma_max * max(a.shape) * np.finfo(np.float64).eps if tolerance is None else tolerance
    inv = np.where(s > tol, 1.0 / np.where(s > tol, s, 1.0), 0.0)
    return (v * inv[None, :]) @ u.T, int(np.count_nonzero(s > tol))


class AffineTransform(_Linear):
    kind = 1

    def __init__(self, from_points=None, to_points=None):
        super().__init__()
        if from_points is not None:
            self.find(from_points, to_points)

    def find(self, from_points, to_points):  # transforms.zig:155-191
        f, t = _pts(from_points), _pts(to_points)
        assert len(f) == len(t) and len(f) >= 3
        p = np.vstack([f.T, np.ones(len(f))])  # 3 x N
        q = t.T                                # 2 x N
        pi, rank = pinv(p)
        if rank < 3:
            raise RankDeficient()
        m = q @ pi
        self.matrix = m[:, :2].copy()
        self.bias = m[:, 2].copy()


def _orientation_is_collinear(a, b, c):  # Point.zig:200-214
    swap = (b[0] > c[0]) or (b[0] == c[0] and b[1] > c[1])
    p1, p2 = (c, b) if swap else (b, c)
    u = a[0] * (p1[1] - p2[1]) + p1[0] * (p2[1] - a[1]) + p2[0] * (a[1] - p1[1])
    return u == 0


def are_all_collinear(points) -> bool:  # Point.zig:232-253
    pts = _pts(points)
    if len(pts) < 3:
        return True
    p1 = pts[0]
    i = 1
    while i < len(pts) and np.array_equal(pts[i], p1):
        i += 1
    if i == len(pts):
        return True
    p2 = pts[i]
    return all(_orientation_is_collinear(p1, p2, p) for p in pts[i + 1:])


class ProjectiveTransform:
    kind = 2

    def __init__(self, from_points=None, to_points=None):
        self.matrix = np.eye(3)
        if from_points is not None:
            self.find(from_points, to_points)

    def project(self, point):  # transforms.zig:224-231
        p = np.asarray(point, np.float64)
        d = self.matrix @ np.array([p[0], p[1], 1.0])
        if d[2] != 0:
            d = d * (1 / d[2])
        return d[:2]

    def inv(self):  # transforms.zig:234-236
        if abs(np.linalg.det(self.matrix)) == 0:
            return None
        out = ProjectiveTransform()
        out.matrix = np.linalg.inv(self.matrix)
        return out

    def as_f32(self):
        return self.kind, np.ascontiguousarray(self.matrix, dtype=np.float32).ravel()

    def find(self, from_points, to_points):  # transforms.zig:242-290
        f, t = _pts(from_points), _pts(to_points)
        assert len(f) >= 4 and len(f) == len(t)
        if are_all_collinear(f) or are_all_collinear(t):
            raise RankDeficient()
        if len(f) == 4:
            a = np.zeros((8, 8))
            b = np.zeros(8)
            for i, (fp, tp) in enumerate(zip(f, t)):
                a[2 * i] = [fp[0], fp[1], 1, 0, 0, 0, -tp[0] * fp[0], -tp[0] * fp[1]]
                a[2 * i + 1] = [0, 0, 0, fp[0], fp[1], 1, -tp[1] * fp[0], -tp[1] * fp[1]]
                b[2 * i], b[2 * i + 1] = tp[0], tp[1]
            try:
                h = np.linalg.solve(a, b)
            except np.linalg.LinAlgError:
                raise RankDeficient()
            self.matrix = np.array([[h[0], h[1], h[2]], [h[3], h[4], h[5]], [h[6], h[7], 1.0]])
            return
        accum = np.zeros((9, 9))
        for fp, tp in zip(f, t):
            fr = np.array([fp[0], fp[1], 1.0])
            b = np.zeros((2, 9))
            b[0, 0:3] = fr * tp[1]
            b[1, 0:3] = fr
            b[0, 3:6] = fr * (-tp[0])
            b[1, 6:9] = fr * (-tp[0])
            accum = accum + b.T @ b
        u, s, _, conv = matrix.svd(accum, "full_u", False)
        if conv != 0:
            raise NotConverged()
        self.matrix = u[:, -1].reshape(3, 3).copy()
