"""Host-side transform fits (reference geometry/transforms.zig:294-520), CPU only: they use the library's host SVD."""
import numpy as np
import pytest

from zignal_b200.geometry import AffineTransform, ProjectiveTransform, RankDeficient, SimilarityTransform, pinv


def test_rank_deficient_inputs():  # :294-337
    with pytest.raises(RankDeficient):
        SimilarityTransform([(0, 0), (0, 0)], [(1, 1), (1, 1)])
    with pytest.raises(RankDeficient):
        AffineTransform([(0, 0), (1, 0), (2, 0)], [(0, 0), (1, 0), (2, 0)])
    with pytest.raises(RankDeficient):
        ProjectiveTransform([(0, 0), (1, 0), (2, 0), (3, 0)], [(0, 0), (1, 0), (2, 0), (3, 0)])


def test_affine3_and_extra_correspondences():  # :339-405
    f = [(0, 0), (0, 1), (1, 1)]
    t = [(0, 1), (1, 1), (1, 0)]
    tf = AffineTransform(f, t)
    assert np.allclose(tf.matrix, [[0, 1], [-1, 0]], atol=1e-9) and np.allclose(tf.bias, [0, 1], atol=1e-9)
    itf = AffineTransform(t, f)
    for a, b in zip(f, t):
        assert np.allclose(tf.project(a), b, atol=1e-9) and np.allclose(itf.project(b), a, atol=1e-9)
    tf4 = AffineTransform(f + [(1, 0)], t + [(0, 0)])
    assert np.allclose(tf4.matrix, [[0, 1], [-1, 0]], atol=1e-9) and np.allclose(tf4.bias, [0, 1], atol=1e-9)


def test_projection4_dlib_golden():  # :407-452
    f = [(199.67754364, 200.17905235), (167.90229797, 175.55920601), (270.33649445, 207.96521187), (267.53637314, 188.24442387)]
    t = [(440.68012238, 275.45248032), (429.62512970, 262.64307976), (484.23328400, 279.44332123), (488.08315277, 272.79547691)]
    tr = ProjectiveTransform(f, t)
    gold = np.array([[-5.9291612941280800e-03, 7.0341614664190845e-03, -8.9922894648198459e-01],
                     [-2.8361695646354147e-03, 2.9060176209597761e-03, -4.3735741833190661e-01],
                     [-1.0156215756801098e-05, 1.3270311721030187e-05, -2.1603199531972065e-03]])
    scaled = tr.matrix * (gold[2, 2] / tr.matrix[2, 2])
    assert np.allclose(scaled, gold, atol=1e-3)
    for a, b in zip(f, t):
        assert np.allclose(tr.project(a), b, rtol=1e-5)
    inv = tr.inv()
    for a in f:
        assert np.allclose(inv.project(tr.project(a)), a, rtol=1e-5)


def test_projection8_dlib_golden():  # :454-495
    f = [(319.48406982, 240.21486282), (268.64367676, 210.67104721), (432.53839111, 249.55825424), (428.05819702, 225.89330864),
         (687.00787354, 240.97020721), (738.32287598, 208.32876205), (574.62890625, 250.60971451), (579.63378906, 225.37580109)]
    t = [(330.48120117, 408.22596359), (317.55538940, 393.26282501), (356.74267578, 411.06428146), (349.94784546, 400.26379395),
         (438.15582275, 411.75442886), (452.01367188, 398.08815765), (398.66107178, 413.83139420), (395.29974365, 401.73685455)]
    tr = ProjectiveTransform(f, t)
    gold = np.array([[7.9497770144471079e-05, 8.6315632819330035e-04, -6.3240797603906806e-01],
                     [3.9739851020393160e-04, 6.4356336568222570e-04, -7.7463154396817901e-01],
                     [1.0207719920241196e-06, 2.6961794891002063e-06, -2.1907681782918601e-03]])
    tol = np.sqrt(np.finfo(float).eps)
    m = tr.matrix if np.sign(tr.matrix[2, 2]) == np.sign(gold[2, 2]) else -tr.matrix
    assert np.allclose(m, gold, atol=tol)
    for a, b in zip(f, t):
        assert np.allclose(tr.project(a), b, rtol=1e-2)


def test_projective_exact_four_points():  # :497-520
    src = [(0, 0), (100, 0), (0, 100), (100, 100)]
    dst = [(50, 20), (150, 40), (30, 120), (130, 140)]
    fwd = ProjectiveTransform(src, dst)
    for s, d in zip(src, dst):
        assert np.allclose(fwd.project(s), d, atol=1e-9)
    back = ProjectiveTransform(dst, src)
    assert np.allclose(back.project(fwd.project((33, 71))), (33, 71), atol=1e-9)
    with pytest.raises(RankDeficient):
        ProjectiveTransform([(0, 0), (1, 1), (2, 2), (3, 3)], dst)


def test_similarity_recovers_rotation_scale_translation():
    rng = np.random.default_rng(0)
    pts = rng.standard_normal((6, 2))
    th, sc, tr = 0.7, 1.8, np.array([3.0, -2.0])
    rot = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    to = (sc * (rot @ pts.T)).T + tr
    s = SimilarityTransform(pts, to)
    assert np.allclose(s.matrix, sc * rot, atol=1e-9) and np.allclose(s.bias, tr, atol=1e-9)
    kind, m = s.as_f32()
    assert kind == 0 and m.dtype == np.float32 and m.size == 6


def test_pinv_matches_numpy():
    rng = np.random.default_rng(1)
    for shape in [(5, 3), (3, 5), (4, 4)]:
        a = rng.standard_normal(shape)
        p, rank = pinv(a)
        assert rank == min(shape) and np.allclose(p, np.linalg.pinv(a), atol=1e-10)
