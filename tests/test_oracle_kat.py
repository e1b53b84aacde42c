"""Known-answer tests that pin the CPU oracle to the facts the reference's own tests hold for the hot
path.  Each test cites the reference test it ports (paths under /root/reference/src/).
CPU only; the oracle is test infrastructure (oracle/zignal_oracle.h)."""
import numpy as np
import pytest

import oracle_lib as zo


# --------------------------------------------------------------------------------------- border.zig:65-139
def test_resolve_index_basic():
    for b in ("zero", "replicate", "mirror", "wrap"):
        assert zo.resolve_index(5, 10, b) == 5
        assert zo.resolve_index(0, 0, b) is None


def test_resolve_index_zero():
    for i in (-1, -5, 10, 15):
        assert zo.resolve_index(i, 10, "zero") is None


def test_resolve_index_replicate():
    assert zo.resolve_index(-1, 10, "replicate") == 0
    assert zo.resolve_index(-5, 10, "replicate") == 0
    assert zo.resolve_index(10, 10, "replicate") == 9
    assert zo.resolve_index(15, 10, "replicate") == 9


def test_resolve_index_mirror():
    assert zo.resolve_index(-1, 5, "mirror") == 1
    assert zo.resolve_index(-2, 5, "mirror") == 2
    assert zo.resolve_index(5, 5, "mirror") == 3
    assert zo.resolve_index(6, 5, "mirror") == 2
    assert zo.resolve_index(-1, 1, "mirror") == 0
    assert zo.resolve_index(5, 1, "mirror") == 0


def test_resolve_index_wrap():
    assert zo.resolve_index(-1, 5, "wrap") == 4
    assert zo.resolve_index(-6, 5, "wrap") == 4
    assert zo.resolve_index(5, 5, "wrap") == 0
    assert zo.resolve_index(6, 5, "wrap") == 1


# --------------------------------------------------------------------------------------- meta.zig:229-249
def test_meta_clamp_rounding():
    L = zo.lib()
    assert L.zo_clamp_u8_f32(-5.0) == 0
    assert L.zo_clamp_u8_f32(300.0) == 255
    assert L.zo_clamp_u8_f32(0.5) == 1      # half away from zero
    assert L.zo_clamp_u8_f32(1.5) == 2
    assert L.zo_clamp_u8_f32(2.5) == 3
    assert L.zo_clamp_u8_f32(254.49) == 254
    # divClampU8 (convolution.zig:18-22): symmetric rounding
    assert L.zo_div_clamp_u8(128, 256) == 1
    assert L.zo_div_clamp_u8(127, 256) == 0
    assert L.zo_div_clamp_u8(-128, 256) == 0
    assert L.zo_div_clamp_u8(255 * 256, 256) == 255
    assert L.zo_div_clamp_u8(10 ** 9, 256) == 255


# --------------------------------------------------------------------------------------- tests/integral.zig
def test_integral_all_ones_scalar():  # :11-28
    img = np.ones((21, 13), np.uint8)
    sat = zo.integral_plane(img)
    r, c = np.mgrid[0:21, 0:13]
    assert np.array_equal(sat, ((r + 1) * (c + 1)).astype(np.float32))


def test_integral_all_ones_view():  # :30-47
    img = np.ones((21, 13), np.uint8)
    view = img[3:10, 2:8]
    sat = zo.integral_plane(view)
    r, c = np.mgrid[0:7, 0:6]
    assert np.array_equal(sat, ((r + 1) * (c + 1)).astype(np.float32))


# --------------------------------------------------------------------------------------- tests/filters.zig
def test_box_blur_uniform():  # :87-103
    img = np.full((5, 5), 128, np.uint8)
    assert np.all(zo.box_blur(img, 1) == 128)


def test_box_blur_radius0_views():  # :50-75, :105-126
    base = np.arange(36, dtype=np.uint8).reshape(6, 6)
    view = base[1:5, 1:5]
    out = zo.box_blur(view, 0, out=np.zeros((4, 4), np.uint8))
    assert np.array_equal(out, view)


def test_box_blur_border_effects():  # :128-154
    img = np.zeros((5, 5), np.uint8)
    img[2, 2] = 255
    out = zo.box_blur(img, 1)
    assert out[0, 0] < out[2, 2] < 255
    assert out[2, 2] == 28  # 255/9 = 28.33 -> 28


def test_box_blur_border_area():  # :186-207: every border pixel of a uniform image keeps its value
    img = np.full((12, 12), 200, np.uint8)
    assert np.all(zo.box_blur(img, 3) == 200)


@pytest.mark.parametrize("size", [8, 32])
@pytest.mark.parametrize("radius", [1, 3])
def test_box_blur_rgba_alpha_preserved(size, radius):  # :234-264
    rng = np.random.default_rng(size * 10 + radius)
    img = rng.integers(0, 256, (size, size, 4), dtype=np.uint8)
    img[..., 3] = 255
    out = zo.box_blur(img, radius)
    assert np.all(out[..., 3] == 255)


def test_sharpen_uniform():  # :327-343
    img = np.full((5, 5), 100, np.uint8)
    assert np.all(zo.sharpen(img, 1) == 100)


def test_convolve_identity():  # :370-398, :662-699
    rng = np.random.default_rng(1)
    ident = np.zeros((3, 3), np.float32)
    ident[1, 1] = 1
    for shape in [(5, 5), (7, 9, 3), (6, 8, 4)]:
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        assert np.array_equal(zo.convolve(img, ident, "zero"), img)
    imgf = rng.random((9, 11), dtype=np.float32)
    assert np.array_equal(zo.convolve(imgf, ident, "mirror"), imgf)


def test_convolve_identity_into_view_sentinels():  # :701-744
    rng = np.random.default_rng(2)
    src = rng.integers(0, 256, (4, 4), dtype=np.uint8)
    big = np.full((8, 8), 0xAA, np.uint8)
    ident = np.zeros((3, 3), np.float32)
    ident[1, 1] = 1
    zo.convolve(src, ident, "zero", out=big[2:6, 2:6])
    assert np.array_equal(big[2:6, 2:6], src)
    mask = np.ones((8, 8), bool)
    mask[2:6, 2:6] = False
    assert np.all(big[mask] == 0xAA)


def test_convolve_zero_corner_white_rgb():  # :571-600
    img = np.full((5, 5, 3), 255, np.uint8)
    k = np.full((3, 3), 1.0 / 9.0, np.float32)
    out = zo.convolve(img, k, "zero")
    assert out[0, 0, 0] != 255
    assert abs(int(out[0, 0, 0]) - 113) <= 1
    # Q8 fixed point: round(256/9) = 28; 4 taps * 255 * 28 = 28560 -> /256 = 111.56 -> 112
    assert out[0, 0, 0] == 112
    assert out[2, 2, 0] == zo.lib().zo_div_clamp_u8(9 * 28 * 255, 256)


def test_convolve_issue_255():  # :1302-1342
    img = np.ones((10, 20), np.uint8)
    k = np.array([[1, 1, 1], [1, 0, 1], [1, 1, 1]], np.float32)
    out = zo.convolve(img, k, "zero", out=np.full((10, 20), 0xAA, np.uint8))
    assert np.all(out[1:9, 0] == 5)
    assert np.all(out[1:9, 1] == 8)


def test_convolve_replicate_corner():  # :427-463
    img = np.zeros((3, 3), np.uint8)
    img[1, 1] = 255
    k = np.array([[0.25, 0.25, 0], [0.25, 0.25, 0], [0, 0, 0]], np.float32)
    assert zo.convolve(img, k, "replicate")[0, 0] == 0


def test_separable_identity_view_f32():  # :602-632
    base = (np.arange(5)[:, None] * 10 + np.arange(5)[None, :]).astype(np.float32)
    view = base[1:4, 1:4]
    out = zo.conv_separable(view, [1.0], [1.0], "zero", out=np.zeros((3, 3), np.float32))
    assert np.array_equal(out, view)


def test_separable_identity_view_u8_sentinels():  # :746-783
    rng = np.random.default_rng(3)
    src = rng.integers(0, 256, (4, 4), dtype=np.uint8)
    big = np.full((8, 8), 0x55, np.uint8)
    zo.conv_separable(src, [1.0], [1.0], "zero", out=big[2:6, 2:6])
    assert np.array_equal(big[2:6, 2:6], src)
    assert np.count_nonzero(big != 0x55) <= 16


def test_separable_impulse():  # :469-491
    img = np.zeros((7, 7), np.float32)
    img[3, 3] = 1
    g = [0.25, 0.5, 0.25]
    out = zo.conv_separable(img, g, g, "zero")
    assert out[3, 3] < 1.0 and out[3, 2] > 0 and out[3, 3] > out[3, 2]
    assert out[3, 3] == np.float32(0.25) and out[3, 2] == np.float32(0.125)


def test_gaussian_sigma0_copy_and_negative():  # :1159-1180, image.zig:966-970
    rng = np.random.default_rng(4)
    img = rng.integers(0, 256, (6, 7), dtype=np.uint8)
    assert np.array_equal(zo.gaussian_blur(img, 0.0), img)
    with pytest.raises(ValueError):
        zo.gaussian_blur(img, -1.0)


def test_gaussian_taps_shape():  # image.zig:973-990
    t = zo.gaussian_taps(2.25)
    assert t.size == 15
    assert abs(float(t.sum()) - 1.0) < 1e-6
    assert np.array_equal(t, t[::-1])
    assert zo.gaussian_taps(1.0).size == 7
    assert zo.gaussian_taps(0.5).size == 5


def test_gaussian_ordering_property():  # :521-546
    img = np.zeros((21, 21), np.uint8)
    img[10, 10] = 255
    a = zo.gaussian_blur(img, 0.5)
    b = zo.gaussian_blur(img, 2.0)
    assert a[10, 10] > b[10, 10]
    assert b[10, 14] >= a[10, 14]


def test_gaussian_rgb_red_square_stays_red():  # :785-815
    img = np.zeros((20, 20, 3), np.uint8)
    img[5:15, 5:15, 0] = 255
    out = zo.gaussian_blur(img, 1.0)
    assert out[10, 10, 0] > 200 and out[10, 10, 1] == 0 and out[10, 10, 2] == 0


def test_uniform_channel_shortcut_equals_full_convolution():  # convolution.zig:363-429 is a pure optimisation
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (12, 17, 4), dtype=np.uint8)
    img[..., 3] = 200  # uniform alpha -> shortcut path
    taps = zo.gaussian_taps(1.0)
    out = zo.conv_separable(img, taps, taps, "mirror")
    alpha_plane = zo.conv_separable(np.ascontiguousarray(img[..., 3]), taps, taps, "mirror")
    assert np.array_equal(out[..., 3], alpha_plane)


# --------------------------------------------------------------------------------------- tests/interpolation.zig
def _gradient(rows, cols):
    r, c = np.mgrid[0:rows, 0:cols]
    return np.minimum(255, (r + c) * 255 // (rows + cols - 2)).astype(np.uint8)


def _checker(rows, cols):
    r, c = np.mgrid[0:rows, 0:cols]
    return np.where((r + c) % 2 == 0, 0, 255).astype(np.uint8)


def test_nearest_exact_and_rounding():  # :36-70
    img = _gradient(10, 10)
    for p in (0, 5, 9):
        assert zo.interpolate(img, p, p, "nearest") == img[p, p]
    ch = _checker(10, 10)
    assert zo.interpolate(ch, 0.4, 0.4, "nearest") == 0
    assert zo.interpolate(ch, 0.6, 0.6, "nearest") == 0
    assert zo.interpolate(ch, 1.5, 0.5, "nearest") == 255


def test_bilinear_midpoints():  # :72-108
    img = _gradient(10, 10)
    assert zo.interpolate(img, 0, 0, "bilinear") == img[0, 0]
    assert zo.interpolate(img, 5, 5, "bilinear") == img[5, 5]
    p = np.tile(np.array([0, 100, 200], np.uint8), (3, 1))
    assert zo.interpolate(p, 0.5, 0, "bilinear") == 50
    assert zo.interpolate(p, 0.5, 0.5, "bilinear") == 50


@pytest.mark.parametrize("method,pts,tol", [("bicubic", (2, 5), 0), ("catmull_rom", (2, 5), 0),
                                            ("lanczos", (3, 5), 1), ("mitchell", (2, 5), 1)])
def test_kernel_exact_pixels(method, pts, tol):  # :110-162
    img = _gradient(10, 10)
    for p in pts:
        v = zo.interpolate(img, p, p, method)
        assert abs(int(v) - int(img[p, p])) <= tol


def test_rgb_interpolation():  # :234-268
    r, c = np.mgrid[0:4, 0:4]
    img = np.stack([r * 85, c * 85, np.full_like(r, 128)], axis=-1).astype(np.uint8)
    assert np.array_equal(zo.interpolate(img, 1.6, 1.4, "nearest"), img[1, 2])
    v = zo.interpolate(img, 0.5, 0.5, "bilinear")
    assert tuple(v) == (43, 43, 128)
    assert zo.interpolate(img, 1.5, 1.5, "mitchell") is not None


def test_float_bilinear():  # :333-353
    r, c = np.mgrid[0:4, 0:4]
    img = (r * 0.25 + c * 0.25).astype(np.float32)
    assert abs(float(zo.interpolate(img, 1.5, 1.5, "bilinear")) - 0.75) < 1e-3
    assert zo.interpolate(img, 1.5, 1.5, "bicubic") is not None


def test_resize_bilinear_range():  # :270-305
    img = np.array([[0, 80, 160, 240]] * 4, np.uint8)
    out = zo.resize(img, (8, 8), "bilinear")
    assert out.max() <= 240


def test_catmull_rom_within_range():  # :307-331
    r, c = np.mgrid[0:5, 0:5]
    img = (50 + (r + c) * 20).astype(np.uint8)
    rr = np.float32(1.5)
    while rr < 3.5:
        cc = np.float32(1.5)
        while cc < 3.5:
            v = zo.interpolate(img, float(cc), float(rr), "catmull_rom")
            assert v is not None and 50 <= v <= 200
            cc = np.float32(cc + np.float32(0.1))
        rr = np.float32(rr + np.float32(0.1))


def test_single_pixel_image():  # :587-601
    img = np.array([[77]], np.uint8)
    for m in zo.INTERP:
        assert zo.interpolate(img, 0.0, 0.0, m) == 77


def test_interpolation_zig_remaining_cases():
    """image/tests/interpolation.zig: mitchell custom parameters (:164-180), boundary conditions under .mirror (:182-232), exact linear bilinear
    (:422-449), nearest discontinuity (:451-470), symmetry (:472-493), mitchell parameter effects (:495-529), lanczos weight normalisation
    (:531-548), extreme values (:550-585) and the clamping stress tests (:355-420, :603-630: every sample exists and stays a valid u8)."""
    img = _gradient(10, 10)
    for b, c in ((1.0, 0.0), (0.0, 0.5), (0.0, 0.75)):
        assert zo.interpolate(img, 5.5, 5.5, "mitchell", b=b, c=c) is not None
    for x, y in ((-0.4, 0), (9.4, 9.4), (-1, 0), (0, -1), (10, 0), (0, 10)):
        assert zo.interpolate(img, x, y, "nearest", "mirror") is not None
    for x, y in ((0, 0), (8.9, 8.9), (9.1, 9.1), (-0.1, 0)):
        assert zo.interpolate(img, x, y, "bilinear", "mirror") is not None
    for x, y in ((1, 1), (7.9, 7.9), (0.5, 0.5), (8.1, 8.1)):
        assert zo.interpolate(img, x, y, "bicubic", "mirror") is not None
    for x, y in ((2, 2), (6.9, 6.9), (1.5, 1.5), (7.1, 7.1)):
        assert zo.interpolate(img, x, y, "lanczos", "mirror") is not None
    q = np.array([[0, 100], [50, 150]], np.uint8)
    assert zo.interpolate(q, 0.5, 0, "bilinear") == 50 and zo.interpolate(q, 0, 0.5, "bilinear") == 25
    assert zo.interpolate(q, 0.5, 0.5, "bilinear") == 75 and zo.interpolate(q, 0.25, 0, "bilinear") == 25
    d = np.array([[0, 255], [100, 200]], np.uint8)
    assert zo.interpolate(d, 0.49, 0, "nearest") == 0 and zo.interpolate(d, 0.51, 0, "nearest") == 255
    r, c = np.mgrid[0:5, 0:5]
    sym = np.minimum(255, (np.abs(r - 2) + np.abs(c - 2)) * 50).astype(np.uint8)
    assert zo.interpolate(sym, 1.5, 2, "bilinear") == zo.interpolate(sym, 2.5, 2, "bilinear")
    assert zo.interpolate(sym, 2, 1.5, "bilinear") == zo.interpolate(sym, 2, 2.5, "bilinear")
    band = np.full((6, 6), 50, np.uint8)
    band[2:4] = 200
    vals = [int(zo.interpolate(band, 2.5, 1.8, "mitchell", b=b, c=c)) for b, c in ((1 / 3, 1 / 3), (1.0, 0.0), (0.0, 0.75))]
    assert len(set(vals)) > 1
    assert zo.interpolate(np.full((8, 8), 128, np.uint8), 4.3, 4.7, "lanczos") == 128
    ch = _checker(4, 4)
    assert abs(int(zo.interpolate(ch, 2, 2, "lanczos"))) <= 1
    assert zo.interpolate(ch, 0, 0, "nearest") == 0 and zo.interpolate(ch, 0, 0, "bilinear") == 0
    edge = np.tile(np.where(np.arange(6) < 3, 0, 255).astype(np.uint8), (6, 1))
    assert zo.interpolate(edge, 3.1, 2.5, "bicubic") is not None
    ch8 = _checker(8, 8)
    for method, kw in (("bicubic", {}), ("catmull_rom", {}), ("lanczos", {}), ("mitchell", {}), ("mitchell", {"b": 0.0, "c": 0.75})):
        y = np.float32(2.0)
        while y < 6.0:
            x = np.float32(2.0)
            while x < 6.0:
                assert zo.interpolate(ch8, float(x), float(y), method, **kw) is not None
                x = np.float32(x + np.float32(0.3))
            y = np.float32(y + np.float32(0.3))
    rgb = np.repeat(_checker(4, 4)[..., None], 3, axis=2)
    v = zo.interpolate(rgb, 1.3, 1.7, "bicubic")
    assert v is not None and v[0] == v[1] == v[2]                      # identical channels stay identical through the clamp


def test_interpolate_rejects_non_finite():  # interpolation.zig:73-75
    img = _gradient(4, 4)
    assert zo.interpolate(img, float("nan"), 0.0, "bilinear") is None
    assert zo.interpolate(img, float("inf"), 0.0, "nearest") is None


# --------------------------------------------------------------------------------------- tests/resize.zig:258-298 (shapes are host logic; see test_host_api)
def test_resize_same_shape_is_copy():
    img = _gradient(6, 7)
    assert np.array_equal(zo.resize(img, (6, 7), "lanczos"), img)


def test_resize_4to1_bicubic_constant_weights():  # at 4:1, fx = fy = 128 -> [-32,160,160,-32]
    rng = np.random.default_rng(6)
    img = rng.integers(0, 256, (16, 16, 3), dtype=np.uint8)
    out = zo.resize(img, (4, 4), "bicubic")
    w1 = np.array([-32, 160, 160, -32], np.int64)
    w = (np.outer(w1, w1) / 256).astype(np.int64)  # exact: all products divisible by 256
    assert w.sum() == 256
    blk = img[4:8, 4:8, 1].astype(np.int64)
    s = int((blk * w).sum())
    expect = int(np.clip(int(s / 256) if s >= 0 else -int(-s / 256), 0, 255))
    assert out[1, 1, 1] == expect


# --------------------------------------------------------------------------------------- tests/transforms.zig
def test_rotate_orthogonal_shapes_and_content():  # :160-209
    img = np.arange(1, 13, dtype=np.uint8).reshape(3, 4)
    r0 = zo.rotate(img, 0.0, "bilinear", "mirror")
    assert r0.shape == (3, 4) and np.array_equal(r0, img)
    r90 = zo.rotate(img, np.pi / 2, "bilinear", "mirror")
    assert r90.shape == (4, 3) and r90[3, 0] == 1 and np.array_equal(r90, np.rot90(img, 1))
    r180 = zo.rotate(img, np.pi, "bilinear", "mirror")
    assert r180.shape == (3, 4) and np.array_equal(r180, np.rot90(img, 2))
    r270 = zo.rotate(img, 3 * np.pi / 2, "bilinear", "mirror")
    assert r270.shape == (4, 3) and np.array_equal(r270, np.rot90(img, 3))


def test_rotate_45_grows():  # :211-229
    img = _checker(10, 10)
    out = zo.rotate(img, np.pi / 4, "bilinear", "mirror")
    assert out.shape[0] > 10 and out.shape[1] > 10
    assert zo.rotate_bounds(1080, 1920, np.float32(np.pi / 4)) == (2122, 2122)


def test_warp_identity_and_translation():
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, (9, 11, 4), dtype=np.uint8)
    out = zo.warp(img, np.zeros_like(img), "affine", [1, 0, 0, 1, 0, 0], "bilinear")
    assert np.array_equal(out, img)
    out = zo.warp(img, np.zeros_like(img), "affine", [1, 0, 0, 1, 2, 1], "nearest")
    assert np.array_equal(out[:8, :9], img[1:9, 2:11])
    outp = zo.warp(img, np.zeros_like(img), "projective", [2, 0, 0, 0, 2, 0, 0, 0, 2], "bilinear")
    assert np.array_equal(outp, img)


# --------------------------------------------------------------------------------------- matrix/svd.zig:498-636
def test_svd_wikipedia():
    a = np.array([[1, 0, 0, 0], [0, 0, 0, 2], [0, 3, 0, 0], [0, 0, 0, 0], [2, 0, 0, 0]], np.float64)
    u, s, v, rc = zo.svd(a, "full_u", True)
    assert rc == 0 and u.shape == (5, 5) and v.shape == (4, 4)
    assert np.all(s >= 0) and np.all(np.diff(s) <= 0)
    assert np.allclose(s, [3, np.sqrt(5), 2, 0], atol=1e-12)
    assert np.allclose(u[:, :4] @ np.diag(s) @ v.T, a, atol=1e-12)
    assert np.allclose(u.T @ u, np.eye(5), atol=1e-12)


def test_svd_modes_agree():  # :546-589
    rng = np.random.default_rng(8)
    a = rng.standard_normal((6, 4))
    _, s_full, _, _ = zo.svd(a, "full_u", True)
    u_sk, s_sk, v_sk, _ = zo.svd(a, "skinny_u", True)
    _, s_no, _, _ = zo.svd(a, "no_u", False)
    tol = np.sqrt(np.finfo(np.float64).eps)
    assert np.allclose(s_full, s_sk, rtol=tol) and np.allclose(s_full, s_no, rtol=tol)
    assert np.allclose(u_sk @ np.diag(s_sk) @ v_sk.T, a, atol=1e-12)
    assert np.allclose(s_sk, np.linalg.svd(a, compute_uv=False), rtol=1e-12)


def test_svd_identity_and_rank1():  # :591-636
    _, s, _, rc = zo.svd(np.eye(3), "full_u", True)
    assert rc == 0 and np.allclose(s, 1.0, rtol=np.sqrt(np.finfo(float).eps))
    a = np.array([[1, 2, 3], [2, 4, 6], [1, 2, 3]], np.float64)
    _, s, _, _ = zo.svd(a, "full_u", True)
    assert np.count_nonzero(s < np.sqrt(np.finfo(float).eps)) == 2


def test_svd_f32():
    rng = np.random.default_rng(9)
    a = rng.standard_normal((8, 5)).astype(np.float32)
    u, s, v, rc = zo.svd(a, "skinny_u", True)
    assert rc == 0
    assert np.allclose(u @ np.diag(s) @ v.T, a, atol=2e-5)


# --------------------------------------------------------------------------------------- matrix/test_ops_gemm.zig:246-299
def test_gemm_9x9_known_values():
    a = np.repeat(np.arange(1, 10, dtype=np.float32)[:, None], 9, axis=1)
    r1 = zo.gemm(a, a)
    assert r1[0, 0] == 45 and r1[1, 0] == 90 and r1[8, 8] == 405
    r2 = zo.gemm(a, a, trans_a=True)
    assert r2[0, 0] == 285 and r2[8, 8] == 285
    r3 = zo.gemm(a, a, trans_b=True)
    assert r3[0, 0] == 9 and r3[1, 1] == 36 and r3[8, 8] == 729
    r4 = zo.gemm(a, a, trans_a=True, trans_b=True)
    assert r4[0, 0] == 45 and r4[0, 8] == 405 and r4[8, 0] == 45 and r4[8, 8] == 405


@pytest.mark.parametrize("ta", [False, True])
@pytest.mark.parametrize("tb", [False, True])
@pytest.mark.parametrize("big", [False, True])
def test_gemm_matches_numpy(ta, tb, big):
    rng = np.random.default_rng(10)
    m, k, n = (13, 11, 9) if big else (2, 3, 2)
    a = rng.standard_normal((k, m) if ta else (m, k))
    b = rng.standard_normal((n, k) if tb else (k, n))
    c = rng.standard_normal((m, n))
    out = zo.gemm(a, b, ta, tb, alpha=0.5, beta=2.0, c=c)
    ref = 0.5 * ((a.T if ta else a) @ (b.T if tb else b)) + 2.0 * c
    assert np.allclose(out, ref, rtol=1e-12, atol=1e-12)
    with pytest.raises(ValueError):
        zo.gemm(rng.standard_normal((3, 4)), rng.standard_normal((5, 2)))


# --------------------------------------------------------------------------------------- fdm.zig:325-604
def test_fdm_gray_mean_exact():  # :429-464
    src = np.arange(100, dtype=np.uint8).reshape(100, 1)
    tgt = (100 + np.arange(100)).astype(np.uint8).reshape(100, 1)
    out = zo.fdm_match(src, tgt)
    assert out.astype(np.float64).mean() == 149.5


def test_fdm_color_mean_and_variance():  # :325-427
    i = np.arange(2500)
    x, y = i % 50, i // 50
    src = np.stack([100 + x % 20, 150 + y % 15, 80 + (x + y) % 25], -1).astype(np.uint8).reshape(50, 50, 3)
    tgt = np.stack([50 + x % 30, 70 + y % 20, 90 + (x + y) % 35], -1).astype(np.uint8).reshape(50, 50, 3)
    out = zo.fdm_match(src, tgt).reshape(-1, 3).astype(np.float64)
    t = tgt.reshape(-1, 3).astype(np.float64)
    assert np.all(np.abs(out.mean(0) - t.mean(0)) <= 2.0)
    assert np.all(np.abs(out.var(0) - t.var(0)) <= 1.0)


def test_fdm_gray_target_on_color_source():  # :531-581
    rng = np.random.default_rng(11)
    src = rng.integers(0, 256, (20, 20, 3), dtype=np.uint8)
    g = rng.integers(60, 200, (20, 20), dtype=np.uint8)
    tgt = np.repeat(g[..., None], 3, axis=2)
    out = zo.fdm_match(src, tgt)
    assert np.array_equal(out[..., 0], out[..., 1]) and np.array_equal(out[..., 1], out[..., 2])
    assert abs(out[..., 0].astype(float).mean() - g.astype(float).mean()) <= 2.0


def test_fdm_stats_match_numpy():
    rng = np.random.default_rng(12)
    img = rng.integers(0, 256, (31, 17, 3), dtype=np.uint8)
    mean, cov, gray = zo.fdm_stats(img)
    x = img.reshape(-1, 3).astype(np.float64) / 255.0
    assert np.allclose(mean, x.mean(0), rtol=1e-12)
    assert np.allclose(cov, np.cov(x.T), rtol=1e-10)
    assert not gray


# --------------------------------------------------------------------------------------- pca.zig:431-671
def _reconstruct(mean, comps, coeffs):
    return mean + comps @ coeffs


def test_pca_2d_reconstruct():  # :441-475
    data = np.array([[1, 2], [3, 4], [5, 6], [7, 8]], np.float64)
    mean, comps, eig = zo.pca_fit(data)
    v = np.array([4.0, 5.0])
    coeffs = comps.T @ (v - mean)
    assert np.allclose(_reconstruct(mean, comps, coeffs), v, atol=1e-10)


def test_pca_gram_path():  # :519-557
    data = np.array([[1, 0, 0], [3, 0, 0]], np.float64)
    mean, comps, eig = zo.pca_fit(data, 1)
    assert abs(eig[0] - 2.0) < 1e-9
    assert abs(abs(comps[0, 0]) - 1.0) < 1e-9 and abs(comps[1, 0]) < 1e-12 and abs(comps[2, 0]) < 1e-12
    t = zo.pca_transform(data, mean, comps)
    assert abs(t[0, 0] - (comps[:, 0] @ (data[0] - mean))) < 1e-12


def test_pca_10d_reconstruct():  # :559-588
    i, j = np.mgrid[0:5, 0:10]
    data = (i + j).astype(np.float64)
    mean, comps, eig = zo.pca_fit(data)
    v = np.arange(10, dtype=np.float64)
    coeffs = comps.T @ (v - mean)
    assert np.allclose(_reconstruct(mean, comps, coeffs), v, atol=1e-10)


def test_filters_zig_remaining_cases():
    """The tests of image/tests/filters.zig that bound a property rather than a value, ported so that every reference test on this path has a
    counterpart: boxBlur basic / zero radius / struct / struct comprehensive (:84-270), sharpen basic / zero radius / struct (:272-362),
    convolve blur kernel / border modes / 3x3 edge kernel / colour channels (:391-466, :628-690), convolveSeparable Gaussian approximation
    (:468-490), gaussianBlur basic / sigma variations (:492-545), the uniform-channel .zero-border regression (:570-596)."""
    assert np.all(zo.box_blur(np.full((5, 5), 128, np.uint8), 1) == 128)
    seq = np.arange(9, dtype=np.uint8).reshape(3, 3)
    assert np.array_equal(zo.box_blur(seq, 0), seq) and np.array_equal(zo.sharpen(seq + 10, 0), seq + 10)
    rgba = np.array([[[255, 0, 0, 255], [0, 255, 0, 255], [0, 0, 255, 255]], [[255, 255, 0, 255], [255, 255, 255, 255], [255, 0, 255, 255]],
                     [[0, 255, 255, 255], [128, 128, 128, 255], [0, 0, 0, 255]]], np.uint8)
    centre = zo.box_blur(rgba, 1)[1, 1]
    assert centre[0] != 255 and centre[1] != 255 and centre[2] != 255
    for size in (8, 32):
        for radius in (1, 3):
            rr, cc = np.mgrid[0:size, 0:size]
            img = np.stack([(255 * cc) // size, np.full_like(cc, 128), (255 * rr) // size, np.full_like(cc, 255)], axis=-1).astype(np.uint8)
            b = zo.box_blur(img, radius)
            assert np.all(b[..., 3] == 255)
            col = b[1:size, size // 2, 0].astype(int)
            assert np.all(np.abs(np.diff(col)) <= 15)
    edge = np.tile(np.where(np.arange(5) < 2, 64, 192).astype(np.uint8), (5, 1))
    sh = zo.sharpen(edge, 1)
    assert sh[2, 0] <= 64 and sh[2, 4] >= 192
    spot = np.full((3, 3, 4), 64, np.uint8)
    spot[..., 3] = 255
    spot[1, 1, :3] = 192
    assert np.all(zo.sharpen(spot, 1)[1, 1, :3] >= 192)
    step = np.tile(np.where(np.arange(5) < 2, 0, 255).astype(np.uint8), (5, 1))
    box = np.full((3, 3), np.float32(1.0) / np.float32(9.0), np.float32)
    assert 0 < zo.convolve(step, box, "replicate")[2, 2] < 255
    dot = np.zeros((3, 3), np.uint8)
    dot[1, 1] = 255
    k = np.array([[0.25, 0.25, 0], [0.25, 0.25, 0], [0, 0, 0]], np.float32)
    assert zo.convolve(dot, k, "replicate")[0, 0] == 0
    for border in ("zero", "mirror"):
        zo.convolve(dot, k, border)
    rr, cc = np.mgrid[0:10, 0:10]
    pat = ((rr * 7 + cc * 13) % 256).astype(np.uint8)
    lap = np.array([[-1, -1, -1], [-1, 8, -1], [-1, -1, -1]], np.float32)
    res = zo.convolve(pat, lap, "zero")
    assert res.shape == pat.shape
    want = np.zeros((10, 10), np.int64)                        # independent restatement of the u8 path: SCALE 256, round half up, clamp
    padded = np.pad(pat.astype(np.int64), 1)
    for dy in range(3):
        for dx in range(3):
            want += int(lap[dy, dx] * 256) * padded[dy:dy + 10, dx:dx + 10]
    assert np.array_equal(res, np.clip((want + 128) >> 8, 0, 255).astype(np.uint8))
    rr, cc = np.mgrid[0:5, 0:5]
    rgb = np.stack([(rr * 20) % 256, (cc * 20) % 256, ((rr + cc) * 10) % 256], axis=-1).astype(np.uint8)
    ident = np.array([[0, 0, 0], [0, 1, 0], [0, 0, 0]], np.float32)
    assert np.array_equal(zo.convolve(rgb, ident, "zero")[1:-1, 1:-1], rgb[1:-1, 1:-1])
    imp = np.zeros((7, 7), np.float32)
    imp[3, 3] = 1.0
    g1 = np.array([0.25, 0.5, 0.25], np.float32)
    r = zo.conv_separable(imp, g1, g1, "zero")
    assert r[3, 3] < 1.0 and r[3, 2] > 0 and r[3, 3] > r[3, 2] and r[3, 3] == np.float32(0.25) and r[3, 2] == np.float32(0.125)
    sq = np.zeros((11, 11), np.uint8)
    sq[3:8, 3:8] = 255
    bl = zo.gaussian_blur(sq, 1.0)
    assert sq[2, 5] == 0 and bl[2, 5] > 0 and bl[5, 5] > 200
    pix = np.zeros((15, 15), np.float32)
    pix[7, 7] = 1.0
    small, large = zo.gaussian_blur(pix, 0.5), zo.gaussian_blur(pix, 2.0)
    assert small[7, 7] > large[7, 7] and large[7, 5] > small[7, 5]
    white = np.full((5, 5, 3), 255, np.uint8)
    corner = zo.convolve(white, box, "zero")[0, 0]
    assert corner[0] != 255 and abs(int(corner[0]) - 113) <= 1


def test_integral_and_transform_remaining_cases():
    """image/tests/integral.zig:49-150 (per-channel planes of an all-ones Rgba image, Rgb vs Rgba planes agree, box sums from the table) and
    image/tests/transforms.zig:211-229, 382-406 (a 45 degree rotation grows the canvas; insert copies under .none and composites under .normal)."""
    ones = np.ones((21, 13), np.uint8)
    area = (np.arange(1, 22)[:, None] * np.arange(1, 14)[None, :]).astype(np.float32)
    assert np.array_equal(zo.integral_plane(ones), area)                     # the struct test applies this plane per channel (:49-68)
    seed, vals = 0, []
    for _ in range(100):
        seed = (seed + 17) % 256
        vals.append((seed, (seed + 50) % 256, (seed + 100) % 256))
    rgb = np.array(vals, np.uint8).reshape(10, 10, 3)
    rgba = np.concatenate([rgb, np.full((10, 10, 1), 255, np.uint8)], axis=-1)
    for ch in range(3):
        assert np.array_equal(zo.integral_plane(np.ascontiguousarray(rgb[..., ch])), zo.integral_plane(np.ascontiguousarray(rgba[..., ch])))
    sat = zo.integral_plane(np.arange(1, 10, dtype=np.uint8).reshape(3, 3))

    def box_sum(r1, c1, r2, c2):                                              # integral.zig:85-90
        return sat[r2, c2] - (sat[r2, c1 - 1] if c1 > 0 else 0) - (sat[r1 - 1, c2] if r1 > 0 else 0) + (sat[r1 - 1, c1 - 1] if r1 > 0 and c1 > 0 else 0)
    assert box_sum(0, 0, 2, 2) == 45 and box_sum(0, 0, 1, 1) == 12 and box_sum(1, 1, 2, 2) == 28 and box_sum(1, 1, 1, 1) == 5
    chk = _checker(10, 10)
    rot = zo.rotate(255 - chk, float(np.float32(np.pi / 4)), "bilinear", "mirror")
    assert rot.shape[0] > 10 and rot.shape[1] > 10
    base, overlay = (0, 0, 255, 255), (255, 0, 0, 128)
    dest = np.array([[base]], np.uint8)
    src = np.array([[overlay]], np.uint8)
    assert tuple(zo.insert(dest, src, (0.0, 0.0, 1.0, 1.0), 0.0, "nearest", blend="none")[0, 0]) == overlay
    assert tuple(zo.insert(dest, src, (0.0, 0.0, 1.0, 1.0), 0.0, "nearest", blend="normal")[0, 0]) == zo.blend_rgba8(base, overlay, "normal")


def test_extract_kats():
    """image/tests/transforms.zig:228-315: extract of a 3x3 region at 0 and 90 degrees, and the single-pixel-axis centring rule."""
    img = (np.arange(5)[:, None] * 10 + np.arange(5)[None, :]).astype(np.uint8)
    rect = (1.0, 1.0, 3.0, 3.0)
    out0 = zo.extract(img, np.zeros((3, 3), np.uint8), rect, 0.0, "nearest", "mirror")
    assert out0.tolist() == [[11, 12, 13], [21, 22, 23], [31, 32, 33]]
    out90 = zo.extract(img, np.zeros((3, 3), np.uint8), rect, np.float32(np.pi / 2.0), "nearest", "mirror")
    assert out90.tolist() == [[13, 23, 33], [12, 22, 32], [11, 21, 31]]
    assert zo.extract(img, np.zeros((1, 1), np.uint8), rect, 0.0, "nearest", "mirror").tolist() == [[22]]
    assert zo.extract(img, np.zeros((1, 3), np.uint8), rect, 0.0, "nearest", "mirror").tolist() == [[21, 22, 23]]
    assert zo.extract(img, np.zeros((3, 1), np.uint8), rect, 0.0, "nearest", "mirror").tolist() == [[12], [22], [32]]
    # "extract from empty image regression" (:411-424): replicate / wrap on a 0x0 source give zeros, no panic
    empty = np.zeros((0, 0), np.uint8)
    for border in ("replicate", "wrap"):
        assert not zo.extract(empty, np.full((2, 2), 9, np.uint8), (0.0, 0.0, 2.0, 2.0), 0.0, "nearest", border).any()


def test_sobel_kat():
    """image/tests/filters.zig:548-569: a vertical step edge gives > 200 on the edge column and < 50 away from it."""
    img = np.where(np.arange(5)[None, :] < 2, 0, 255).astype(np.uint8).repeat(5, axis=0).reshape(5, 5)
    img = np.tile(np.where(np.arange(5) < 2, 0, 255).astype(np.uint8), (5, 1))
    e = zo.sobel(img)
    assert e[2, 2] > 200 and e[2, 0] < 50
    assert e[2, 2] == 255 and e[2, 1] == 255 and e[2, 0] == 0 and e[2, 4] == 0     # |gx| = 4 * 255 -> 255 after /4


def test_canny_kat():
    """image/tests/filters.zig:1182-1300: a vertical step is found near its column (u8 and Rgb), sigma = 0 is valid, parameter
    validation (InvalidSigma / InvalidThreshold / InvalidParameter).  Plus an independent breadth-first restatement of the
    hysteresis on a hand-made chain (strong seed, weak tail, isolated weak pixel)."""
    step = np.tile(np.where(np.arange(10) < 5, 0, 255).astype(np.uint8), (10, 1))
    e = zo.canny(step, 1.0, 50, 100)
    assert e.shape == step.shape and set(np.unique(e)) <= {0, 255}
    assert e[:, 4:7].any()
    assert not e[0].any() and not e[-1].any() and not e[:, 0].any() and not e[:, -1].any()   # NMS never marks the border (edges.zig:713)
    rgb = np.zeros((8, 8, 3), np.uint8)
    rgb[:, :4, 0] = 255
    rgb[:, 4:, 1] = 255
    assert zo.canny(rgb, 1.0, 30, 90)[:, 3:6].any()
    ramp = (np.arange(5)[:, None] * 10 + np.arange(5)[None, :]).astype(np.uint8)
    zo.canny(ramp, 0.0, 50, 100)
    for args, status in [((-1.0, 50, 100), 2), ((1.0, -1, 100), 14), ((1.0, 50, -1), 14), ((1.0, 100, 50), 14), ((1.0, 50, 50), 14),
                         ((np.nan, 50, 100), 5), ((1.0, np.nan, 100), 5), ((1.0, 50, np.nan), 5), ((np.inf, 50, 100), 5),
                         ((1.0, np.inf, 100), 5), ((1.0, 50, np.inf), 5), ((-np.inf, 50, 100), 5)]:
        with pytest.raises(zo.OracleStatus) as ei:
            zo.canny(ramp, *args)
        assert ei.value.status == status, args
    assert not zo.canny(np.full((2, 9), 200, np.uint8), 0.0, 1, 2).any()       # < 3 rows: no interior, nothing survives
    # sigma = 0 on an integer image: every stage is exact in f32, so a float64 numpy restatement must agree bit for bit.
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, (23, 31), dtype=np.uint8)
    p = np.pad(img.astype(np.float64), 1, mode="edge")
    gx = (p[:-2, 2:] - p[:-2, :-2]) + 2 * (p[1:-1, 2:] - p[1:-1, :-2]) + (p[2:, 2:] - p[2:, :-2])
    gy = (p[2:, :-2] - p[:-2, :-2]) + 2 * (p[2:, 1:-1] - p[:-2, 1:-1]) + (p[2:, 2:] - p[:-2, 2:])
    mag = np.sqrt((gx * gx + gy * gy).astype(np.float32)).astype(np.float32)
    K = np.float32(0.414213562)
    ax, ay = np.abs(gx).astype(np.float32), np.abs(gy).astype(np.float32)
    nms = np.zeros(img.shape, bool)
    for r in range(1, img.shape[0] - 1):
        for c in range(1, img.shape[1] - 1):
            if ay[r, c] <= K * ax[r, c]:
                d = (0, -1, 0, 1)
            elif ax[r, c] <= K * ay[r, c]:
                d = (-1, 0, 1, 0)
            elif gx[r, c] * gy[r, c] > 0:
                d = (-1, 1, 1, -1)
            else:
                d = (-1, -1, 1, 1)
            nms[r, c] = mag[r, c] >= mag[r + d[0], c + d[1]] and mag[r, c] >= mag[r + d[2], c + d[3]]
    low, high = 300.0, 700.0
    out = nms & (mag >= high)
    weak = nms & (mag >= low)
    while True:                                  # fixed point of "weak next to an edge becomes an edge" == the BFS closure
        grown = np.pad(out, 1)
        nb = np.zeros_like(out)
        for dr in range(3):
            for dc in range(3):
                nb |= grown[dr:dr + out.shape[0], dc:dc + out.shape[1]]
        nxt = out | (weak & nb)
        if np.array_equal(nxt, out):
            break
        out = nxt
    got = zo.canny(img, 0.0, low, high)
    assert np.array_equal(got, np.where(out, 255, 0).astype(np.uint8))
    assert 0 < int(out.sum()) < out.size and int((weak & ~out).sum()) > 0       # the case exercises both hysteresis outcomes


def _order_brute(img, radius, border, fn):
    """Brute-force window gather (border-resolved indices, zero for out-of-range) -> fn(sorted window)."""
    planes = img[..., None] if img.ndim == 2 else img
    rows, cols, ch = planes.shape
    ri = [zo.resolve_index(i, rows, border) for i in range(-radius, rows + radius)]
    ci = [zo.resolve_index(i, cols, border) for i in range(-radius, cols + radius)]
    out = np.zeros_like(planes)
    for k in range(ch):
        padded = np.zeros((rows + 2 * radius, cols + 2 * radius), np.uint8)
        for y, r in enumerate(ri):
            for x, c in enumerate(ci):
                if r is not None and c is not None:
                    padded[y, x] = planes[r, c, k]
        for r in range(rows):
            for c in range(cols):
                out[r, c, k] = fn(np.sort(padded[r:r + 2 * radius + 1, c:c + 2 * radius + 1].ravel()).astype(np.int64))
    return out.reshape(img.shape)


def test_order_statistic_kats():
    """image/tests/filters.zig:817-966 (median removes an impulse, percentile 1.0 with .zero, Rgb median, min == percentile 0,
    max == percentile 1, midpoint, alpha-trimmed mean, InvalidTrim) and a brute-force sort-based restatement on random images."""
    imp = np.zeros((5, 5), np.uint8)
    imp[2, 2] = 255
    med = zo.order_blur(imp, 1, "percentile", 0.5, "mirror")
    assert med[2, 2] == 0 and med[2, 1] == 0 and med[1, 2] == 0
    seq = np.arange(9, dtype=np.uint8).reshape(3, 3)
    mx = zo.order_blur(seq, 1, "percentile", 1.0, "zero")
    assert mx[1, 1] == 8 and mx[0, 0] == 4
    rgb = np.tile(np.array([32, 64, 96], np.uint8), (3, 3, 1))
    rgb[1, 1] = (255, 0, 0)
    m = zo.order_blur(rgb, 1, "percentile", 0.5, "mirror")
    assert tuple(m[1, 1]) == (32, 64, 96) and tuple(m[0, 0]) == (32, 64, 96)
    assert zo.order_blur(seq, 1, "midpoint", 0.0, "replicate")[1, 1] == 4
    assert zo.order_blur(seq, 1, "alpha_trimmed", 0.12, "replicate")[1, 1] == 4
    for trim in (0.6, 0.5, -0.1, np.nan):
        with pytest.raises(zo.OracleStatus) as ei:
            zo.order_blur(seq, 1, "alpha_trimmed", trim, "replicate")
        assert ei.value.status == 16
    for pct in (-0.01, 1.01):
        with pytest.raises(zo.OracleStatus) as ei:
            zo.order_blur(seq, 1, "percentile", pct, "mirror")
        assert ei.value.status == 15
    assert np.array_equal(zo.order_blur(seq, 0, "percentile", 7.0, "mirror"), seq)      # radius 0 copies before the percentile check (:43-50)
    with pytest.raises(zo.OracleStatus) as ei:
        zo.order_blur(np.zeros((4, 4), np.float32), 1, "percentile", 0.5, "mirror")
    assert ei.value.status == 3
    rng = np.random.default_rng(11)
    for shape, radius in [((9, 13), 1), ((7, 6), 2), ((5, 4, 3), 3), ((6, 8, 4), 1), ((1, 5), 2), ((3, 1), 1)]:
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        area = (2 * radius + 1) ** 2
        for border in ("zero", "replicate", "mirror", "wrap"):
            for pct in (0.0, 0.1, 0.5, 0.73, 1.0):
                rank = min(int(np.floor(pct * (area - 1) + 1e-12)), area - 1)
                assert np.array_equal(zo.order_blur(img, radius, "percentile", pct, border),
                                      _order_brute(img, radius, border, lambda w: w[rank])), (shape, radius, border, pct)
            assert np.array_equal(zo.order_blur(img, radius, "midpoint", 0.0, border),
                                  _order_brute(img, radius, border, lambda w: (w[0] + w[-1] + 1) // 2)), (shape, radius, border)
            for trim in (0.0, 0.12, 0.3, 0.49):
                t = min(int(np.floor(trim * area)), area // 2)
                kept = area - 2 * t
                assert np.array_equal(zo.order_blur(img, radius, "alpha_trimmed", trim, border),
                                      _order_brute(img, radius, border, lambda w: min(255, (int(w[t:area - t].sum()) + kept // 2) // kept))), (shape, trim)


def test_metrics_kats():
    """image/tests/psnr.zig:13-130 and image/metrics.zig:253-293 (meanPixelError RGB example, ssim on a checkerboard), plus a float64
    numpy restatement of SSIM's 11x11 Gaussian-window statistics."""
    a = np.array([[100, 150], [200, 250]], np.uint8)
    b = np.array([[110, 140], [205, 245]], np.uint8)
    assert zo.psnr(a, a) == np.inf
    assert abs(zo.psnr(a, b) - 30.171) < 0.01
    assert zo.psnr(a, b) == 20.0 * np.log10(255.0) - 10.0 * np.log10(62.5)
    with pytest.raises(zo.OracleStatus) as ei:
        zo.psnr(a, np.zeros((2, 3), np.uint8))
    assert ei.value.status == 1
    rgb1 = np.tile(np.array([100, 150, 200], np.uint8), (2, 2, 1))
    rgb2 = np.tile(np.array([110, 140, 205], np.uint8), (2, 2, 1))
    assert abs(zo.psnr(rgb1, rgb2) - 29.38) < 0.01
    ra = np.array([[[255, 0, 0, 255], [0, 255, 0, 255]]], np.uint8)
    rb = np.array([[[250, 5, 0, 255], [0, 250, 5, 255]]], np.uint8)
    assert abs(zo.psnr(ra, rb) - 37.16) < 0.01
    f1 = np.array([[0.5, 0.7], [0.3, 0.9]], np.float32)
    f2 = np.array([[0.4, 0.8], [0.2, 1.0]], np.float32)
    assert abs(zo.psnr(f1, f2) - 20.0) < 0.01
    one = np.array([[[255, 0, 0]]], np.uint8)
    assert abs(zo.mean_pixel_error(one, np.zeros_like(one)) - 1.0 / 3.0) < 1e-9
    assert zo.mean_pixel_error(np.zeros((0, 0), np.uint8), np.zeros((0, 0), np.uint8)) == 0.0
    chk = np.zeros((12, 12, 3), np.uint8)
    rr, cc = np.mgrid[0:12, 0:12]
    chk[(rr + cc) % 2 == 0] = (255, 0, 0)
    chk[(rr + cc) % 2 == 1] = (0, 255, 0)
    assert zo.ssim(chk, np.zeros_like(chk)) < 0.99
    assert zo.ssim(chk, chk) == 1.0
    with pytest.raises(zo.OracleStatus) as ei:
        zo.ssim(np.zeros((10, 30), np.uint8), np.zeros((10, 30), np.uint8))
    assert ei.value.status == 17
    rng = np.random.default_rng(3)
    x = rng.integers(0, 256, (19, 23), dtype=np.uint8)
    y = np.clip(x.astype(int) + rng.integers(-30, 31, x.shape), 0, 255).astype(np.uint8)
    yy, xx = np.mgrid[-5:6, -5:6].astype(np.float64)
    w = np.exp(-(xx * xx + yy * yy) / (2.0 * 1.5 * 1.5))
    w /= w.sum()
    c1, c2 = (0.01 * 255.0) ** 2, (0.03 * 255.0) ** 2
    vals = []
    for r in range(5, x.shape[0] - 5):
        for c in range(5, x.shape[1] - 5):
            px, py = x[r - 5:r + 6, c - 5:c + 6].astype(np.float64), y[r - 5:r + 6, c - 5:c + 6].astype(np.float64)
            mx, my = (w * px).sum(), (w * py).sum()
            sx, sy, sxy = max(0.0, (w * px * px).sum() - mx * mx), max(0.0, (w * py * py).sum() - my * my), (w * px * py).sum() - mx * my
            vals.append((2 * mx * my + c1) * (2 * sxy + c2) / ((mx * mx + my * my + c1) * (sx + sy + c2)))
    assert abs(zo.ssim(x, y) - np.mean(vals)) < 1e-12
    xf = rng.random((14, 15, 4), dtype=np.float32)
    yf = np.clip(xf + rng.normal(0, 0.05, xf.shape).astype(np.float32), 0, 1).astype(np.float32)
    assert 0.0 < zo.ssim(xf, yf) < 1.0 and abs(zo.mean_pixel_error(xf, yf) - np.abs(xf.astype(np.float64) - yf).mean()) < 1e-12


def test_blend_kats():
    """blending.zig:194-421: the reference's own blendColors(u8, ...) expectations, and insert with a blend mode on an Rgba(u8) canvas."""
    B = zo.blend_rgba8
    r = B((100, 100, 100, 255), (200, 200, 200, 128), "normal")
    assert all(140 < v < 160 for v in r[:3])
    assert B((255, 255, 255, 255), (128, 128, 128, 255), "multiply")[:3] == (128, 128, 128)
    assert B((0, 0, 0, 255), (128, 128, 128, 255), "screen")[:3] == (128, 128, 128)
    assert B((100, 100, 100, 255), (200, 200, 200, 0), "normal") == (100, 100, 100, 255)
    r = B((100, 100, 100, 128), (200, 200, 200, 128), "normal")
    assert 190 <= r[3] <= 192 and 130 < r[0] < 170
    r = B((0, 0, 0, 0), (200, 150, 100, 180), "normal")
    assert r[3] == 180 and abs(r[0] - 200) <= 1 and abs(r[1] - 150) <= 1 and abs(r[2] - 100) <= 1
    m, sc = B((100, 100, 100, 200), (50, 50, 50, 100), "multiply"), B((100, 100, 100, 200), (50, 50, 50, 100), "screen")
    assert abs(m[3] - 221) <= 2 and abs(sc[3] - 221) <= 2 and m[0] < sc[0]
    for mode in ("multiply", "screen", "exclusion"):
        assert B((25, 75, 125, 0), (200, 150, 100, 180), mode) == (200, 150, 100, 180)
    assert B((100, 100, 100, 255), (200, 200, 200, 255), "none") == (200, 200, 200, 255)
    assert B((100, 200, 100, 255), (200, 100, 100, 255), "darken")[:3] == (100, 100, 100)
    assert B((100, 200, 100, 255), (200, 100, 100, 255), "lighten")[:3] == (200, 200, 100)
    assert B((200, 100, 50, 255), (50, 200, 200, 255), "difference")[:3] == (150, 100, 150)
    assert B((0, 128, 255, 255), (255, 255, 255, 255), "color_dodge")[:3] == (0, 255, 255)      # B = 0 -> 0; S = 1 -> 1 (W3C)
    assert B((255, 128, 0, 255), (0, 0, 0, 255), "color_burn")[:3] == (255, 0, 0)                # B = 1 -> 1; S = 0 -> 0
    # float identities from the f32 tests, on values exactly representable as n / 255: overlay / hard light swap roles.
    for b in (0, 51, 102, 153, 204, 255):
        for o in (0, 85, 170, 255):
            assert B((b, b, b, 255), (o, o, o, 255), "overlay") == B((o, o, o, 255), (b, b, b, 255), "hard_light")
            e = b / 255 + o / 255 - 2 * (b / 255) * (o / 255)
            assert abs(B((b, b, b, 255), (o, o, o, 255), "exclusion")[0] - 255 * e) <= 0.5 + 1e-3
    canvas = np.full((6, 6, 4), 255, np.uint8)
    canvas[..., :3] = 100
    patch = np.zeros((2, 2, 4), np.uint8)
    patch[...] = (200, 200, 200, 128)
    out = zo.insert(canvas, patch, (2.0, 2.0, 4.0, 4.0), 0.0, "nearest", blend="normal")
    assert tuple(out[2, 2]) == B((100, 100, 100, 255), (200, 200, 200, 128), "normal") and tuple(out[0, 0]) == (100, 100, 100, 255)
    gray = np.full((6, 6), 9, np.uint8)
    assert np.array_equal(zo.insert(gray, np.full((2, 2), 77, np.uint8), (2.0, 2.0, 4.0, 4.0), 0.0, "nearest", blend="multiply")[2:4, 2:4],
                          np.full((2, 2), 77, np.uint8))                                         # non-Rgba samples are assigned (image.zig:90-94)


def test_motion_blur_kats():
    """image/tests/filters.zig:969-1160: horizontal / vertical / diagonal streaks of a bright spot, zero distance, RGB, radial zoom keeps the
    centre, radial spin spreads a point, zero strength copies; plus a float64 numpy restatement of the diagonal line integral."""
    spot = np.zeros((5, 5), np.uint8)
    spot[2, 2] = 255
    h = zo.motion_blur_linear(spot, 0.0, 3)
    assert h[2, 1] > 0 and h[2, 2] > 0 and h[2, 3] > 0 and h[1, 2] == 0 and h[3, 2] == 0
    v = zo.motion_blur_linear(spot, float(np.float32(np.pi / 2)), 3)
    assert v[1, 2] > 0 and v[2, 2] > 0 and v[3, 2] > 0 and v[2, 1] == 0 and v[2, 3] == 0
    d = zo.motion_blur_linear(spot, float(np.float32(np.pi / 4)), 3)
    assert d[1, 1] > 0 and d[2, 2] > 0 and d[3, 3] > 0
    assert np.array_equal(zo.motion_blur_linear(spot, 0.0, 0), spot)
    yy, xx = np.mgrid[0:7, 0:7].astype(np.float32)
    dist = np.sqrt((xx - 3) ** 2 + (yy - 3) ** 2)
    ring = np.where((dist > 1.5) & (dist < 2.5), 255, 0).astype(np.uint8)
    assert abs(int(zo.motion_blur_radial(ring, 0.5, 0.5, 0.5)[3, 3]) - int(ring[3, 3])) < 20
    pt = np.zeros((7, 7), np.uint8)
    pt[2, 4] = 255
    sp = zo.motion_blur_radial(pt, 0.5, 0.5, 0.5, spin=True)
    assert sp[2, 4] > 0 and int((sp > 0).sum()) > 1
    seq = np.arange(9, dtype=np.uint8).reshape(3, 3)
    assert np.array_equal(zo.motion_blur_radial(seq, 0.5, 0.5, 0.0), seq)
    rng = np.random.default_rng(17)
    img = rng.random((9, 11)).astype(np.float32)
    angle, distance = np.float32(0.6), 4
    ca, sa = np.float64(np.cos(angle, dtype=np.float32)), np.float64(np.sin(angle, dtype=np.float32))
    want = np.zeros(img.shape)
    for r in range(img.shape[0]):
        for c in range(img.shape[1]):
            acc, n, t = 0.0, 0, -distance / 2.0
            while t <= distance / 2.0:
                sx, sy = np.float32(np.float32(c) + np.float32(np.float32(t) * np.float32(ca))), np.float32(np.float32(r) + np.float32(np.float32(t) * np.float32(sa)))
                if 0 <= sx < img.shape[1] and 0 <= sy < img.shape[0]:
                    x0, y0 = int(np.floor(sx)), int(np.floor(sy))
                    x1, y1 = min(x0 + 1, img.shape[1] - 1), min(y0 + 1, img.shape[0] - 1)
                    fx, fy = float(sx) - x0, float(sy) - y0
                    acc += (img[y0, x0] * (1 - fx) + img[y0, x1] * fx) * (1 - fy) + (img[y1, x0] * (1 - fx) + img[y1, x1] * fx) * fy
                    n += 1
                t += 1.0
            want[r, c] = acc / n if n else img[r, c]
    assert np.allclose(zo.motion_blur_linear(img, float(angle), distance), want, rtol=2e-6, atol=1e-7)


def test_convert_kats():
    """color.zig:1556-1582, 1854-1866 (scalar colours, grayscale conversion, clamping) through Image.convert's per-pixel convertColor."""
    U8, F32, RGB8, RGBA8, RGBAF32 = range(5)
    one = lambda v, dt: np.array([[v]], dt)
    assert tuple(zo.convert(one(128, np.uint8), RGB8)[0, 0]) == (128, 128, 128)            # convertColor(Rgb(u8), u8 128)
    assert tuple(zo.convert(one(0.5, np.float32), RGB8)[0, 0]) == (128, 128, 128)          # convertColor(Rgb(u8), 0.5)
    assert abs(float(zo.convert(one(128, np.uint8), F32)[0, 0]) - 128.0 / 255.0) < 1e-7
    assert zo.convert(one(0.5, np.float32), U8)[0, 0] == 128
    assert zo.convert(one(-0.5, np.float32), U8)[0, 0] == 0 and zo.convert(one(1.5, np.float32), U8)[0, 0] == 255
    assert zo.convert(np.array([[[128, 128, 128]]], np.uint8), U8)[0, 0] == 128            # Rgb.to(.gray)
    assert zo.convert(np.array([[[255, 0, 0]]], np.uint8), U8)[0, 0] == 54
    assert zo.convert(np.array([[[255, 0, 0, 128]]], np.uint8), U8)[0, 0] == 54            # Rgba ignores alpha
    over = np.array([[[1.2, -0.2, 0.5, 2.0]]], np.float32)                                 # Rgb(f32){1.2, -0.2, 0.5}.as(u8) = 255, 0, 128
    assert tuple(zo.convert(over, RGBA8)[0, 0]) == (255, 0, 128, 255) and tuple(zo.convert(over, RGB8)[0, 0]) == (255, 0, 128)
    assert tuple(zo.convert(np.array([[[10, 20, 30]]], np.uint8), RGBA8)[0, 0]) == (10, 20, 30, 255)
    assert tuple(zo.convert(np.array([[[10, 20, 30, 40]]], np.uint8), RGB8)[0, 0]) == (10, 20, 30)
    f = zo.convert(np.array([[[255, 0, 51]]], np.uint8), RGBAF32)[0, 0]
    assert f[0] == 1.0 and f[1] == 0.0 and f[2] == np.float32(51) / np.float32(255) and f[3] == 1.0
    g = zo.convert(one(0.25, np.float32), RGBAF32)[0, 0]
    assert tuple(g) == (0.25, 0.25, 0.25, 1.0)
    y = zo.convert(np.array([[[1.0, 0.0, 0.0, 0.3]]], np.float32), F32)[0, 0]
    assert abs(float(y) - 0.2126) < 1e-6
    rng = np.random.default_rng(4)
    img = rng.integers(0, 256, (5, 7, 4), dtype=np.uint8)
    assert np.array_equal(zo.convert(zo.convert(img, RGBAF32), RGBA8), img)                 # u8 -> f32 -> u8 is the identity
    assert np.array_equal(zo.convert(img, RGBA8), img)                                      # same type copies


def test_insert_extract_inverse_kat():
    """image/tests/transforms.zig:316-381: extract then insert reproduces the centre of the source (avg error < 25); and the
    documented properties of insert: pixels outside the rectangle stay untouched, an empty source is a no-op."""
    src = ((np.arange(64)[:, None] + np.arange(64)[None, :]) % 256).astype(np.uint8)
    for rect, angle, size, method in [((10.0, 10.0, 50.0, 50.0), 0.0, 40, "bilinear"), ((15.0, 15.0, 45.0, 45.0), np.pi / 4, 30, "bilinear"),
                                      ((20.0, 20.0, 40.0, 40.0), 0.0, 40, "bicubic")]:
        ext = zo.extract(src, np.zeros((size, size), np.uint8), rect, angle, method, "mirror")
        canvas = zo.insert(np.zeros((64, 64), np.uint8), ext, rect, angle, method)
        cx, cy = (rect[0] + rect[2]) * 0.5, (rect[1] + rect[3]) * 0.5
        cs = min(rect[2] - rect[0], rect[3] - rect[1]) * 0.6
        r0, r1, c0, c1 = int(cy - cs / 2), int(cy + cs / 2), int(cx - cs / 2), int(cx + cs / 2)
        assert np.abs(src[r0:r1, c0:c1].astype(int) - canvas[r0:r1, c0:c1].astype(int)).mean() < 25
    base = np.full((20, 20), 7, np.uint8)
    out = zo.insert(base, np.full((4, 4), 200, np.uint8), (5.0, 5.0, 9.0, 9.0), 0.0, "nearest")
    assert np.all(out[5:9, 5:9] == 200) and int((out != 7).sum()) == 16
    assert np.array_equal(zo.insert(base, np.zeros((0, 0), np.uint8), (1.0, 1.0, 5.0, 5.0), 0.3, "bilinear"), base)
