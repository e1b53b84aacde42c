"""callers that are compositions of the hot path, device-resident across the chain, against the same
compositions of the oracle's functions (bit-exact for 8-bit formats)."""
import numpy as np
import pytest

import oracle_lib as zo
from gpu_utils import rand_image

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def zb():
    import torch
    assert torch.cuda.is_available()
    import zignal_b200 as zb
    return zb


@pytest.mark.parametrize("shape", [(97, 131), (120, 90, 3), (150, 160, 4)])
def test_image_pyramid_levels(zb, shape):
    from zignal_b200.compose import ImagePyramid
    rng = np.random.default_rng(len(shape) + shape[0])
    img = rand_image(rng, shape, np.uint8)
    pyr = ImagePyramid.build(zb.Image.from_numpy(img), 8, 1.2, 1.6)            # buildDefault parameters
    plan = ImagePyramid.level_plan(shape[0], shape[1], 8, 1.2, 1.6)
    assert pyr.n_levels == len(plan) + 1 and pyr.levels[0].rows == shape[0]
    for (rows, cols, sigma), level in zip(plan, pyr.levels[1:]):
        base = zo.gaussian_blur(img, sigma) if sigma is not None else img
        want = zo.resize(base, (rows, cols), "bilinear")
        assert (level.rows, level.cols) == (rows, cols)
        assert np.array_equal(level.to_numpy(), want), (rows, cols, sigma)
    # truncation when a level would drop below 8 pixels (pyramid.zig:61-71)
    small = ImagePyramid.build(zb.Image.from_numpy(img[:20, :20].copy()), 8, 1.5, 1.6)
    assert small.n_levels == 3 and small.levels[-1].rows == 8


@pytest.mark.parametrize("shape", [(60, 75), (48, 50, 4)])
@pytest.mark.parametrize("distance", [1, 4, 9])
def test_motion_blur_linear_axis_aligned(zb, shape, distance):
    from zignal_b200.compose import motion_blur_linear
    rng = np.random.default_rng(distance)
    img = rand_image(rng, shape, np.uint8)
    dev = zb.Image.from_numpy(img)
    k = np.full(distance, np.float32(1.0) / np.float32(distance), np.float32)
    one = np.ones(1, np.float32)
    got_h = motion_blur_linear(dev, zb.Image.init_like(dev), 0.0, distance).to_numpy()
    got_v = motion_blur_linear(dev, zb.Image.init_like(dev), float(np.float32(np.pi / 2)), distance).to_numpy()
    assert np.array_equal(got_h, zo.conv_separable(img, k, one, "replicate"))
    assert np.array_equal(got_v, zo.conv_separable(img, one, k, "replicate"))
    assert np.array_equal(motion_blur_linear(dev, zb.Image.init_like(dev), 0.3, 0).to_numpy(), img)   # distance 0: copy


MOTION_CASES = [((37, 45), np.uint8), ((40, 33, 3), np.uint8), ((29, 31, 4), np.uint8), ((35, 36), np.float32), ((23, 27, 4), np.float32),
                ((1, 9), np.uint8), ((7, 1), np.uint8)]


@pytest.mark.parametrize("shape,dtype", MOTION_CASES)
def test_motion_blur_linear_diagonal(zb, shape, dtype):
    """The line-integral branch of MotionBlur.linear (motion_blur.zig:115-250): bit-exact for every pixel type, u8 and f32."""
    from zignal_b200.compose import motion_blur_linear
    rng = np.random.default_rng(shape[0] + 31 * shape[1])
    img = rand_image(rng, shape, dtype)
    dev = zb.Image.from_numpy(img)
    for angle, distance in [(np.pi / 4, 3), (0.7, 5), (-1.1, 12), (2.5, 1), (0.0011, 7), (3.0, 40)]:
        got = motion_blur_linear(dev, zb.Image.init_like(dev), float(angle), distance).to_numpy()
        assert zb.lib().zb_last_kernel().decode() == "motion_line"
        assert np.array_equal(got, zo.motion_blur_linear(img, float(angle), distance)), (angle, distance)
    with pytest.raises(zb.ZignalError):
        motion_blur_linear(dev, dev, 0.7, 5)                      # a gather cannot run in place


@pytest.mark.parametrize("shape,dtype", MOTION_CASES)
def test_motion_blur_radial(zb, shape, dtype):
    """radial_zoom is bit-exact; radial_spin evaluates atan2 / cos / sin per sample, so its interior is compared at 1e-5 (f32) / one 8-bit
    step on at most 0.5 % of the samples (motion_blur.zig:252-436)."""
    from zignal_b200.compose import motion_blur_radial
    rng = np.random.default_rng(shape[0] * 5 + shape[1])
    img = rand_image(rng, shape, dtype)
    dev = zb.Image.from_numpy(img)
    for cx, cy, strength in [(0.5, 0.5, 0.5), (0.2, 0.8, 1.0), (0.0, 0.0, 0.3), (1.0, 0.4, 0.05), (0.5, 0.5, 7.0)]:
        got = motion_blur_radial(dev, zb.Image.init_like(dev), cx, cy, strength).to_numpy()
        assert zb.lib().zb_last_kernel().decode() == "motion_zoom"
        assert np.array_equal(got, zo.motion_blur_radial(img, cx, cy, strength)), (cx, cy, strength)
        got = motion_blur_radial(dev, zb.Image.init_like(dev), cx, cy, strength, spin=True).to_numpy()
        want = zo.motion_blur_radial(img, cx, cy, strength, spin=True)
        # On the outermost ring the t = 0 sample is the pixel itself recomputed as centre + distance * (cos, sin)(atan2(dy, dx)): it lands
        # within an ulp of the image edge, so whether the reference's bounds test (:323) keeps it depends on libm's last bit -- in the
        # reference as much as here.  The interior has no such discontinuity.
        gi, wi = got[1:-1, 1:-1], want[1:-1, 1:-1]
        if gi.size == 0:
            continue
        if dtype == np.uint8:
            diff = np.abs(gi.astype(int) - wi.astype(int))
            assert diff.max() <= 1 and (diff != 0).mean() <= 0.005, (cx, cy, strength, int(diff.max()), float((diff != 0).mean()))
        else:
            # an angle error of a few f32 ulps moves a sample by ~1e-7 * distance pixels: <= 2e-5 of the value range on these image sizes
            assert np.allclose(gi, wi, rtol=1e-5, atol=2e-5), (cx, cy, strength, float(np.abs(gi - wi).max()))
    assert np.array_equal(motion_blur_radial(dev, zb.Image.init_like(dev), 0.5, 0.5, 0.0).to_numpy(), img)       # strength 0: copy


def test_motion_blur_reference_cases(zb):
    """image/tests/filters.zig:1021-1160 through the device path."""
    from zignal_b200.compose import motion_blur_linear, motion_blur_radial
    spot = np.zeros((5, 5), np.uint8)
    spot[2, 2] = 255
    d = zb.Image.from_numpy(spot)
    b = motion_blur_linear(d, zb.Image.init_like(d), float(np.float32(np.pi / 4)), 3).to_numpy()
    assert b[1, 1] > 0 and b[2, 2] > 0 and b[3, 3] > 0
    yy, xx = np.mgrid[0:7, 0:7].astype(np.float32)
    dist = np.sqrt((xx - 3) ** 2 + (yy - 3) ** 2)
    ring = np.where((dist > 1.5) & (dist < 2.5), 255, 0).astype(np.uint8)
    dr = zb.Image.from_numpy(ring)
    z = motion_blur_radial(dr, zb.Image.init_like(dr), 0.5, 0.5, 0.5).to_numpy()
    assert abs(int(z[3, 3]) - int(ring[3, 3])) < 20
    pt = np.zeros((7, 7), np.uint8)
    pt[2, 4] = 255
    dp = zb.Image.from_numpy(pt)
    sp = motion_blur_radial(dp, zb.Image.init_like(dp), 0.5, 0.5, 0.5, spin=True).to_numpy()
    assert sp[2, 4] > 0 and int((sp > 0).sum()) > 1


@pytest.mark.parametrize("src_shape,dst_shape", [((60, 100, 3), (64, 64)), ((100, 60, 4), (48, 90)), ((40, 40, 3), (80, 80)), ((30, 50, 3), (30, 50))])
@pytest.mark.parametrize("method", ["bilinear", "bicubic", "nearest"])
def test_letterbox(zb, src_shape, dst_shape, method):
    from zignal_b200.compose import letterbox
    from gpu_utils import method_enum
    rng = np.random.default_rng(src_shape[0] + dst_shape[1])
    img = rand_image(rng, src_shape, np.uint8)
    out = zb.Image.from_numpy(np.full(dst_shape + (src_shape[2],), 0xAB, np.uint8))
    rect = letterbox(zb.Image.from_numpy(img), out, method_enum(zb, method))
    got = out.to_numpy()
    rs, cs = np.float32(dst_shape[0]) / np.float32(src_shape[0]), np.float32(dst_shape[1]) / np.float32(src_shape[1])
    if src_shape[:2] == dst_shape:
        assert np.array_equal(got, img) and (rect.l, rect.t, rect.r, rect.b) == (0, 0, dst_shape[1], dst_shape[0])
    elif rs == cs:
        assert np.array_equal(got, zo.resize(img, dst_shape, method))
    else:
        a = min(rs, cs)
        sr = int(np.floor(np.float64(np.float32(a * np.float32(src_shape[0]))) + 0.5))
        sc = int(np.floor(np.float64(np.float32(a * np.float32(src_shape[1]))) + 0.5))
        orow, ocol = (dst_shape[0] - sr) // 2, (dst_shape[1] - sc) // 2
        assert (rect.l, rect.t, rect.r, rect.b) == (ocol, orow, ocol + sc, orow + sr)
        want = np.zeros_like(got)
        want[orow:orow + sr, ocol:ocol + sc] = zo.resize(img, (sr, sc), method)
        assert np.array_equal(got, want)


@pytest.mark.parametrize("shape,dtype", [((37, 45), np.uint8), ((40, 33, 3), np.uint8), ((29, 31, 4), np.uint8), ((35, 36), np.float32), ((1, 7), np.uint8),
                                         ((300, 1000), np.uint8), ((130, 777), np.uint8), ((140, 530, 3), np.uint8), ((97, 600, 4), np.uint8)])
def test_sobel(zb, shape, dtype):
    rng = np.random.default_rng(shape[0])
    img = rand_image(rng, shape, dtype)
    if dtype == np.float32:
        img = (img * 255.0).astype(np.float32)     # float scalars pass through un-normalised (edges.zig:38-48)
    got = zb.Image.from_numpy(img).sobel().to_numpy()
    gray8 = dtype == np.uint8            # every 8-bit input (gray, Rgb, Rgba: luma inside the loader) takes the byte-tile kernel
    assert zb.lib().zb_last_kernel().decode() == ("sobel_tile_u8" if gray8 else "sobel_fused")
    assert np.array_equal(got, zo.sobel(img))
    if gray8:
        zb.lib().zb_tune(b"sobel.tile", 0)
        try:
            per_px = zb.Image.from_numpy(img).sobel().to_numpy()
            assert zb.lib().zb_last_kernel().decode() == "sobel_fused"
        finally:
            zb.lib().zb_tune(b"sobel.tile", 1)
        assert np.array_equal(per_px, got)
    zb.lib().zb_set_force_generic(1)          # the three-kernel composition must agree with the one-pass kernel
    try:
        comp = zb.Image.from_numpy(img).sobel().to_numpy()
        assert zb.lib().zb_last_kernel().decode() == "sobel"
    finally:
        zb.lib().zb_set_force_generic(0)
    assert np.array_equal(comp, got)
    step = np.tile(np.where(np.arange(5) < 2, 0, 255).astype(np.uint8), (5, 1))    # the reference's own test image
    e = zb.Image.from_numpy(step).sobel().to_numpy()
    assert e[2, 2] > 200 and e[2, 0] < 50


def _structured(rng, rows, cols):
    """Smooth blobs plus a little noise: long contours whose strength varies along them, so hysteresis has work to do."""
    y, x = np.mgrid[0:rows, 0:cols].astype(np.float64)
    img = np.zeros((rows, cols))
    for _ in range(12):
        cy, cx, rad, amp = rng.uniform(0, rows), rng.uniform(0, cols), rng.uniform(15, min(rows, cols) / 3), rng.uniform(20, 90)
        img += amp * (np.hypot(y - cy, x - cx) < rad)
    img += rng.normal(0, 2.0, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


@pytest.mark.parametrize("shape,dtype", [((37, 45), np.uint8), ((40, 33, 3), np.uint8), ((29, 31, 4), np.uint8), ((35, 36), np.float32),
                                         ((3, 3), np.uint8), ((2, 40), np.uint8), ((1, 1), np.uint8)])
def test_canny_matches_oracle(zb, shape, dtype):
    """Image.canny (edges.zig:212-274), bit-exact: every stage reproduces the oracle's f32 values (generic convolution paths, separately
    rounded magnitude), so the thresholded decisions agree pixel for pixel."""
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    img = rand_image(rng, shape, dtype)
    for sigma, low, high in [(0.0, 150.0, 400.0), (1.0, 20.0, 60.0), (1.4, 10.0, 30.0), (0.5, 0.0, 1.0)]:
        got = zb.Image.from_numpy(img).canny(sigma, low, high).to_numpy()
        want = zo.canny(img, sigma, low, high)
        assert zb.lib().zb_last_kernel().decode() == "canny"
        assert np.array_equal(got, want), (sigma, low, high, int((got != want).sum()))
        assert set(np.unique(got)) <= {0, 255}


def test_canny_reference_cases_and_validation(zb):
    """image/tests/filters.zig:1182-1300 through the device path."""
    step = np.tile(np.where(np.arange(10) < 5, 0, 255).astype(np.uint8), (10, 1))
    e = zb.Image.from_numpy(step).canny(1.0, 50, 100).to_numpy()
    assert e.shape == step.shape and e[:, 4:7].any()
    assert np.array_equal(e, zo.canny(step, 1.0, 50, 100))
    rgb = np.zeros((8, 8, 3), np.uint8)
    rgb[:, :4, 0] = 255
    rgb[:, 4:, 1] = 255
    e = zb.Image.from_numpy(rgb).canny(1.0, 30, 90).to_numpy()
    assert e[:, 3:6].any() and np.array_equal(e, zo.canny(rgb, 1.0, 30, 90))
    ramp = zb.Image.from_numpy((np.arange(5)[:, None] * 10 + np.arange(5)[None, :]).astype(np.uint8))
    ramp.canny(0.0, 50, 100)
    for args, name in [((-1.0, 50, 100), "InvalidSigma"), ((1.0, -1, 100), "InvalidThreshold"), ((1.0, 50, -1), "InvalidThreshold"),
                       ((1.0, 100, 50), "InvalidThreshold"), ((np.nan, 50, 100), "InvalidArgument"), ((1.0, np.nan, 100), "InvalidArgument"),
                       ((1.0, 50, np.nan), "InvalidArgument"), ((np.inf, 50, 100), "InvalidArgument"), ((1.0, np.inf, 100), "InvalidArgument"),
                       ((1.0, 50, np.inf), "InvalidArgument"), ((-np.inf, 50, 100), "InvalidArgument")]:
        with pytest.raises(zb.ZignalError) as ei:
            ramp.canny(*args)
        assert ei.value.name == name, args
    with pytest.raises(zb.ZignalError) as ei:
        ramp.canny(1.0, 10, 20, out=zb.Image.init(4, 5, zb.PixFmt.U8, device="cuda"))
    assert ei.value.name == "DimensionMismatch"


@pytest.mark.parametrize("sigma,low,high", [(2.0, 4.0, 12.0), (1.0, 10.0, 40.0), (0.0, 30.0, 200.0)])
def test_canny_hysteresis_across_tiles(zb, sigma, low, high):
    """Contours several hundred pixels long cross many 64x64 relaxation tiles; the fixed point must equal the breadth-first closure."""
    rng = np.random.default_rng(int(high))
    img = _structured(rng, 389, 523)
    got = zb.Image.from_numpy(img).canny(sigma, low, high).to_numpy()
    want = zo.canny(img, sigma, low, high)
    assert np.array_equal(got, want), int((got != want).sum())
    assert 0 < int(want.sum()) // 255 < want.size // 4


def test_canny_long_weak_chain(zb):
    """One seed at the end of a serpentine of candidates that winds through every tile: thousands of promotions along a single path,
    the worst case for the tile relaxation.  Built directly on the magnitude by a ridge image: a bright one-pixel-wide path on black."""
    rows, cols = 200, 330
    img = np.zeros((rows, cols), np.uint8)
    level = 60
    for k, r in enumerate(range(4, rows - 4, 8)):                # horizontal runs joined alternately at the right and left ends
        img[r, 4:cols - 4] = level
        if r + 8 < rows - 4:
            c = cols - 5 if k % 2 == 0 else 4
            img[r:r + 9, c] = level
    img[4, 4:12] = 255                                           # the only strong stretch
    got = zb.Image.from_numpy(img).canny(0.0, 100.0, 600.0).to_numpy()
    want = zo.canny(img, 0.0, 100.0, 600.0)
    assert np.array_equal(got, want), int((got != want).sum())
    assert int(want[rows - 20:, :].sum()) > 0                    # the closure really reached the far end of the serpentine
