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
    with pytest.raises(zb.ZignalError):
        motion_blur_linear(dev, zb.Image.init_like(dev), 0.7, 5)


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


@pytest.mark.parametrize("shape,dtype", [((37, 45), np.uint8), ((40, 33, 3), np.uint8), ((29, 31, 4), np.uint8), ((35, 36), np.float32), ((1, 7), np.uint8)])
def test_sobel(zb, shape, dtype):
    rng = np.random.default_rng(shape[0])
    img = rand_image(rng, shape, dtype)
    if dtype == np.float32:
        img = (img * 255.0).astype(np.float32)     # float scalars pass through un-normalised (edges.zig:38-48)
    got = zb.Image.from_numpy(img).sobel().to_numpy()
    assert zb.lib().zb_last_kernel().decode() == "sobel_fused"
    assert np.array_equal(got, zo.sobel(img))
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
