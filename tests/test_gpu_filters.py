"""Box blur / sharpen / integral image on the GPU vs the oracle: bit-exact for every pixel format,
including images whose f32 summed-area table exceeds 2^24 (where the summation order is part of the
result, reference integral.zig:41-78)."""
import ctypes as C

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


CASES = [((5, 5), np.uint8), ((21, 13), np.uint8), ((64, 70, 3), np.uint8), ((33, 47, 4), np.uint8), ((40, 37), np.float32),
         ((31, 29, 4), np.float32), ((1, 9), np.uint8), ((9, 1), np.uint8), ((600, 700), np.uint8), ((520, 530, 4), np.uint8)]


@pytest.mark.parametrize("shape,dtype", CASES)
@pytest.mark.parametrize("radius", [1, 2, 3, 7, 50])
def test_box_blur_bit_exact(zb, shape, dtype, radius):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1] + radius)
    img = rand_image(rng, shape, dtype)
    if shape[0] >= 500 and dtype == np.uint8:
        img |= 0xC0  # bright: SAT values pass 2^24, f32 adds become inexact
    got = zb.Image.from_numpy(img).box_blur(radius).to_numpy()
    assert np.array_equal(got, zo.box_blur(img, radius))


@pytest.mark.parametrize("shape,dtype", CASES)
@pytest.mark.parametrize("radius", [1, 3, 9])
def test_sharpen_bit_exact(zb, shape, dtype, radius):
    rng = np.random.default_rng(shape[0] * 77 + shape[1] + radius)
    img = rand_image(rng, shape, dtype)
    got = zb.Image.from_numpy(img).sharpen(radius).to_numpy()
    assert np.array_equal(got, zo.sharpen(img, radius))


@pytest.mark.parametrize("ch", [1, 4])
@pytest.mark.parametrize("radius", [1, 3, 6, 15])
def test_box_fused_strips_bands_and_ragged_views(zb, ch, radius):
    """The fused 8-bit path (row-offset table, checkpointed column chains, shared-memory SAT ring): several strips and bands,
    a ragged last unit (gray view whose width is not a multiple of 4), against the oracle and against the SAT-in-HBM path."""
    import torch
    L = zb.lib()
    rng = np.random.default_rng(radius * 10 + ch)
    shape = (333, 1008) if ch == 1 else (333, 300, 4)
    base = rand_image(rng, shape, np.uint8) | 0xE0
    view_cols = 1003 if ch == 1 else 297
    dev = zb.Image.from_numpy(base)
    v = dev.view(zb.Rectangle(0, 0, view_cols, shape[0]))
    host = np.ascontiguousarray(base[:, :view_cols])
    out = zb.Image.from_numpy(np.zeros_like(base))
    ov = out.view(zb.Rectangle(0, 0, view_cols, shape[0]))
    v.box_blur(radius, out=ov)
    assert L.zb_last_kernel().decode() == "box_fused_blur"
    assert np.array_equal(out.to_numpy()[:, :view_cols], zo.box_blur(host, radius))
    assert not out.to_numpy()[:, view_cols:].any()
    v.sharpen(radius, out=ov)
    assert L.zb_last_kernel().decode() == "box_fused_sharpen"
    assert np.array_equal(out.to_numpy()[:, :view_cols], zo.sharpen(host, radius))
    # large image: fused vs the three-kernel SAT path (both on the GPU)
    g = torch.Generator(device="cuda").manual_seed(radius)
    big = torch.randint(128, 256, (3000, 2048, 4) if ch == 4 else (3000, 4096), device="cuda", dtype=torch.uint8, generator=g)
    bi = zb.Image.from_tensor(big)
    a = bi.box_blur(radius).tensor().clone()
    a2 = bi.sharpen(radius).tensor().clone()
    L.zb_set_force_generic(1)
    try:
        b = bi.box_blur(radius).tensor().clone()
        assert L.zb_last_kernel().decode() == "sat_box_blur"
        b2 = bi.sharpen(radius).tensor().clone()
    finally:
        L.zb_set_force_generic(0)
    assert torch.equal(a, b) and torch.equal(a2, b2)


def test_integral_plane_bit_exact(zb):
    import torch
    rng = np.random.default_rng(3)
    for shape, dtype in [((21, 13), np.uint8), ((300, 517), np.uint8), ((129, 65), np.float32), ((1000, 1100), np.uint8)]:
        img = rand_image(rng, shape, dtype)
        if shape[0] >= 1000:
            img |= 0xF0
        dev = zb.Image.from_numpy(img)
        sat = torch.empty(shape, dtype=torch.float32, device="cuda")
        a = dev._zb()
        zb._ffi.check(zb.lib().zb_integral_plane(a, int(dev.pixfmt), C.cast(sat.data_ptr(), C.POINTER(C.c_float)), None))
        torch.cuda.synchronize()
        assert np.array_equal(sat.cpu().numpy(), zo.integral_plane(img))
    ones = np.ones((21, 13), np.uint8)  # tests/integral.zig:11-28
    r, c = np.mgrid[0:21, 0:13]
    assert np.array_equal(zo.integral_plane(ones), ((r + 1) * (c + 1)).astype(np.float32))


def test_reference_kats_and_views(zb):
    img = np.full((5, 5), 128, np.uint8)  # filters.zig:87-103
    assert np.all(zb.Image.from_numpy(img).box_blur(1).to_numpy() == 128)
    img = np.full((12, 12), 200, np.uint8)  # :186-207
    assert np.all(zb.Image.from_numpy(img).box_blur(3).to_numpy() == 200)
    img = np.full((5, 5), 100, np.uint8)  # :327-343
    assert np.all(zb.Image.from_numpy(img).sharpen(1).to_numpy() == 100)
    rng = np.random.default_rng(4)
    for size in (8, 32):  # :234-264 alpha preserved
        im = rand_image(rng, (size, size, 4), np.uint8)
        im[..., 3] = 255
        for r in (1, 3):
            assert np.all(zb.Image.from_numpy(im).box_blur(r).to_numpy()[..., 3] == 255)
    # radius 0 = copy, views on both sides (:50-75, :105-126)
    base = rand_image(rng, (30, 40, 3), np.uint8)
    dev = zb.Image.from_numpy(base)
    v = dev.view(zb.Rectangle(5, 4, 35, 24))
    out_big = zb.Image.from_numpy(np.full((30, 40, 3), 9, np.uint8))
    ov = out_big.view(zb.Rectangle(2, 3, 32, 23))
    v.box_blur(0, out=ov)
    assert np.array_equal(out_big.to_numpy()[3:23, 2:32], base[4:24, 5:35])
    v.box_blur(2, out=ov)
    full = out_big.to_numpy()
    assert np.array_equal(full[3:23, 2:32], zo.box_blur(np.ascontiguousarray(base[4:24, 5:35]), 2))
    mask = np.ones((30, 40), bool)
    mask[3:23, 2:32] = False
    assert np.all(full[mask] == 9)
    # in place (examples/src/face_alignment.zig:95,97 use boxBlur / sharpen in place)
    img = rand_image(rng, (50, 60, 4), np.uint8)
    d = zb.Image.from_numpy(img)
    d.box_blur(2, out=d)
    assert np.array_equal(d.to_numpy(), zo.box_blur(img, 2))
    d = zb.Image.from_numpy(img)
    d.sharpen(2, out=d)
    assert np.array_equal(d.to_numpy(), zo.sharpen(img, 2))
    # host twins
    assert np.array_equal(zb.host_box_blur(img, 3), zo.box_blur(img, 3))
    assert np.array_equal(zb.host_sharpen(img, 3), zo.sharpen(img, 3))


def test_config1_box_blur_512_u8(zb):
    """BASELINE config 1: 3x3 box blur (radius 1) on 512x512 u8."""
    rng = np.random.default_rng(1)
    img = rand_image(rng, (512, 512), np.uint8)
    assert np.array_equal(zb.Image.from_numpy(img).box_blur(1).to_numpy(), zo.box_blur(img, 1))


# ---- order-statistic filters (order_statistic_blur.zig) ----
@pytest.mark.parametrize("shape", [(37, 45), (40, 33, 3), (29, 31, 4), (1, 9), (7, 1), (3, 3)])
def test_order_statistic_filters_match_oracle(zb, shape):
    """Bit-exact against the oracle's sliding-histogram restatement: percentile (incl. median / min / max), midpoint and alpha-trimmed
    mean, every border mode, windows larger than the image included."""
    rng = np.random.default_rng(shape[0] * 13 + shape[1])
    img = rand_image(rng, shape, np.uint8)
    dev = zb.Image.from_numpy(img)
    for radius in (1, 2, 3, 5):
        for bname, border in [("zero", zb.BorderMode.ZERO), ("replicate", zb.BorderMode.REPLICATE), ("mirror", zb.BorderMode.MIRROR),
                              ("wrap", zb.BorderMode.WRAP)]:
            for pct in (0.0, 0.25, 0.5, 0.9, 1.0):
                got = dev.percentile_blur(radius, pct, border).to_numpy()
                assert zb.lib().zb_last_kernel().decode() == "order_statistic"
                assert np.array_equal(got, zo.order_blur(img, radius, "percentile", pct, bname)), (radius, bname, pct)
            assert np.array_equal(dev.midpoint_blur(radius, border).to_numpy(), zo.order_blur(img, radius, "midpoint", 0.0, bname)), (radius, bname)
            for trim in (0.0, 0.12, 0.33, 0.49):
                got = dev.alpha_trimmed_mean_blur(radius, trim, border).to_numpy()
                assert np.array_equal(got, zo.order_blur(img, radius, "alpha_trimmed", trim, bname)), (radius, bname, trim)
        assert np.array_equal(dev.median_blur(radius).to_numpy(), zo.order_blur(img, radius, "percentile", 0.5, "mirror"))
        assert np.array_equal(dev.min_blur(radius, zb.BorderMode.REPLICATE).to_numpy(), zo.order_blur(img, radius, "percentile", 0.0, "replicate"))
        assert np.array_equal(dev.max_blur(radius, zb.BorderMode.REPLICATE).to_numpy(), zo.order_blur(img, radius, "percentile", 1.0, "replicate"))


@pytest.mark.parametrize("shape", [(45, 61), (33, 40, 4)])
def test_order_statistic_register_and_tile_kernels_agree(zb, shape):
    """Radius 1-3 run kernels whose window lives in registers; zb_set_force_generic selects the any-radius kernel that re-reads the shared
    tile.  Both must give the oracle's result."""
    rng = np.random.default_rng(shape[1])
    img = rand_image(rng, shape, np.uint8)
    dev = zb.Image.from_numpy(img)
    for radius in (1, 2, 3):
        fast = [dev.percentile_blur(radius, 0.5, zb.BorderMode.WRAP).to_numpy(), dev.percentile_blur(radius, 0.81, zb.BorderMode.ZERO).to_numpy(),
                dev.midpoint_blur(radius, zb.BorderMode.MIRROR).to_numpy(), dev.alpha_trimmed_mean_blur(radius, 0.27, zb.BorderMode.REPLICATE).to_numpy()]
        zb.lib().zb_set_force_generic(1)
        try:
            slow = [dev.percentile_blur(radius, 0.5, zb.BorderMode.WRAP).to_numpy(), dev.percentile_blur(radius, 0.81, zb.BorderMode.ZERO).to_numpy(),
                    dev.midpoint_blur(radius, zb.BorderMode.MIRROR).to_numpy(), dev.alpha_trimmed_mean_blur(radius, 0.27, zb.BorderMode.REPLICATE).to_numpy()]
        finally:
            zb.lib().zb_set_force_generic(0)
        want = [zo.order_blur(img, radius, "percentile", 0.5, "wrap"), zo.order_blur(img, radius, "percentile", 0.81, "zero"),
                zo.order_blur(img, radius, "midpoint", 0.0, "mirror"), zo.order_blur(img, radius, "alpha_trimmed", 0.27, "replicate")]
        for f, g, w in zip(fast, slow, want):
            assert np.array_equal(f, w) and np.array_equal(g, w), radius


def test_order_statistic_reference_cases(zb):
    """image/tests/filters.zig:817-966 through the device path, plus aliasing, radius 0 and the error order."""
    imp = np.zeros((5, 5), np.uint8)
    imp[2, 2] = 255
    med = zb.Image.from_numpy(imp).median_blur(1).to_numpy()
    assert med[2, 2] == 0 and med[2, 1] == 0 and med[1, 2] == 0
    seq = np.arange(9, dtype=np.uint8).reshape(3, 3)
    dseq = zb.Image.from_numpy(seq)
    mx = dseq.percentile_blur(1, 1.0, zb.BorderMode.ZERO).to_numpy()
    assert mx[1, 1] == 8 and mx[0, 0] == 4
    rgb = np.tile(np.array([32, 64, 96], np.uint8), (3, 3, 1))
    rgb[1, 1] = (255, 0, 0)
    m = zb.Image.from_numpy(rgb).median_blur(1).to_numpy()
    assert tuple(m[1, 1]) == (32, 64, 96) and tuple(m[0, 0]) == (32, 64, 96)
    assert dseq.midpoint_blur(1, zb.BorderMode.REPLICATE).to_numpy()[1, 1] == 4
    assert dseq.alpha_trimmed_mean_blur(1, 0.12, zb.BorderMode.REPLICATE).to_numpy()[1, 1] == 4
    for trim in (0.6, 0.5, -0.1, float("nan")):
        with pytest.raises(zb.ZignalError) as ei:
            dseq.alpha_trimmed_mean_blur(1, trim, zb.BorderMode.REPLICATE)
        assert ei.value.name == "InvalidTrim"
    for pct in (-0.01, 1.01):
        with pytest.raises(zb.ZignalError) as ei:
            dseq.percentile_blur(1, pct)
        assert ei.value.name == "InvalidPercentile"
    assert np.array_equal(dseq.percentile_blur(0, 7.0).to_numpy(), seq)           # radius 0 copies before the percentile check
    with pytest.raises(zb.ZignalError) as ei:
        zb.Image.from_numpy(np.zeros((4, 4), np.float32)).median_blur(1)
    assert ei.value.name == "Unsupported"
    with pytest.raises(zb.ZignalError) as ei:
        dseq.median_blur(1, out=zb.Image.init(3, 4, zb.PixFmt.U8, device="cuda"))
    assert ei.value.name == "DimensionMismatch"
    rng = np.random.default_rng(5)
    img = rand_image(rng, (50, 70, 4), np.uint8)
    dev = zb.Image.from_numpy(img)
    dev.median_blur(2, out=dev)                                                   # in place (order_statistic_blur.zig:52-60)
    assert np.array_equal(dev.to_numpy(), zo.order_blur(img, 2, "percentile", 0.5, "mirror"))


def test_order_statistic_large(zb):
    """A size where the grid has many tiles in both directions and the widest supported window."""
    rng = np.random.default_rng(77)
    img = rand_image(rng, (301, 517), np.uint8)
    dev = zb.Image.from_numpy(img)
    assert np.array_equal(dev.median_blur(3).to_numpy(), zo.order_blur(img, 3, "percentile", 0.5, "mirror"))
    assert np.array_equal(dev.alpha_trimmed_mean_blur(4, 0.2, zb.BorderMode.WRAP).to_numpy(), zo.order_blur(img, 4, "alpha_trimmed", 0.2, "wrap"))
    small = np.ascontiguousarray(img[:64, :96])
    dsm = zb.Image.from_numpy(small)
    assert np.array_equal(dsm.percentile_blur(31, 0.37, zb.BorderMode.MIRROR).to_numpy(), zo.order_blur(small, 31, "percentile", 0.37, "mirror"))
    # radius > 31: the window tile opts in to the SM's full shared memory (the reference has no radius limit)
    assert np.array_equal(dsm.median_blur(40).to_numpy(), zo.order_blur(small, 40, "percentile", 0.5, "mirror"))
    rgba = rand_image(rng, (48, 64, 4), np.uint8)
    assert np.array_equal(zb.Image.from_numpy(rgba).max_blur(60, zb.BorderMode.REPLICATE).to_numpy(), zo.order_blur(rgba, 60, "percentile", 1.0, "replicate"))
    with pytest.raises(zb.ZignalError) as ei:
        dsm.median_blur(400)                                                       # beyond one SM's shared memory
    assert ei.value.name == "Unsupported"
