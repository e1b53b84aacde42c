"""Image.convert on the device (reference image.zig:396-421 / color.zig:108-151): all 25 (source, target) pairs of the five
pixel formats, bit-exact against the oracle, including out-of-range floats, strided views and the exhaustive 8-bit luma table."""
import numpy as np
import pytest

import oracle_lib as zo
from gpu_utils import rand_image

pytestmark = pytest.mark.gpu

SHAPES = {0: ((), np.uint8), 1: ((), np.float32), 2: ((3,), np.uint8), 3: ((4,), np.uint8), 4: ((4,), np.float32)}


@pytest.fixture(scope="module")
def zb():
    import torch
    assert torch.cuda.is_available()
    import zignal_b200 as zb
    return zb


@pytest.mark.parametrize("src_fmt", range(5))
@pytest.mark.parametrize("dst_fmt", range(5))
def test_convert_all_pairs(zb, src_fmt, dst_fmt):
    rng = np.random.default_rng(src_fmt * 5 + dst_fmt)
    tail, dtype = SHAPES[src_fmt]
    img = rand_image(rng, (37, 53) + tail, dtype)
    if dtype == np.float32:
        img = (img * 1.6 - 0.3).astype(np.float32)             # values below 0 and above 1 exercise the clamps
        img.flat[::17] = np.float32(0.5)                       # 127.5: the round-half-away tie
    got = zb.Image.from_numpy(img).convert(zb.PixFmt(dst_fmt)).to_numpy()
    want = zo.convert(img, dst_fmt)
    assert got.dtype == want.dtype and got.shape == want.shape
    assert np.array_equal(got, want), (src_fmt, dst_fmt, int((got != want).sum()))
    if src_fmt != dst_fmt:
        assert zb.lib().zb_last_kernel().decode() == "convert"


def test_convert_exhaustive_tables_views_and_errors(zb):
    r, g = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    for b in (0, 1, 77, 128, 254, 255):                          # 6 x 65536 colours through the integer luma and the /255 path
        rgb = np.stack([r, g, np.full_like(r, b)], axis=-1)
        dev = zb.Image.from_numpy(rgb)
        assert np.array_equal(dev.convert(zb.PixFmt.U8).to_numpy(), zo.convert(rgb, 0))
        assert np.array_equal(dev.convert(zb.PixFmt.F32).to_numpy(), zo.convert(rgb, 1))
    ramp = np.linspace(-0.01, 1.01, 70000, dtype=np.float32).reshape(280, 250)
    dev = zb.Image.from_numpy(ramp)
    for dst in (0, 2, 3):
        assert np.array_equal(dev.convert(zb.PixFmt(dst)).to_numpy(), zo.convert(ramp, dst)), dst
    rng = np.random.default_rng(1)
    big = rand_image(rng, (60, 80, 4), np.uint8)
    view = zb.Image.from_numpy(big).view(zb.Rectangle(5, 3, 71, 44))
    out_big = zb.Image.init(60, 80, zb.PixFmt.RGBAF32, device="cuda")
    out_view = out_big.view(zb.Rectangle(2, 7, 68, 48))
    view.convert(zb.PixFmt.RGBAF32, out=out_view)
    assert np.array_equal(out_big.to_numpy()[7:48, 2:68], zo.convert(np.ascontiguousarray(big[3:44, 5:71]), 4))
    with pytest.raises(zb.ZignalError) as ei:
        zb.Image.from_numpy(big).convert(zb.PixFmt.U8, out=zb.Image.init(60, 81, zb.PixFmt.U8, device="cuda"))
    assert ei.value.name == "DimensionMismatch"
    empty = zb.Image.init(0, 0, zb.PixFmt.RGB8, device="cuda")
    assert empty.convert(zb.PixFmt.F32).rows == 0
