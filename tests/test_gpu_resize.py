"""Image.resize on the GPU vs the oracle: integer plane resizers for Rgb/Rgba u8 (bit-exact) and the
generic per-pixel samplers for u8 / f32 / RGBA f32."""
import numpy as np
import pytest

import oracle_lib as zo
from gpu_utils import METHODS, method_enum, rand_image, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def zb():
    import torch
    assert torch.cuda.is_available()
    import zignal_b200 as zb
    return zb


SHAPES = [((16, 16), (4, 4)), ((33, 47), (70, 19)), ((64, 48), (16, 12)), ((7, 5), (40, 31)), ((50, 60), (51, 59)), ((1, 1), (5, 7)), ((9, 13), (1, 1))]


@pytest.mark.parametrize("ch", [3, 4])
@pytest.mark.parametrize("method", METHODS)
def test_plane_resize_bit_exact(zb, ch, method):
    rng = np.random.default_rng(ch * 10 + METHODS.index(method))
    for src_shape, dst_shape in SHAPES:
        img = rand_image(rng, src_shape + (ch,), np.uint8)
        out = zb.Image.init(dst_shape[0], dst_shape[1], zb.Image.from_numpy(img).pixfmt)
        got = zb.Image.from_numpy(img).resize(out, method_enum(zb, method)).to_numpy()
        assert np.array_equal(got, zo.resize(img, dst_shape, method)), (src_shape, dst_shape)


@pytest.mark.parametrize("fmt", ["u8", "f32", "rgbaf32"])
@pytest.mark.parametrize("method", METHODS)
def test_generic_resize(zb, fmt, method):
    rng = np.random.default_rng(METHODS.index(method))
    for src_shape, dst_shape in SHAPES:
        shape = src_shape + ((4,) if fmt == "rgbaf32" else ())
        img = rand_image(rng, shape, np.uint8 if fmt == "u8" else np.float32)
        dev = zb.Image.from_numpy(img)
        out = zb.Image.init(dst_shape[0], dst_shape[1], dev.pixfmt)
        for b, c in [(1 / 3, 1 / 3), (0.0, 0.5)] if method == "mitchell" else [(1 / 3, 1 / 3)]:
            got = dev.resize(out, method_enum(zb, method), b, c).to_numpy()
            want = zo.resize(img, dst_shape, method, b, c)
            if fmt == "u8":
                assert np.array_equal(got, want), (src_shape, dst_shape)
            else:
                assert rel_err(got, want) <= 1e-5, (src_shape, dst_shape)
                assert np.array_equal(got, want), "f32 samplers are expected to be bit-identical (unfused f32, same order)"


def test_same_shape_copy_views_and_host_twin(zb):
    rng = np.random.default_rng(3)
    img = rand_image(rng, (20, 30, 3), np.uint8)
    dev = zb.Image.from_numpy(img)
    assert np.array_equal(dev.resize(zb.Image.init_like(dev), zb.Interpolation.LANCZOS).to_numpy(), img)  # interpolation.zig:91-108
    v = dev.view(zb.Rectangle(4, 2, 28, 18))
    crop = np.ascontiguousarray(img[2:18, 4:28])
    big = zb.Image.from_numpy(np.full((40, 50, 3), 5, np.uint8))
    ov = big.view(zb.Rectangle(1, 1, 41, 33))
    v.resize(ov, zb.Interpolation.BICUBIC)
    assert np.array_equal(big.to_numpy()[1:33, 1:41], zo.resize(crop, (32, 40), "bicubic"))
    assert np.array_equal(zb.host_resize(img, (9, 11), zb.Interpolation.BILINEAR), zo.resize(img, (9, 11), "bilinear"))
    s = dev.scale(0.5, zb.Interpolation.BILINEAR)  # tests/resize.zig:258-298
    assert (s.rows, s.cols) == (10, 15)
    assert np.array_equal(s.to_numpy(), zo.resize(img, (10, 15), "bilinear"))


@pytest.mark.parametrize("ch", [3, 4])
@pytest.mark.parametrize("method", ["bicubic", "catmull_rom", "mitchell"])
def test_uniform_phase_cubic_kernel(zb, ch, method):
    """Integer down-scale ratios give every column (row) the same tap weights: the constant-weight kernel must agree with the
    oracle bit for bit, including odd row pitches (unaligned 32-bit windows), mirrored edge taps and strided views."""
    rng = np.random.default_rng(ch + len(method))
    for src_shape, dst_shape in [((96, 99), (32, 33)), ((120, 100), (60, 50)), ((64, 64), (16, 16)), ((40, 35), (8, 7)), ((12, 9), (4, 3))]:
        img = rand_image(rng, src_shape + (ch,), np.uint8)
        dev = zb.Image.from_numpy(img)
        out = zb.Image.init(dst_shape[0], dst_shape[1], dev.pixfmt)
        got = dev.resize(out, method_enum(zb, method)).to_numpy()
        assert zb.lib().zb_last_kernel().decode() in ("resize_cubic_uniform_u8", "resize_cubic_r4_u8"), (src_shape, dst_shape)
        assert np.array_equal(got, zo.resize(img, dst_shape, method)), (src_shape, dst_shape)
    # exact 4:1 with 16-byte friendly rows: the coalesced strip kernel (full warps, a ragged last warp, several row blocks)
    if method == "bicubic":
        for src_shape, dst_shape in [((64, 2048), (16, 512)), ((40, 4144), (10, 1036)), ((256, 528), (64, 132))]:
            img = rand_image(rng, src_shape + (ch,), np.uint8)
            dev = zb.Image.from_numpy(img)
            got = dev.resize(zb.Image.init(dst_shape[0], dst_shape[1], dev.pixfmt), method_enum(zb, method)).to_numpy()
            assert zb.lib().zb_last_kernel().decode() == "resize_cubic_r4_u8", (src_shape, dst_shape)
            assert np.array_equal(got, zo.resize(img, dst_shape, method)), (src_shape, dst_shape)
    # a view whose rows start on odd bytes, into a view of a larger destination
    base = rand_image(rng, (70, 67, ch), np.uint8)
    dev = zb.Image.from_numpy(base)
    v = dev.view(zb.Rectangle(3, 5, 63, 65))                       # 60 x 60 window
    big = zb.Image.from_numpy(np.full((40, 41, ch), 9, np.uint8))
    ov = big.view(zb.Rectangle(2, 1, 32, 31))                      # 30 x 30 window
    v.resize(ov, method_enum(zb, method))
    want = zo.resize(np.ascontiguousarray(base[5:65, 3:63]), (30, 30), method)
    res = big.to_numpy()
    assert np.array_equal(res[1:31, 2:32], want)
    res[1:31, 2:32] = 9
    assert np.all(res == 9)


def test_config3_bicubic_4to1_rgb(zb):
    """BASELINE config 3 at reduced size for the oracle (2048^2 -> 512^2) plus the 4:1 constant-weight property at
    full size 16384^2 -> 4096^2: every output is clamp(trunc(sum(block * W) / 256)) with W = outer([-32,160,160,-32])/256."""
    import torch
    rng = np.random.default_rng(3)
    img = rand_image(rng, (2048, 2048, 3), np.uint8)
    dev = zb.Image.from_numpy(img)
    got = dev.resize(zb.Image.init(512, 512, dev.pixfmt), zb.Interpolation.BICUBIC).to_numpy()
    assert np.array_equal(got, zo.resize(img, (512, 512), "bicubic"))
    gen = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randint(0, 256, (16384, 16384, 3), device="cuda", dtype=torch.uint8, generator=gen)
    big = zb.Image.from_tensor(x)
    out = big.resize(zb.Image.init(4096, 4096, big.pixfmt), zb.Interpolation.BICUBIC).tensor()
    w1 = torch.tensor([-32, 160, 160, -32], device="cuda", dtype=torch.int64)
    w = (torch.outer(w1, w1) // 256)  # all products are multiples of 256
    blocks = x.view(4096, 4, 4096, 4, 3).permute(0, 2, 4, 1, 3).to(torch.int64)  # (R, C, ch, 4, 4)
    s = (blocks * w).sum(dim=(-1, -2))
    q = torch.where(s >= 0, s // 256, -((-s) // 256)).clamp(0, 255).to(torch.uint8)
    assert torch.equal(out, q)
