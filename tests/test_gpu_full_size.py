"""The BASELINE.json configurations at FULL size, whole output against the CPU oracle (all host threads).

C2: 15x15 Gaussian, 8192x8192 RGBA f32 (every pixel, 1e-5 relative; bit-identical in the exact build)
C3: bicubic 16384^2 -> 4096^2 Rgb u8 (bit-exact, every pixel)
C4: rotate 45 deg bilinear zero border on a 128-frame batch of 1920x1080 Rgba u8 (bit-exact, every frame)
C5: fdm.match on 4096x4096 Rgb u8 (the +-1 flips that the exact-integer statistics may cause are counted)
Reference: src/image/convolution.zig:441-647, src/image/channel_ops.zig:217-289, src/image/transforms.zig:163-212,
src/fdm.zig:141-273.
"""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as zo
from gpu_utils import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def zb():
    import torch
    assert torch.cuda.is_available()
    import zignal_b200 as zb
    zo.set_threads(zo.hw_threads())
    yield zb
    zo.set_threads(1)


def test_c2_whole_image_vs_oracle(zb):
    import torch
    R = Cc = 8192
    gen = torch.Generator(device="cuda").manual_seed(2)
    x = torch.rand(R, Cc, 4, device="cuda", dtype=torch.float32, generator=gen)
    src = zb.Image.from_tensor(x)
    taps = zb.gaussian_taps(2.25)
    assert taps.size == 15
    got = src.gaussian_blur(2.25).tensor()
    assert zb.lib().zb_last_kernel().decode() == "fused_sep_rgbaf32"
    zb.lib().zb_set_exact_f32(1)
    try:
        got_exact = src.gaussian_blur(2.25).tensor()
    finally:
        zb.lib().zb_set_exact_f32(0)
    xh = x.cpu().numpy()
    # the reference convolves float pixels one Image(f32) plane at a time (convolution.zig:322-436)
    worst = 0.0
    for ch in range(4):
        plane = np.ascontiguousarray(xh[..., ch])
        want = zo.conv_separable(plane, taps, taps, "mirror")
        g = got[..., ch].cpu().numpy()
        worst = max(worst, rel_err(g, want))
        assert np.array_equal(got_exact[..., ch].cpu().numpy(), want), f"exact build differs from the oracle in channel {ch}"
    assert worst <= 1e-5, worst


def test_c3_whole_image_vs_oracle(zb):
    import torch
    gen = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randint(0, 256, (16384, 16384, 3), device="cuda", dtype=torch.uint8, generator=gen)
    big = zb.Image.from_tensor(x)
    got = big.resize(zb.Image.init(4096, 4096, big.pixfmt), zb.Interpolation.BICUBIC).tensor().cpu().numpy()
    want = zo.resize(x.cpu().numpy(), (4096, 4096), "bicubic")
    assert np.array_equal(got, want)


def test_c4_batch_128_frames_vs_oracle(zb):
    import torch
    n, R, Cc = 128, 1080, 1920
    gen = torch.Generator(device="cuda").manual_seed(4)
    x = torch.randint(0, 256, (n, R, Cc, 4), device="cuda", dtype=torch.uint8, generator=gen)
    angle = np.float32(np.pi / 4)
    a64 = np.float64(angle)
    cs = (np.float32(np.cos(a64)), np.float32(np.sin(a64)))
    orows, ocols = zb.Image.from_tensor(x[0]).rotate_bounds(angle)
    assert (orows, ocols) == (2122, 2122)
    y = torch.empty((n, orows, ocols, 4), device="cuda", dtype=torch.uint8)
    src0 = zb.Image.from_tensor(x[0])._zb()
    dst0 = zb.Image.from_tensor(y[0])._zb()
    zb._ffi.check(zb.lib().zb_rotate_into_batch(src0, R * Cc, dst0, orows * ocols, n, int(zb.PixFmt.RGBA8), C.c_float(angle), C.c_float(cs[0]),
                                                C.c_float(cs[1]), int(zb.Interpolation.BILINEAR), C.c_float(1 / 3), C.c_float(1 / 3),
                                                int(zb.BorderMode.ZERO), zb.image.current_stream()))
    torch.cuda.synchronize()
    xh, yh = x.cpu().numpy(), y.cpu().numpy()
    for i in range(n):
        want = zo.rotate(xh[i], angle, "bilinear", "zero", cos_sin=cs)
        assert np.array_equal(yh[i], want), f"frame {i}"


def test_c5_whole_image_vs_oracle_flip_count(zb):
    import torch
    from zignal_b200.fdm import FeatureDistributionMatching
    g = torch.Generator(device="cuda").manual_seed(5)
    src = torch.randint(0, 256, (4096, 4096, 3), device="cuda", dtype=torch.uint8, generator=g)
    base = torch.randint(0, 256, (4096, 4096, 3), device="cuda", dtype=torch.uint8, generator=g).to(torch.float32)
    mix = torch.tensor([[0.5, 0.2, 0.0], [0.1, 0.4, 0.1], [0.0, 0.2, 0.45]], device="cuda")
    tgt = (base @ mix.T * 0.6 + 50.0).clamp(0, 255).to(torch.uint8)
    sh, th = src.cpu().numpy(), tgt.cpu().numpy()
    si, ti = zb.Image.from_tensor(src), zb.Image.from_tensor(tgt)
    f = FeatureDistributionMatching(si.pixfmt)
    f.match(si, ti)
    f.status()
    got = si.tensor().cpu().numpy()
    f.deinit()
    want = zo.fdm_match(sh, th)
    diff = got.astype(np.int16) - want.astype(np.int16)
    flips = int(np.count_nonzero(diff))
    print(f"C5 4096^2: {flips} of {diff.size} values differ from the Welford-statistics oracle (all by one count)")
    # exact integer moments vs the reference's sequential Welford stream (stats.zig:261-320): the 3x3 map differs in the last
    # bits of its f64 entries, so a value within ~1e-12 of a rounding boundary k + 1/2 may land on the other side
    assert np.abs(diff).max() <= 1
    assert flips <= diff.size // 100000, f"{flips} flips"


def test_image_mirror_rejects_mismatched_out(zb):
    """the C ABI takes one pixfmt for both buffers; the Python mirror must refuse an `out` of another pixel type
    (Zig's Image(T) typing makes that a compile error) instead of writing 16 B/px into a 1 B/px allocation."""
    a = zb.Image.from_numpy(np.zeros((32, 32, 4), np.float32))
    wrong = zb.Image.from_numpy(np.zeros((32, 32), np.uint8))
    for call in (lambda: a.gaussian_blur(1.0, out=wrong), lambda: a.box_blur(1, out=wrong), lambda: a.resize(wrong),
                 lambda: a.rotate_into(wrong, 0.3), lambda: a.copy(wrong), lambda: a.psnr(wrong)):
        with pytest.raises(TypeError):
            call()
