"""Size-independent properties of the oracle's restatements of the SURVEY 8(f) additions (CPU only): facts that follow from the reference's
definitions and hold for every input, checked on seeded random images.  They guard the restatement against transcription slips that the
handful of reference KATs would not see."""
import numpy as np
import pytest

import oracle_lib as zo
from gpu_utils import rand_image


@pytest.mark.parametrize("shape", [(23, 31), (17, 19, 3), (14, 16, 4)])
def test_order_statistics_are_ordered(shape):
    """Percentiles are monotone in p (so min <= median <= max), midpoint = ceil((min + max) / 2), the untrimmed mean lies between min and
    max, and trimming down to one sample leaves the median."""
    rng = np.random.default_rng(shape[0])
    img = rand_image(rng, shape, np.uint8)
    for radius, border in [(1, "mirror"), (2, "zero"), (3, "wrap")]:
        ps = [zo.order_blur(img, radius, "percentile", p, border).astype(int) for p in (0.0, 0.2, 0.5, 0.8, 1.0)]
        for lo, hi in zip(ps, ps[1:]):
            assert np.all(lo <= hi)
        mid = zo.order_blur(img, radius, "midpoint", 0.0, border).astype(int)
        assert np.array_equal(mid, (ps[0] + ps[-1] + 1) // 2)
        area = (2 * radius + 1) ** 2
        mean = zo.order_blur(img, radius, "alpha_trimmed", 0.0, border).astype(int)
        assert np.all(mean >= ps[0]) and np.all(mean <= ps[-1])
        trimmed = zo.order_blur(img, radius, "alpha_trimmed", 0.49, border).astype(int)                     # keeps area - 2 * floor(0.49 area) samples
        kept = area - 2 * min(int(np.floor(0.49 * area)), area // 2)
        if kept == 1:
            assert np.array_equal(trimmed, ps[2])                                                           # one sample left: the median


def test_canny_is_monotone_in_its_thresholds():
    """Raising the high threshold can only remove seeds, raising the low one can only remove candidates: the edge set shrinks (hysteresis is
    a closure over a fixed candidate graph).  All edges lie on NMS survivors, which do not depend on the thresholds."""
    rng = np.random.default_rng(5)
    yy, xx = np.mgrid[0:90, 0:120]
    img = np.clip(128 + 80 * np.sin(xx / 9.0) * np.cos(yy / 7.0) + rng.normal(0, 6, (90, 120)), 0, 255).astype(np.uint8)
    base = zo.canny(img, 1.2, 10.0, 30.0) > 0
    assert np.all((zo.canny(img, 1.2, 10.0, 60.0) > 0) <= base)
    assert np.all((zo.canny(img, 1.2, 20.0, 30.0) > 0) <= base)
    everything = zo.canny(img, 1.2, 0.0, 1e-30) > 0                         # every NMS survivor with magnitude >= ~0 is a seed
    assert np.all(base <= everything)
    assert not everything[0].any() and not everything[:, 0].any()


def test_blend_identities():
    """mode none returns the overlay; a transparent overlay returns the base; an opaque overlay under `normal` returns the overlay; darken and
    lighten of opaque pixels are the channel-wise min / max; difference and exclusion, multiply and screen are symmetric in opaque inputs."""
    rng = np.random.default_rng(8)
    for _ in range(200):
        b = tuple(int(v) for v in rng.integers(0, 256, 4))
        o = tuple(int(v) for v in rng.integers(0, 256, 4))
        assert zo.blend_rgba8(b, o, "none") == o
        assert zo.blend_rgba8(b, o[:3] + (0,), "multiply") == b
        assert zo.blend_rgba8(b, o[:3] + (255,), "normal") == o[:3] + (255,)
        bo, oo = b[:3] + (255,), o[:3] + (255,)
        assert zo.blend_rgba8(bo, oo, "darken")[:3] == tuple(min(x, y) for x, y in zip(bo[:3], oo[:3]))
        assert zo.blend_rgba8(bo, oo, "lighten")[:3] == tuple(max(x, y) for x, y in zip(bo[:3], oo[:3]))
        for mode in ("difference", "exclusion", "multiply", "screen"):
            assert zo.blend_rgba8(bo, oo, mode) == zo.blend_rgba8(oo, bo, mode), mode
        a = zo.blend_rgba8(b, o, "overlay")[3]
        assert a >= max(b[3], o[3]) - 1 or b[3] == 0 or o[3] == 0           # "over" never lowers coverage


def test_convert_round_trips_and_consistency():
    rng = np.random.default_rng(12)
    rgba = rng.integers(0, 256, (20, 30, 4), dtype=np.uint8)
    rgb = np.ascontiguousarray(rgba[..., :3])
    assert np.array_equal(zo.convert(zo.convert(rgba, 4), 3), rgba)          # u8 -> f32 -> u8
    assert np.array_equal(zo.convert(zo.convert(rgb, 3), 2), rgb)            # Rgb -> Rgba -> Rgb
    assert np.array_equal(zo.convert(rgba, 0), zo.convert(rgb, 0))           # alpha does not enter the luma
    gray = zo.convert(rgb, 0)
    assert np.array_equal(zo.convert(zo.convert(gray, 2), 0), gray)          # gray -> Rgb -> gray: the luma weights sum to 65536
    assert np.array_equal(zo.convert(gray, 1), gray.astype(np.float32) / np.float32(255))
    f = rng.random((20, 30)).astype(np.float32)
    assert np.array_equal(zo.convert(f, 0), np.floor(f.astype(np.float64) * 255.0 + 0.5).astype(np.uint8))
    assert np.array_equal(zo.convert(f, 4)[..., 3], np.ones((20, 30), np.float32))


def test_metrics_identities():
    rng = np.random.default_rng(21)
    a = rng.integers(0, 256, (24, 28, 3), dtype=np.uint8)
    b = rng.integers(0, 256, (24, 28, 3), dtype=np.uint8)
    assert zo.psnr(a, b) == zo.psnr(b, a) and zo.mean_pixel_error(a, b) == zo.mean_pixel_error(b, a)
    assert abs(zo.ssim(a, b) - zo.ssim(b, a)) < 1e-15 and zo.ssim(a, a) == 1.0 and zo.psnr(a, a) == np.inf
    mse = np.mean((a.astype(np.float64) - b) ** 2)
    assert abs(zo.psnr(a, b) - (20 * np.log10(255.0) - 10 * np.log10(mse))) < 1e-12
    assert abs(zo.mean_pixel_error(a, b) - np.mean(np.abs(a.astype(np.float64) - b)) / 255.0) < 1e-15
    assert -1.0 <= zo.ssim(a, b) <= 1.0


def test_motion_blur_preserves_constants_and_mass():
    """A constant image is a fixed point of every motion blur (each output is a mean of samples of the constant); a horizontal blur of odd
    length keeps row sums away from the border."""
    const = np.full((30, 40, 3), 93, np.uint8)
    assert np.array_equal(zo.motion_blur_linear(const, 0.6, 7), const)
    assert np.array_equal(zo.motion_blur_radial(const, 0.3, 0.6, 0.8), const)
    assert np.array_equal(zo.motion_blur_radial(const, 0.3, 0.6, 0.8, spin=True), const)
    rng = np.random.default_rng(3)
    f = rng.random((12, 40)).astype(np.float32)
    h = zo.motion_blur_linear(f, 0.0, 5)
    want = np.stack([f[:, i:i + 36] for i in range(5)]).mean(axis=0)
    assert np.allclose(h[:, 2:38], want, atol=1e-6)


@pytest.mark.parametrize("n", [2, 5, 9])
def test_eigh_invariants(n):
    rng = np.random.default_rng(n)
    m = rng.normal(size=(n, n))
    a = (m + m.T) * 0.5
    vals, vecs = zo.eigh(a)
    assert np.allclose(vecs.T @ vecs, np.eye(n), atol=1e-12)
    assert abs(vals.sum() - np.trace(a)) < 1e-10 and np.all(np.diff(vals) >= 0)
    shifted, _ = zo.eigh(a + 3.0 * np.eye(n))
    assert np.allclose(shifted, vals + 3.0, atol=1e-10)
