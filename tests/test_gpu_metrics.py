"""Image.psnr / ssim / meanPixelError on the device (reference image/metrics.zig) against the oracle.  8-bit psnr and
meanPixelError are exact (integer sums); float formats and ssim differ from the oracle only in how the final f64 sum is associated,
checked at 1e-12 relative."""
import numpy as np
import pytest

import oracle_lib as zo
from gpu_utils import rand_image

pytestmark = pytest.mark.gpu

REL = 1e-12


@pytest.fixture(scope="module")
def zb():
    import torch
    assert torch.cuda.is_available()
    import zignal_b200 as zb
    return zb


def _pair(rng, shape, dtype):
    a = rand_image(rng, shape, dtype)
    if dtype == np.uint8:
        b = np.clip(a.astype(np.int32) + rng.integers(-25, 26, a.shape), 0, 255).astype(np.uint8)
    else:
        b = np.clip(a + rng.normal(0, 0.05, a.shape).astype(np.float32), 0, 1).astype(np.float32)
    return a, b


CASES = [((37, 45), np.uint8), ((40, 33, 3), np.uint8), ((29, 31, 4), np.uint8), ((35, 36), np.float32), ((23, 27, 4), np.float32),
         ((11, 11), np.uint8), ((300, 419, 4), np.uint8), ((257, 300), np.float32)]


@pytest.mark.parametrize("shape,dtype", CASES)
def test_metrics_match_oracle(zb, shape, dtype):
    rng = np.random.default_rng(shape[0] * 3 + shape[1])
    a, b = _pair(rng, shape, dtype)
    da, db = zb.Image.from_numpy(a), zb.Image.from_numpy(b)
    want_psnr, want_mpe, want_ssim = zo.psnr(a, b), zo.mean_pixel_error(a, b), zo.ssim(a, b)
    got_psnr, got_mpe, got_ssim = da.psnr(db), da.mean_pixel_error(db), da.ssim(db)
    if dtype == np.uint8:
        assert got_psnr == want_psnr and got_mpe == want_mpe
    else:
        assert abs(got_psnr - want_psnr) <= REL * abs(want_psnr) and abs(got_mpe - want_mpe) <= REL * abs(want_mpe)
    assert abs(got_ssim - want_ssim) <= REL * abs(want_ssim), (got_ssim, want_ssim)
    assert da.psnr(da) == np.inf and da.ssim(da) == 1.0 and da.mean_pixel_error(da) == 0.0
    assert da.ssim(db) == got_ssim                                         # reproducible: fixed block order, host-side final sum


def test_metrics_reference_cases(zb):
    """image/tests/psnr.zig:13-130 and metrics.zig:253-293 through the device path."""
    def img(x):
        return zb.Image.from_numpy(np.ascontiguousarray(x))
    a = np.array([[100, 150], [200, 250]], np.uint8)
    b = np.array([[110, 140], [205, 245]], np.uint8)
    assert abs(img(a).psnr(img(b)) - 30.171) < 0.01
    rgb1 = np.tile(np.array([100, 150, 200], np.uint8), (2, 2, 1))
    rgb2 = np.tile(np.array([110, 140, 205], np.uint8), (2, 2, 1))
    assert abs(img(rgb1).psnr(img(rgb2)) - 29.38) < 0.01
    ra = np.array([[[255, 0, 0, 255], [0, 255, 0, 255]]], np.uint8)
    rb = np.array([[[250, 5, 0, 255], [0, 250, 5, 255]]], np.uint8)
    assert abs(img(ra).psnr(img(rb)) - 37.16) < 0.01
    f1 = np.array([[0.5, 0.7], [0.3, 0.9]], np.float32)
    f2 = np.array([[0.4, 0.8], [0.2, 1.0]], np.float32)
    assert abs(img(f1).psnr(img(f2)) - 20.0) < 0.01
    one = np.array([[[255, 0, 0]]], np.uint8)
    assert abs(img(one).mean_pixel_error(img(np.zeros_like(one))) - 1.0 / 3.0) < 1e-9
    chk = np.zeros((12, 12, 3), np.uint8)
    rr, cc = np.mgrid[0:12, 0:12]
    chk[(rr + cc) % 2 == 0] = (255, 0, 0)
    chk[(rr + cc) % 2 == 1] = (0, 255, 0)
    assert img(chk).ssim(img(np.zeros_like(chk))) < 0.99
    with pytest.raises(zb.ZignalError) as ei:
        img(a).psnr(img(np.zeros((2, 3), np.uint8)))
    assert ei.value.name == "DimensionMismatch"
    with pytest.raises(zb.ZignalError) as ei:
        img(np.zeros((10, 30), np.uint8)).ssim(img(np.zeros((10, 30), np.uint8)))
    assert ei.value.name == "ImageTooSmall"


def test_metrics_on_views(zb):
    """Strided views (Image.view, image.zig:426-430) of a larger device image."""
    rng = np.random.default_rng(9)
    a, b = _pair(rng, (90, 120, 4), np.uint8)
    da, db = zb.Image.from_numpy(a), zb.Image.from_numpy(b)
    va, vb = da.view(zb.Rectangle(9, 7, 86, 68)), db.view(zb.Rectangle(9, 7, 86, 68))
    ca, cb = np.ascontiguousarray(a[7:68, 9:86]), np.ascontiguousarray(b[7:68, 9:86])
    assert va.psnr(vb) == zo.psnr(ca, cb)
    assert abs(va.ssim(vb) - zo.ssim(ca, cb)) <= REL
