"""FeatureDistributionMatching and GEMM on the GPU vs the oracle."""
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


def _fdm(zb, src, tgt):
    from zignal_b200.fdm import FeatureDistributionMatching
    f = FeatureDistributionMatching(zb.image.pixfmt_of_array(src))
    s = zb.Image.from_numpy(src)
    t = zb.Image.from_numpy(tgt)
    f.match(s, t)
    f.status()          # the 3x3 solve runs on the device; this waits and raises SvdFailed if it did not converge
    out = s.to_numpy()
    f.deinit()
    return out


def _moments_numpy(img, as_luma=False):
    x = img.reshape(-1, img.shape[2] if img.ndim == 3 else 1).astype(np.int64)
    if x.shape[1] == 1:
        r = g = b = x[:, 0]
        ng = 0
    else:
        r, g, b = x[:, 0], x[:, 1], x[:, 2]
        ng = int(np.count_nonzero((r != g) | (g != b)))
        if as_luma:
            y = np.clip((13933 * r + 46871 * g + 4732 * b + 32768) >> 16, 0, 255)
            r = g = b = y
    return [len(r), r.sum(), g.sum(), b.sum(), (r * r).sum(), (r * g).sum(), (r * b).sum(), (g * g).sum(), (g * b).sum(), (b * b).sum(), ng]


@pytest.mark.parametrize("shape", [(37, 41), (64, 50, 3), (33, 47, 4), (512, 700, 3)])
def test_moments_exact(zb, shape):
    rng = np.random.default_rng(shape[0])
    img = rand_image(rng, shape, np.uint8)
    dev = zb.Image.from_numpy(img)
    for luma in ([0] if len(shape) == 2 else [0, 1]):
        sums = (C.c_uint64 * 11)()
        a = dev._zb()
        zb._ffi.check(zb.lib().zb_fdm_moments(a, int(dev.pixfmt), luma, sums, None))
        assert list(sums) == _moments_numpy(img, bool(luma))


def test_fdm_kats(zb):
    src = np.arange(100, dtype=np.uint8).reshape(100, 1)  # fdm.zig:429-464
    tgt = (100 + np.arange(100)).astype(np.uint8).reshape(100, 1)
    out = _fdm(zb, src, tgt)
    assert out.astype(np.float64).mean() == 149.5
    assert np.array_equal(out, zo.fdm_match(src, tgt))
    i = np.arange(2500)  # fdm.zig:325-427
    x, y = i % 50, i // 50
    s = np.stack([100 + x % 20, 150 + y % 15, 80 + (x + y) % 25], -1).astype(np.uint8).reshape(50, 50, 3)
    t = np.stack([50 + x % 30, 70 + y % 20, 90 + (x + y) % 35], -1).astype(np.uint8).reshape(50, 50, 3)
    out = _fdm(zb, s, t)
    o, tt = out.reshape(-1, 3).astype(np.float64), t.reshape(-1, 3).astype(np.float64)
    assert np.all(np.abs(o.mean(0) - tt.mean(0)) <= 2.0) and np.all(np.abs(o.var(0) - tt.var(0)) <= 1.0)


@pytest.mark.parametrize("ch", [1, 3, 4])
def test_fdm_matches_oracle(zb, ch):
    """u8 outputs must equal the oracle's; the statistics are exact integer moments instead of the reference's
    sequential Welford stream, so a pixel whose mapped value lands within ~1e-13 of a rounding boundary may flip
    by one count -- the test counts such flips and allows a vanishing fraction."""
    rng = np.random.default_rng(ch)
    shape = (240, 311) + ((ch,) if ch > 1 else ())
    src = rand_image(rng, shape, np.uint8)
    tgt = (rand_image(rng, shape, np.uint8) // 2 + 40).astype(np.uint8)
    if ch > 1:
        tgt[..., 1] = (tgt[..., 1].astype(np.int32) * 3 // 4 + tgt[..., 0] // 4).astype(np.uint8)
    got = _fdm(zb, src, tgt)
    want = zo.fdm_match(src, tgt)
    diff = got.astype(np.int32) - want.astype(np.int32)
    flips = int(np.count_nonzero(diff))
    assert np.abs(diff).max() <= 1 and flips <= max(2, src.size // 100000), f"{flips} flips"
    if ch == 4:
        assert np.array_equal(got[..., 3], src[..., 3])  # colour branch leaves alpha untouched (fdm.zig:268-270)


def test_fdm_gray_target_branch_and_reuse(zb):
    from zignal_b200.fdm import FeatureDistributionMatching
    rng = np.random.default_rng(9)
    src = rand_image(rng, (60, 70, 4), np.uint8)
    g = rng.integers(60, 200, (60, 70), dtype=np.uint8)
    tgt = np.repeat(g[..., None], 4, axis=2)
    got = _fdm(zb, src, tgt)
    want = zo.fdm_match(src, tgt)
    assert np.abs(got.astype(int) - want.astype(int)).max() <= 1
    assert np.all(got[..., 3] == 0)  # appendix A.14: the gray branch rewrites alpha to the struct default 0
    # batch reuse of the target statistics (fdm.zig:466-529) and error order (fdm.zig:583-604)
    f = FeatureDistributionMatching(zb.PixFmt.RGB8)
    with pytest.raises(zb.ZignalError) as e:
        f.update()
    assert e.value.name == "NoTargetSet"
    t3 = rand_image(rng, (20, 20, 3), np.uint8)
    f.set_target(zb.Image.from_numpy(t3))
    with pytest.raises(zb.ZignalError) as e:
        f.update()
    assert e.value.name == "NoSourceSet"
    for seed in (1, 2):
        s3 = rand_image(np.random.default_rng(seed), (20, 20, 3), np.uint8)
        d = zb.Image.from_numpy(s3)
        f.set_source(d)
        f.update()
        assert np.abs(d.to_numpy().astype(int) - zo.fdm_match(s3, t3).astype(int)).max() <= 1
    f.deinit()
    assert np.abs(zb.fdm.host_match(src[..., :3].copy(), tgt[..., :3].copy()).astype(int) - zo.fdm_match(src[..., :3].copy(), tgt[..., :3].copy()).astype(int)).max() <= 1


@pytest.mark.parametrize("ch", [1, 3, 4])
def test_fdm_tail_pixels_and_unaligned_row_views(zb, ch):
    """The kernels work on 4-pixel word groups: a pixel count that is not a multiple of 4 (byte-wise tail) and a full-width row
    view that starts on an odd byte (Rgb: 3 * 311 bytes per row) must give what the oracle gives for that view."""
    from zignal_b200.fdm import FeatureDistributionMatching
    rng = np.random.default_rng(40 + ch)
    shape = (63, 311) + ((ch,) if ch > 1 else ())
    base = rand_image(rng, shape, np.uint8)
    tgt = (rand_image(rng, shape, np.uint8) // 3 + 90).astype(np.uint8)
    for r0, r1 in ((0, 63), (1, 62), (3, 8)):
        dev = zb.Image.from_numpy(base.copy())
        view = dev.view(zb.Rectangle(0, r0, 311, r1))
        f = FeatureDistributionMatching(dev.pixfmt)
        f.match(view, zb.Image.from_numpy(tgt))
        f.status()
        got = dev.to_numpy()
        f.deinit()
        want = zo.fdm_match(np.ascontiguousarray(base[r0:r1]), tgt)
        assert np.abs(got[r0:r1].astype(int) - want.astype(int)).max() <= 1, (r0, r1)
        assert np.array_equal(got[:r0], base[:r0]) and np.array_equal(got[r1:], base[r1:])   # rows outside the view untouched


@pytest.mark.parametrize("dtype,tol", [(np.float32, 2e-6), (np.float64, 1e-12)])
@pytest.mark.parametrize("ta", [False, True])
@pytest.mark.parametrize("tb", [False, True])
def test_gemm(zb, dtype, tol, ta, tb):
    from zignal_b200.matrix import gemm
    rng = np.random.default_rng(int(ta) * 2 + int(tb))
    for m, k, n in [(2, 3, 2), (13, 11, 9), (70, 300, 65), (256, 5000, 256), (1, 1, 1)]:
        a = rng.standard_normal((k, m) if ta else (m, k)).astype(dtype)
        b = rng.standard_normal((n, k) if tb else (k, n)).astype(dtype)
        c = rng.standard_normal((m, n)).astype(dtype)
        got = gemm(a, b, ta, tb, 0.5, 2.0, c)
        ref = 0.5 * ((a.T if ta else a).astype(np.float64) @ (b.T if tb else b).astype(np.float64)) + 2.0 * c.astype(np.float64)
        scale = np.abs(ref).max() + 1e-30
        assert np.abs(got - ref).max() / scale <= tol, (m, k, n)
        got0 = gemm(a, b, ta, tb, 1.0, 0.0, None)
        ref0 = (a.T if ta else a).astype(np.float64) @ (b.T if tb else b).astype(np.float64)
        assert np.abs(got0 - ref0).max() / (np.abs(ref0).max() + 1e-30) <= tol
    a9 = np.repeat(np.arange(1, 10, dtype=dtype)[:, None], 9, axis=1)  # test_ops_gemm.zig:246-299
    r1 = gemm(a9, a9)
    assert r1[0, 0] == 45 and r1[1, 0] == 90 and r1[8, 8] == 405
    assert gemm(a9, a9, True, False)[0, 0] == 285 and gemm(a9, a9, False, True)[8, 8] == 729
    with pytest.raises(zb.ZignalError):
        gemm(np.zeros((3, 4), dtype), np.zeros((5, 2), dtype))


@pytest.mark.parametrize("n,dim,positive", [(4096, 128, False), (8200, 256, False), (8200, 256, True), (5003, 384, False)])
def test_gemm_xtx_tensorcore(zb, n, dim, positive):
    """Pca.fit's covariance contraction (pca.zig:338): the tcgen05 3xTF32 kernel against the oracle's reference-order f32 GEMM
    and against the exact f64 product.  Tolerance 1e-5 of max|C| (north_star's f32 bar); the kernel sits near 1e-6."""
    import torch
    from zignal_b200 import matrix
    rng = np.random.default_rng(n + dim)
    x = (rng.random((n, dim)) * 255.0 if positive else rng.standard_normal((n, dim))).astype(np.float32)
    x[:, 3] *= 7.0
    xd = torch.from_numpy(x).cuda()
    got = matrix.gemm_device(xd, xd, True, False, 1.0, 0.0, None).cpu().numpy()
    assert zb.lib().zb_last_kernel().decode() == "gemm_xtx_tf32x3_tcgen05"
    exact = x.astype(np.float64).T @ x.astype(np.float64)
    scale = np.abs(exact).max()
    assert np.abs(got - exact).max() / scale <= 2e-6
    assert np.array_equal(got, got.T)                      # mirrored tiles: exactly symmetric
    want = zo.gemm(x, x, True, False)                      # the reference's own accumulation order, f32
    assert np.abs(got - want).max() / scale <= 1e-5
    # alpha / beta epilogue and the CUDA-core path agree with it
    c0 = rng.standard_normal((dim, dim)).astype(np.float32)
    got2 = matrix.gemm_device(xd, xd, True, False, 1.0 / (n - 1), 2.0, torch.from_numpy(c0).cuda()).cpu().numpy()
    ref2 = exact / (n - 1) + 2.0 * c0
    assert np.abs(got2 - ref2).max() / np.abs(ref2).max() <= 2e-6
    zb.lib().zb_set_force_generic(1)
    try:
        gen = matrix.gemm_device(xd, xd, True, False, 1.0, 0.0, None).cpu().numpy()
        assert zb.lib().zb_last_kernel().decode() != "gemm_xtx_tf32x3_tcgen05"
    finally:
        zb.lib().zb_set_force_generic(0)
    assert np.abs(gen - got).max() / scale <= 2e-6


def test_pca_fit_and_transform(zb):
    from zignal_b200.pca import Pca
    rng = np.random.default_rng(3)
    for dtype, n, dim in [(np.float64, 200, 6), (np.float32, 500, 12), (np.float64, 5, 10)]:
        x = (rng.standard_normal((n, dim)) @ rng.standard_normal((dim, dim))).astype(dtype)
        p = Pca(dtype)
        p.fit(x)
        mean, comps, eig = zo.pca_fit(x)
        tol = 1e-9 if dtype == np.float64 else 5e-4
        assert np.allclose(p.mean, mean, atol=tol)
        assert np.allclose(p.eigenvalues, eig, rtol=tol, atol=tol * max(1.0, float(eig[0])))
        k = p.num_components
        # compare subspaces up to column sign (SVD contract, appendix A.15)
        for i in range(min(k, 3)):
            d = abs(float(p.components[:, i] @ comps[:, i]))
            assert abs(d - 1.0) < (1e-6 if dtype == np.float64 else 1e-2)
        t = p.transform(x)
        recon = t @ p.components.T + p.mean
        if k == min(n - 1, dim) and n > dim:
            assert np.allclose(recon, x, atol=1e-8 if dtype == np.float64 else 1e-3)
    data = np.array([[1, 0, 0], [3, 0, 0]], np.float64)  # pca.zig:519-557
    p = Pca(np.float64)
    p.fit(data, 1)
    assert abs(p.eigenvalues[0] - 2.0) < 1e-9 and abs(abs(p.components[0, 0]) - 1.0) < 1e-9


def test_config5_full_size_properties(zb):
    """BASELINE config 5 at full size (4096x4096 Rgb u8), through size-independent properties: the integer moments equal a
    torch int64 reduction exactly, and after fdm.match the source's channel means / covariance land on the target's (the map
    is affine, so only the 8-bit quantisation and the [0, 255] clamp separate them)."""
    import torch
    from zignal_b200.fdm import FeatureDistributionMatching
    g = torch.Generator(device="cuda").manual_seed(5)
    src = torch.randint(0, 256, (4096, 4096, 3), device="cuda", dtype=torch.uint8, generator=g)
    base = torch.randint(0, 256, (4096, 4096, 3), device="cuda", dtype=torch.uint8, generator=g).to(torch.float32)
    mix = torch.tensor([[0.5, 0.2, 0.0], [0.1, 0.4, 0.1], [0.0, 0.2, 0.45]], device="cuda")
    tgt = (base @ mix.T * 0.6 + 50.0).clamp(0, 255).to(torch.uint8)          # a different, full-rank colour distribution
    si, ti = zb.Image.from_tensor(src.clone()), zb.Image.from_tensor(tgt)
    m = FeatureDistributionMatching.moments(si)
    x = src.reshape(-1, 3).to(torch.int64)
    assert int(m[0]) == x.shape[0] and [int(v) for v in m[1:4]] == x.sum(0).tolist()
    assert int(m[5]) == int((x[:, 0] * x[:, 1]).sum()) and int(m[9]) == int((x[:, 2] ** 2).sum())
    f = FeatureDistributionMatching(si.pixfmt)
    f.match(si, ti)
    f.status()
    out = si.tensor().reshape(-1, 3).to(torch.float64)
    want = tgt.reshape(-1, 3).to(torch.float64)
    assert (out.mean(0) - want.mean(0)).abs().max().item() < 0.5
    assert (torch.cov(out.T) - torch.cov(want.T)).abs().max().item() < 0.02 * torch.cov(want.T).abs().max().item()
    f.deinit()


def test_pca_gram_full_size(zb):
    """The PCA contraction at BASELINE size (n = 1,048,576 x dim 256, tcgen05 3xTF32): against the f64 product, 5e-6 of max|C|."""
    import torch
    from zignal_b200 import matrix
    g = torch.Generator(device="cuda").manual_seed(7)
    X = torch.randn(1048576, 256, device="cuda", generator=g)
    C_ = matrix.gemm_device(X, X, True, False, 1.0 / (X.shape[0] - 1), 0.0, None)
    assert zb.lib().zb_last_kernel().decode() == "gemm_xtx_tf32x3_tcgen05"
    ref = torch.zeros(256, 256, dtype=torch.float64, device="cuda")
    for i in range(0, X.shape[0], 131072):                                    # f64 reference in slices (keeps memory modest)
        blk = X[i:i + 131072].double()
        ref += blk.T @ blk
    ref /= (X.shape[0] - 1)
    assert ((C_.double() - ref).abs().max() / ref.abs().max()).item() <= 5e-6
    assert torch.equal(C_, C_.T)
