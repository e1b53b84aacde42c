"""Image.rotate / rotateInto / warp on the GPU vs the oracle (bit-exact for u8 formats given identical
cos/sin inputs; f32 expected bit-identical as well)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as zo
from gpu_utils import BORDERS, METHODS, border_enum, method_enum, rand_image, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def zb():
    import torch
    assert torch.cuda.is_available()
    import zignal_b200 as zb
    return zb


FORMATS = [((23, 31), np.uint8), ((20, 27, 3), np.uint8), ((25, 33, 4), np.uint8), ((19, 22), np.float32), ((17, 21, 4), np.float32)]


def _cs(angle):
    a = np.float32(angle)
    return np.float32(np.cos(np.float64(a))), np.float32(np.sin(np.float64(a)))


@pytest.mark.parametrize("shape,dtype", FORMATS)
@pytest.mark.parametrize("method", METHODS)
def test_rotate_general(zb, shape, dtype, method):
    rng = np.random.default_rng(METHODS.index(method) + shape[0])
    img = rand_image(rng, shape, dtype)
    dev = zb.Image.from_numpy(img)
    for angle in (np.pi / 4, 0.3, -1.1, 2.5):
        for border in BORDERS:
            cs = _cs(angle)
            got = dev.rotate(np.float32(angle), method_enum(zb, method), border_enum(zb, border), cos_sin=cs).to_numpy()
            want = zo.rotate(img, np.float32(angle), method, border, cos_sin=cs)
            assert got.shape == want.shape
            if dtype == np.uint8:
                assert np.array_equal(got, want), (angle, border)
            else:
                assert rel_err(got, want) <= 1e-5 and np.array_equal(got, want), (angle, border)


@pytest.mark.parametrize("shape,dtype", FORMATS)
def test_rotate_orthogonal_fast_paths(zb, shape, dtype):
    rng = np.random.default_rng(5)
    img = rand_image(rng, shape, dtype)
    dev = zb.Image.from_numpy(img)
    for k, angle in enumerate([0.0, np.pi / 2, np.pi, 3 * np.pi / 2, 2 * np.pi, -np.pi / 2]):
        got = dev.rotate(np.float32(angle)).to_numpy()
        want = zo.rotate(img, np.float32(angle))
        assert np.array_equal(got, want), angle
    # rotateInto a larger / smaller destination: centred content, zeroed margins only when an offset exists
    for out_shape in [(shape[0] + 6, shape[1] + 4), (shape[0] - 3, shape[1] + 5), (shape[1] + 1, shape[0])]:
        for angle in (0.0, np.pi / 2, np.pi):
            fill = rand_image(rng, out_shape + tuple(shape[2:]), dtype)
            out = zb.Image.from_numpy(fill)
            dev.rotate_into(out, np.float32(angle))
            want = zo.rotate_into(img, fill.copy(), np.float32(angle))
            assert np.array_equal(out.to_numpy(), want), (out_shape, angle)


def test_rotate_kats(zb):
    img = np.arange(1, 13, dtype=np.uint8).reshape(3, 4)  # tests/transforms.zig:160-209
    dev = zb.Image.from_numpy(img)
    for angle, k in [(0.0, 0), (np.pi / 2, 1), (np.pi, 2), (3 * np.pi / 2, 3)]:
        assert np.array_equal(dev.rotate(np.float32(angle), zb.Interpolation.BILINEAR, zb.BorderMode.MIRROR).to_numpy(), np.rot90(img, k))
    chk = (np.indices((10, 10)).sum(0) % 2 == 0).astype(np.uint8) * 255  # :211-229
    out = zb.Image.from_numpy(chk).rotate(np.float32(np.pi / 4), zb.Interpolation.BILINEAR, zb.BorderMode.MIRROR)
    assert out.rows > 10 and out.cols > 10


def test_rotate_batch_and_config4_shape(zb):
    """BASELINE config 4 (batch of 1920x1080 RGBA, 45 degrees, bilinear, zero border) on a small batch: every image of the
    batch must equal the oracle's single-image result; output shape 2122x2122."""
    import torch
    n, R, Cc = 3, 1080, 1920
    gen = torch.Generator(device="cuda").manual_seed(4)
    x = torch.randint(0, 256, (n, R, Cc, 4), device="cuda", dtype=torch.uint8, generator=gen)
    angle = np.float32(np.pi / 4)
    cs = _cs(angle)
    orows, ocols = zb.Image.from_tensor(x[0]).rotate_bounds(angle)
    assert (orows, ocols) == (2122, 2122)
    y = torch.empty((n, orows, ocols, 4), device="cuda", dtype=torch.uint8)
    src0 = zb.Image.from_tensor(x[0])._zb()
    dst0 = zb.Image.from_tensor(y[0])._zb()
    zb._ffi.check(zb.lib().zb_rotate_into_batch(src0, R * Cc, dst0, orows * ocols, n, int(zb.PixFmt.RGBA8), C.c_float(angle), C.c_float(cs[0]),
                                                C.c_float(cs[1]), int(zb.Interpolation.BILINEAR), C.c_float(1 / 3), C.c_float(1 / 3),
                                                int(zb.BorderMode.ZERO), zb.image.current_stream()))
    torch.cuda.synchronize()
    for i in (0, n - 1):
        want = zo.rotate(x[i].cpu().numpy(), angle, "bilinear", "zero", cos_sin=cs)
        assert np.array_equal(y[i].cpu().numpy(), want), i


def _last_kernel(zb):
    return zb.lib().zb_last_kernel().decode()


@pytest.mark.parametrize("shape", [(64, 100), (37, 52), (130, 260), (5, 8), (1, 4), (200, 64)])
def test_rotate_tile_kernel_vs_oracle(zb, shape):
    """Rgba(u8) / bilinear / .zero with a 16-byte aligned row pitch takes the shared-memory tile kernel (zb_rotate_tile.cu: TMA
    source tiles, zero fill as the border, magic-number floor + weight): bit-exact against the oracle for angles in every
    quadrant, near-orthogonal angles, tiny and odd shapes, destinations larger and smaller than rotateBounds."""
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    img = rng.integers(0, 256, shape + (4,), dtype=np.uint8)
    dev = zb.Image.from_numpy(img)
    for angle in (np.pi / 4, 0.3, -1.1, 2.5, 1e-3, np.pi / 2 + 2e-3, -np.pi / 4, 3.0, 4.0, 5.5, 0.7853):
        cs = _cs(angle)
        got = dev.rotate(np.float32(angle), zb.Interpolation.BILINEAR, zb.BorderMode.ZERO, cos_sin=cs).to_numpy()
        assert _last_kernel(zb) == "rotate_tile_rgba8"
        want = zo.rotate(img, np.float32(angle), "bilinear", "zero", cos_sin=cs)
        assert np.array_equal(got, want), angle
    for out_shape in ((shape[0] + 70, shape[1] + 3), (max(1, shape[0] // 2), max(1, shape[1] - 1)), (129, 65)):
        fill = rng.integers(0, 256, out_shape + (4,), dtype=np.uint8)
        out = zb.Image.from_numpy(fill)
        dev.rotate_into(out, np.float32(0.9))
        assert _last_kernel(zb) == "rotate_tile_rgba8"
        assert np.array_equal(out.to_numpy(), zo.rotate_into(img, fill.copy(), np.float32(0.9))), out_shape


def test_rotate_tile_kernel_views_and_gather_cross_check(zb):
    """A view with a row pitch larger than its width (still 16-byte aligned) stays on the tile kernel and reads nothing outside the
    view (TMA fills beyond the view's columns with zeros even though the parent has pixels there); a 1080p frame at several angles
    is bit-identical between the tile kernel and the gather kernel (zb_tune rotate.tile 0)."""
    import torch
    rng = np.random.default_rng(77)
    parent = rng.integers(1, 256, (90, 128, 4), dtype=np.uint8)
    pdev = zb.Image.from_numpy(parent)
    v = pdev.view(zb.Rectangle(8, 5, 8 + 60, 5 + 70)) if hasattr(zb, "Rectangle") else None
    if v is not None:
        sub = parent[5:75, 8:68].copy()
        got = v.rotate(np.float32(0.6), cos_sin=_cs(0.6)).to_numpy()
        assert _last_kernel(zb) == "rotate_tile_rgba8"
        assert np.array_equal(got, zo.rotate(sub, np.float32(0.6), "bilinear", "zero", cos_sin=_cs(0.6)))
    gen = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randint(0, 256, (1080, 1920, 4), device="cuda", dtype=torch.uint8, generator=gen)
    dev = zb.Image.from_tensor(x)
    for angle in (np.pi / 4, 0.1, -2.0, 1.5607):
        cs = _cs(angle)
        a = dev.rotate(np.float32(angle), cos_sin=cs)
        assert _last_kernel(zb) == "rotate_tile_rgba8"
        try:
            zb._ffi.check(zb.lib().zb_tune(b"rotate.tile", 0))
            b = dev.rotate(np.float32(angle), cos_sin=cs)
            assert _last_kernel(zb) == "rotate_gather"
        finally:
            zb._ffi.check(zb.lib().zb_tune(b"rotate.tile", 1))
        assert torch.equal(a._t, b._t), angle
    # destinations whose last tile row / column is 1..3 pixels
    for orows, ocols in ((1089, 1921), (1027, 2050)):
        out = zb.Image.init(orows, ocols, dev.pixfmt)
        dev.rotate_into(out, np.float32(0.4), cos_sin=_cs(0.4))
        assert _last_kernel(zb) == "rotate_tile_rgba8"
        zb._ffi.check(zb.lib().zb_tune(b"rotate.tile", 0))
        try:
            ref = zb.Image.init(orows, ocols, dev.pixfmt)
            dev.rotate_into(ref, np.float32(0.4), cos_sin=_cs(0.4))
        finally:
            zb._ffi.check(zb.lib().zb_tune(b"rotate.tile", 1))
        assert torch.equal(out._t, ref._t), (orows, ocols)


class _Xf:
    def __init__(self, kind, m):
        self.kind, self.m = kind, np.asarray(m, np.float32)

    def as_f32(self):
        return {"similarity": 0, "affine": 1, "projective": 2}[self.kind], self.m


@pytest.mark.parametrize("shape,dtype", FORMATS)
@pytest.mark.parametrize("method", METHODS)
def test_warp(zb, shape, dtype, method):
    rng = np.random.default_rng(METHODS.index(method) + 100)
    img = rand_image(rng, shape, dtype)
    dev = zb.Image.from_numpy(img)
    cases = [("affine", [1, 0, 0, 1, 0, 0]), ("affine", [0.9, -0.2, 0.25, 1.1, 1.5, -2.0]), ("similarity", [0.7, 0.7, -0.7, 0.7, 3, 4]),
             ("projective", [1.0, 0.05, 2, -0.03, 0.95, 1, 1e-3, -2e-3, 1]), ("projective", [2, 0, 0, 0, 2, 0, 0, 0, 2]),
             ("projective", [1, 0, 0, 0, 1, 0, 0, 0, 0])]
    for kind, m in cases:
        for out_shape in [shape[:2], (shape[0] + 5, shape[1] - 3)]:
            out = zb.Image.init(out_shape[0], out_shape[1], dev.pixfmt)
            got = dev.warp(out, _Xf(kind, m), method_enum(zb, method)).to_numpy()
            want = zo.warp(img, np.zeros(out_shape + tuple(shape[2:]), dtype), kind, m, method)
            if dtype == np.uint8:
                assert np.array_equal(got, want), (kind, m)
            else:
                assert rel_err(got, want) <= 1e-5 and np.array_equal(got, want), (kind, m)


def test_host_rotate_and_warp(zb):
    rng = np.random.default_rng(6)
    img = rand_image(rng, (30, 40, 4), np.uint8)
    angle = np.float32(0.6)
    got = zb.host_rotate(img, angle)
    # the library computes cos/sin with the host libm, exactly like the oracle's default
    import math
    cs = (np.float32(math.cos(float(angle))), np.float32(math.sin(float(angle))))
    want = zo.rotate(img, angle, "bilinear", "zero")
    assert got.shape == want.shape
    assert np.mean(got != want) < 1e-3  # cosf/sinf (library) vs numpy cos (oracle default) may differ by an ulp
    out = np.zeros((25, 35, 4), np.uint8)
    m = [0.9, -0.2, 0.25, 1.1, 1.5, -2.0]
    assert np.array_equal(zb.host_warp(img, out, _Xf("affine", m)), zo.warp(img, np.zeros_like(out), "affine", m, "bilinear"))


@pytest.mark.parametrize("fmt", ["u8", "rgb8", "rgba8", "f32"])
def test_extract_and_crop(zb, fmt):
    """Image.extract (transforms.zig:232-283): rotated / rescaled rectangles through every sampler, the copyRect fast path with
    every border mode, and crop (= extract nearest/zero into a rounded chip) -- bit-identical to the oracle."""
    from gpu_utils import METHODS, method_enum
    rng = np.random.default_rng(len(fmt))
    shape = {"u8": (40, 52), "rgb8": (40, 52, 3), "rgba8": (40, 52, 4), "f32": (40, 52)}[fmt]
    img = rand_image(rng, shape, np.float32 if fmt == "f32" else np.uint8)
    dev = zb.Image.from_numpy(img)
    tail = tuple(shape[2:])
    for rect, angle, out_shape in [((5.0, 4.0, 35.0, 28.0), 0.3, (20, 25)), ((-6.5, -3.0, 30.0, 50.0), -1.1, (17, 9)),
                                   ((10.0, 10.0, 20.0, 20.0), 0.0, (1, 1)), ((2.2, 3.7, 48.9, 30.1), 2.0, (33, 1))]:
        for method in METHODS:
            for border in ("zero", "mirror"):
                out = zb.Image.init(out_shape[0], out_shape[1], dev.pixfmt)
                got = dev.extract(out, rect, angle, method_enum(zb, method), getattr(zb.BorderMode, border.upper())).to_numpy()
                want = zo.extract(img, np.zeros(out_shape + tail, img.dtype), rect, angle, method, border)
                assert np.array_equal(got, want), (rect, angle, method, border)
    # copyRect fast path: rect of the output's own size, any border mode, partly or fully outside the image
    for rect in [(3.0, 2.0, 23.0, 17.0), (-4.4, -7.5, 15.6, 7.5), (45.0, 30.0, 65.0, 45.0), (100.0, 100.0, 120.0, 115.0)]:
        for border in ("zero", "replicate", "mirror", "wrap"):
            out = zb.Image.from_numpy(np.full((15, 20) + tail, 7, img.dtype))
            got = dev.extract(out, rect, 0.0, zb.Interpolation.BICUBIC, getattr(zb.BorderMode, border.upper())).to_numpy()
            assert zb.lib().zb_last_kernel().decode() == "extract_copy_rect"
            want = zo.extract(img, np.zeros((15, 20) + tail, img.dtype), rect, 0.0, "bicubic", border)
            assert np.array_equal(got, want), (rect, border)
    chip = dev.crop((4.4, -2.6, 30.5, 19.5))
    assert (chip.rows, chip.cols) == (22, 26)
    want = zo.extract(img, np.zeros((22, 26) + tail, img.dtype), (4.4, -2.6, 30.5, 19.5), 0.0, "nearest", "zero")
    assert np.array_equal(chip.to_numpy(), want)


@pytest.mark.parametrize("fmt", ["u8", "rgb8", "rgba8", "f32"])
def test_insert(zb, fmt):
    """Image.insert with blend_mode .none (transforms.zig:293-376): rotated / scaled rectangles, rectangles hanging over the image
    edge, the axis-aligned copy path, 1-pixel sources -- destination pixels outside the rectangle must stay untouched."""
    from gpu_utils import METHODS, method_enum
    rng = np.random.default_rng(7 + len(fmt))
    tail = {"u8": (), "rgb8": (3,), "rgba8": (4,), "f32": ()}[fmt]
    dtype = np.float32 if fmt == "f32" else np.uint8
    dest = rand_image(rng, (48, 60) + tail, dtype)
    for src_shape, rect, angle in [((20, 25), (5.0, 4.0, 35.0, 28.0), 0.4), ((9, 17), (-6.5, 30.0, 30.0, 58.0), -1.2),
                                   ((1, 1), (10.0, 10.0, 20.0, 20.0), 0.0), ((13, 1), (40.2, 3.7, 70.9, 30.1), 2.0),
                                   ((15, 20), (3.0, 2.0, 23.0, 17.0), 0.0), ((15, 20), (50.0, 40.0, 70.0, 55.0), 0.0)]:
        source = rand_image(rng, src_shape + tail, dtype)
        for method in METHODS:
            dev = zb.Image.from_numpy(dest.copy())
            got = dev.insert(zb.Image.from_numpy(source), rect, angle, method_enum(zb, method)).to_numpy()
            want = zo.insert(dest, source, rect, angle, method)
            assert np.array_equal(got, want), (src_shape, rect, angle, method)
    # extract -> insert round trip of the reference's own test (tests/transforms.zig:316-381): small average error in the centre
    if fmt == "u8":
        src = ((np.arange(64)[:, None] + np.arange(64)[None, :]) % 256).astype(np.uint8)
        rect, angle = (15.0, 15.0, 45.0, 45.0), float(np.float32(np.pi / 4))
        ext = zb.Image.from_numpy(src).extract(zb.Image.init(30, 30, zb.PixFmt.U8), rect, angle, zb.Interpolation.BILINEAR, zb.BorderMode.MIRROR)
        canvas = zb.Image.from_numpy(np.zeros((64, 64), np.uint8)).insert(ext, rect, angle, zb.Interpolation.BILINEAR).to_numpy()
        assert np.abs(src[21:39, 21:39].astype(int) - canvas[21:39, 21:39].astype(int)).mean() < 25


BLEND_MODES = ["none", "normal", "multiply", "screen", "overlay", "soft_light", "hard_light", "color_dodge", "color_burn", "darken", "lighten",
               "difference", "exclusion"]


def test_insert_blend_modes(zb):
    """Image.insert with a Blending mode (blending.zig:26-156 through image.zig:67-95 assignPixel): Rgba(u8) onto Rgba(u8), all thirteen
    modes, bit-exact; alpha values 0 / 255 and the channel extremes 0 / 255 are over-represented so every early return and every
    dodge / burn guard is hit.  Other pixel types ignore the mode."""
    rng = np.random.default_rng(2024)

    def biased(shape):
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        pick = rng.random(shape)
        img[pick < 0.12] = 0
        img[pick > 0.88] = 255
        img[rng.random(shape) < 0.1] = 128
        return img

    dest = biased((70, 90, 4))
    cases = [((40, 50), (10.0, 8.0, 60.0, 48.0), 0.0, "nearest"),          # axis-aligned copy path with blending
             ((23, 31), (5.5, 4.0, 80.0, 66.0), 0.35, "bilinear"),
             ((17, 12), (-8.0, 30.0, 40.0, 75.0), -1.1, "bicubic")]
    for src_shape, rect, angle, method in cases:
        source = biased(src_shape + (4,))
        for i, mode in enumerate(BLEND_MODES):
            dev = zb.Image.from_numpy(dest.copy())
            got = dev.insert(zb.Image.from_numpy(source), rect, angle, getattr(zb.Interpolation, method.upper()), blend=zb.Blending(i)).to_numpy()
            want = zo.insert(dest, source, rect, angle, method, blend=mode)
            assert np.array_equal(got, want), (src_shape, mode, int((got != want).sum()))
    # exhaustive over one channel pair at several alphas: 256 x 256 base / overlay values per mode through the copy path
    b, o = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    for alpha_b, alpha_o in [(255, 255), (255, 128), (77, 201), (1, 254)]:
        base = np.stack([b, b[::-1], o, np.full_like(b, alpha_b)], axis=-1)
        over = np.stack([o, o, b[::-1], np.full_like(b, alpha_o)], axis=-1)
        for i, mode in enumerate(BLEND_MODES[1:], start=1):
            got = zb.Image.from_numpy(base.copy()).insert(zb.Image.from_numpy(over), (0.0, 0.0, 256.0, 256.0), 0.0, zb.Interpolation.NEAREST,
                                                          blend=zb.Blending(i)).to_numpy()
            want = zo.insert(base, over, (0.0, 0.0, 256.0, 256.0), 0.0, "nearest", blend=mode)
            assert np.array_equal(got, want), (mode, alpha_b, alpha_o, int((got != want).sum()))
    rgb = biased((30, 30, 3))
    patch = biased((10, 10, 3))
    got = zb.Image.from_numpy(rgb.copy()).insert(zb.Image.from_numpy(patch), (5.0, 5.0, 15.0, 15.0), 0.0, zb.Interpolation.NEAREST,
                                                 blend=zb.Blending.MULTIPLY).to_numpy()
    assert np.array_equal(got, zo.insert(rgb, patch, (5.0, 5.0, 15.0, 15.0), 0.0, "nearest", blend="multiply"))
    with pytest.raises(zb.ZignalError):
        zb.lib().zb_insert_blend  # symbol exists
        from zignal_b200._ffi import check
        d = zb.Image.from_numpy(rgb.copy())
        import ctypes as C
        check(zb.lib().zb_insert_blend(d._zb(), d._zb(), int(d.pixfmt), C.c_float(0), C.c_float(0), C.c_float(1), C.c_float(1), C.c_float(0),
                                       C.c_float(1), C.c_float(0), 0, C.c_float(0), C.c_float(0), 13, None))


def _oracle_insert_mixed(dest, source, rect, angle, method, blend):
    """Image.insert with a source of another pixel type, composed from oracle pieces exactly as the reference composes it
    (transforms.zig:293-376 + image.zig:67-95 assignPixel): the SAMPLE every destination pixel receives is what the same-type insert
    of the source's type writes (two passes over differently filled canvases tell written pixels from untouched ones), then
    assignPixel: an Rgba(u8) sample with a blend mode composites through Rgba(u8), anything else is convertColor(DestType, sample)."""
    sfmt, dfmt = zo.pixfmt_of(source), zo.pixfmt_of(dest)
    shape = dest.shape[:2] + tuple(source.shape[2:])
    lo = np.zeros(shape, source.dtype)
    hi = np.full(shape, 255 if source.dtype == np.uint8 else 1.0, source.dtype)
    s0 = zo.insert(lo, source, rect, angle, method)
    s1 = zo.insert(hi, source, rect, angle, method)
    same = (s0 == s1)
    written = same if same.ndim == 2 else same.all(axis=2)
    if sfmt == zo.PIX_RGBA8 and blend != "none":
        base = dest if dfmt == zo.PIX_RGBA8 else zo.convert(dest, zo.PIX_RGBA8)
        comp = base.copy()
        for r, c in zip(*np.nonzero(written)):
            comp[r, c] = zo.blend_rgba8(tuple(int(v) for v in base[r, c]), tuple(int(v) for v in s0[r, c]), blend)
        conv = comp if dfmt == zo.PIX_RGBA8 else zo.convert(comp, dfmt)
    else:
        conv = zo.convert(s0, dfmt)
    out = dest.copy()
    out[written] = conv[written]
    return out


_MIX = {"u8": ((24, 30), np.uint8), "f32": ((24, 30), np.float32), "rgb8": ((24, 30, 3), np.uint8), "rgba8": ((24, 30, 4), np.uint8),
        "rgbaf32": ((24, 30, 4), np.float32)}


@pytest.mark.parametrize("src_kind", list(_MIX))
@pytest.mark.parametrize("dst_kind", list(_MIX))
def test_insert_source_of_another_pixel_type(zb, src_kind, dst_kind):
    """Image.insert(source: anytype, ...) (transforms.zig:293): all 20 mixed pairs, rotated + scaled rectangle and the copy fast path,
    bilinear and bicubic; an Rgba(u8) source also under blend modes.  Bit-exact against the oracle composition."""
    if src_kind == dst_kind:
        pytest.skip("same pixel type: test_insert_* above")
    rng = np.random.default_rng(hash((src_kind, dst_kind)) % 2**32)
    dshape, ddt = _MIX[dst_kind]
    dest = rand_image(rng, (40, 52) + tuple(dshape[2:]), ddt)
    source = rand_image(rng, *_MIX[src_kind])
    cases = [((6.0, 5.0, 41.0, 33.0), 0.35, "bilinear", "none"), ((10.0, 8.0, 40.0, 32.0), 0.0, "nearest", "none"),      # the second is the copy path
             ((3.5, 2.0, 30.0, 38.0), -1.1, "bicubic", "none")]
    if src_kind == "rgba8":
        cases += [((6.0, 5.0, 41.0, 33.0), 0.35, "bilinear", "normal"), ((2.0, 4.0, 44.0, 30.0), 2.0, "bilinear", "multiply"),
                  ((10.0, 8.0, 40.0, 32.0), 0.0, "nearest", "screen")]
    for rect, angle, method, blend in cases:
        d = zb.Image.from_numpy(dest)
        d.insert(zb.Image.from_numpy(source), rect, angle, method_enum(zb, method), blend=zb.Blending[blend.upper()])
        want = _oracle_insert_mixed(dest, source, rect, angle, method, blend)
        got = d.to_numpy()
        assert zb.lib().zb_last_kernel().decode() == "insert_mixed"
        assert np.array_equal(got, want), (rect, angle, method, blend, int(np.count_nonzero(got != want)))
