"""Seeded cases of the SURVEY 8(f) additions shared by the fixture generator (tests/golden/make_golden_8f.py), the CPU regression test of
the oracle (tests/test_golden_oracle.py) and the GPU fixture test (tests/test_gpu_golden.py).  Each case builds its input from numpy's
PCG64, names the oracle call and the equivalent call on the device mirror; only bit-exact operations are listed (spin blur and the f64
metrics are compared with tolerances elsewhere)."""
import numpy as np

import oracle_lib as zo
from gpu_utils import rand_image


def _img(seed, shape, dtype=np.uint8):
    return rand_image(np.random.default_rng(seed), shape, dtype)


def _smooth(seed, rows, cols):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:rows, 0:cols].astype(np.float64)
    base = 128 + 90 * np.sin(xx / 17.0) * np.cos(yy / 23.0) + rng.normal(0, 4.0, (rows, cols))
    return np.clip(base, 0, 255).astype(np.uint8)


# name -> (input builder, oracle call, device call).  Device calls take (zb, image array) and return a numpy array.
CASES = {
    "canny_gray_sigma1.4": (lambda: _smooth(61, 240, 317), lambda a: zo.canny(a, 1.4, 8.0, 24.0),
                            lambda zb, a: zb.Image.from_numpy(a).canny(1.4, 8.0, 24.0).to_numpy()),
    "canny_rgb_sigma0": (lambda: np.repeat(_smooth(62, 120, 150)[..., None], 3, axis=2), lambda a: zo.canny(a, 0.0, 40.0, 120.0),
                         lambda zb, a: zb.Image.from_numpy(a).canny(0.0, 40.0, 120.0).to_numpy()),
    "median_r2_rgba": (lambda: _img(63, (150, 201, 4)), lambda a: zo.order_blur(a, 2, "percentile", 0.5, "mirror"),
                       lambda zb, a: zb.Image.from_numpy(a).median_blur(2).to_numpy()),
    "percentile_r5_gray_wrap": (lambda: _img(64, (97, 131)), lambda a: zo.order_blur(a, 5, "percentile", 0.3, "wrap"),
                                lambda zb, a: zb.Image.from_numpy(a).percentile_blur(5, 0.3, zb.BorderMode.WRAP).to_numpy()),
    "alpha_trimmed_r3_rgb": (lambda: _img(65, (88, 120, 3)), lambda a: zo.order_blur(a, 3, "alpha_trimmed", 0.2, "replicate"),
                             lambda zb, a: zb.Image.from_numpy(a).alpha_trimmed_mean_blur(3, 0.2, zb.BorderMode.REPLICATE).to_numpy()),
    "midpoint_r1_gray_zero": (lambda: _img(66, (60, 77)), lambda a: zo.order_blur(a, 1, "midpoint", 0.0, "zero"),
                              lambda zb, a: zb.Image.from_numpy(a).midpoint_blur(1, zb.BorderMode.ZERO).to_numpy()),
    "motion_line_rgba": (lambda: _img(67, (120, 160, 4)), lambda a: zo.motion_blur_linear(a, 0.7, 9),
                         lambda zb, a: _motion_line(zb, a, 0.7, 9)),
    "motion_line_f32": (lambda: _img(68, (90, 110), np.float32), lambda a: zo.motion_blur_linear(a, -1.2, 6),
                        lambda zb, a: _motion_line(zb, a, -1.2, 6)),
    "motion_zoom_rgb": (lambda: _img(69, (100, 140, 3)), lambda a: zo.motion_blur_radial(a, 0.4, 0.6, 0.7),
                        lambda zb, a: _motion_zoom(zb, a, 0.4, 0.6, 0.7)),
    "sobel_rgba": (lambda: _img(70, (130, 170, 4)), lambda a: zo.sobel(a), lambda zb, a: zb.Image.from_numpy(a).sobel().to_numpy()),
    "convert_rgba8_to_u8": (lambda: _img(71, (90, 123, 4)), lambda a: zo.convert(a, 0), lambda zb, a: zb.Image.from_numpy(a).convert(zb.PixFmt.U8).to_numpy()),
    "convert_rgb8_to_rgbaf32": (lambda: _img(72, (70, 91, 3)), lambda a: zo.convert(a, 4),
                                lambda zb, a: zb.Image.from_numpy(a).convert(zb.PixFmt.RGBAF32).to_numpy()),
    "convert_rgbaf32_to_rgb8": (lambda: (_img(73, (64, 85, 4), np.float32) * np.float32(1.3) - np.float32(0.15)).astype(np.float32),
                                lambda a: zo.convert(a, 2), lambda zb, a: zb.Image.from_numpy(a).convert(zb.PixFmt.RGB8).to_numpy()),
    "insert_blend_soft_light": (lambda: _img(74, (90, 110, 4)), lambda a: zo.insert(a, _img(75, (40, 50, 4)), (12.0, 9.0, 92.0, 70.0), 0.4, "bilinear",
                                                                                    blend="soft_light"),
                                lambda zb, a: zb.Image.from_numpy(a.copy()).insert(zb.Image.from_numpy(_img(75, (40, 50, 4))), (12.0, 9.0, 92.0, 70.0), 0.4,
                                                                                   zb.Interpolation.BILINEAR, blend=zb.Blending.SOFT_LIGHT).to_numpy()),
}


def _motion_line(zb, a, angle, distance):
    from zignal_b200.compose import motion_blur_linear
    dev = zb.Image.from_numpy(a)
    return motion_blur_linear(dev, zb.Image.init_like(dev), angle, distance).to_numpy()


def _motion_zoom(zb, a, cx, cy, strength):
    from zignal_b200.compose import motion_blur_radial
    dev = zb.Image.from_numpy(a)
    return motion_blur_radial(dev, zb.Image.init_like(dev), cx, cy, strength).to_numpy()


def eigh_case():
    rng = np.random.default_rng(76)
    m = rng.normal(size=(12, 12))
    return (m + m.T) * 0.5
