"""Callers that are pure compositions of the hot path, kept device-resident across the chain:

* ImagePyramid.build   (reference src/image/pyramid.zig:31-104): per level, Gaussian blur of the ORIGINAL with an adaptive
  sigma, then a bilinear resize -- two kernels per level, no host round trip;
* motion_blur_linear   (reference src/image/motion_blur.zig:65-250): horizontal / vertical motion blur is convolveSeparable
  with a uniform kernel of `distance` taps against the identity, border .replicate; any other angle is the library's line-integral
  kernel (zb_motion_blur_linear decides, with the reference's 0.001 thresholds on |sin| and |cos|);
* motion_blur_radial   (:252-436): radial zoom / spin through zb_motion_blur_radial;
* letterbox / set_border_zero (src/image/transforms.zig:49-108, image.zig:198-229): resize into the content view + zeroed border.

The scalar set-up (level scale, sigma, sizes) is evaluated in f32 exactly as the reference writes it; `pow` is numpy's powf,
which may differ from Zig's std.math.pow in the last ulp (it only matters if a level size or a tap lands on a rounding edge).
"""
from __future__ import annotations

from typing import List

import numpy as np

import ctypes as C

from ._ffi import check, lib
from .image import Image, Interpolation, Rectangle, current_stream


class ImagePyramid:
    def __init__(self, levels: List[Image], scale_factor: float, blur_sigma: float):
        self.levels = levels
        self.scale_factor = np.float32(scale_factor)
        self.n_levels = len(levels)
        self.blur_sigma = np.float32(blur_sigma)

    @staticmethod
    def level_plan(rows: int, cols: int, n_levels: int, scale_factor: float, blur_sigma: float):
        """[(new_rows, new_cols, sigma or None)] for levels 1.. as pyramid.zig:55-89 computes them (f32 throughout)."""
        assert n_levels > 0 and scale_factor > 1.0 and blur_sigma > 0          # pyramid.zig:38-40
        sf, bs = np.float32(scale_factor), np.float32(blur_sigma)
        plan = []
        for i in range(1, n_levels):
            scale = np.power(sf, np.float32(i), dtype=np.float32)
            new_rows = max(1, int(np.trunc(np.float32(rows) / scale)))
            new_cols = max(1, int(np.trunc(np.float32(cols) / scale)))
            if new_rows < 8 or new_cols < 8:                                      # :61-71 truncate the pyramid here
                break
            sigma = np.float32(bs * np.sqrt(np.float32(scale * scale - np.float32(1.0))))
            plan.append((new_rows, new_cols, float(sigma) if sigma > np.float32(0.5) else None))
        return plan

    @classmethod
    def build(cls, source: Image, n_levels: int, scale_factor: float, blur_sigma: float) -> "ImagePyramid":
        levels = [source]                                                        # level 0 references the source (:52)
        for new_rows, new_cols, sigma in cls.level_plan(source.rows, source.cols, n_levels, scale_factor, blur_sigma):
            base = source.gaussian_blur(sigma) if sigma is not None else source   # blur the ORIGINAL, :75-83
            level = Image.init(new_rows, new_cols, source.pixfmt, device=source._t.device)
            base.resize(level, Interpolation.BILINEAR)                            # :89
            levels.append(level)
        return cls(levels, scale_factor, blur_sigma)

    @classmethod
    def build_default(cls, source: Image) -> "ImagePyramid":                      # :107-109
        return cls.build(source, 8, 1.2, 1.6)

    def get_scale(self, level: int) -> np.float32:                               # :121-124
        assert level < self.n_levels
        return np.power(self.scale_factor, np.float32(level), dtype=np.float32)


def motion_blur_linear(image: Image, out: Image, angle: float, distance: int) -> Image:
    """MotionBlur.linear (motion_blur.zig:65-250): copy for distance 0, convolveSeparable with a uniform kernel for axis-aligned motion,
    the per-pixel line integral of bilinear samples for any other angle.  cos / sin are taken in f32 on the host (`@cos(angle)`)."""
    a32 = np.float32(angle)
    cos_a, sin_a = np.cos(a32, dtype=np.float32), np.sin(a32, dtype=np.float32)
    a, d = image._zb(), out._zb()
    check(lib().zb_motion_blur_linear(a, d, int(image.pixfmt), C.c_float(a32), C.c_float(cos_a), C.c_float(sin_a), C.c_uint32(distance),
                                      current_stream()))
    return out


def motion_blur_radial(image: Image, out: Image, center_x: float, center_y: float, strength: float, spin: bool = False) -> Image:
    """MotionBlur.radial_zoom / radial_spin (motion_blur.zig:252-436): center in normalised [0, 1] coordinates, strength in [0, 1]."""
    a, d = image._zb(), out._zb()
    check(lib().zb_motion_blur_radial(a, d, int(image.pixfmt), C.c_float(center_x), C.c_float(center_y), C.c_float(strength), int(bool(spin)),
                                      current_stream()))
    return out


def set_border_zero(img: Image, rect: Rectangle) -> None:
    """Image.setBorder(rect, zeroes(T)) (image.zig:198-229)."""
    a = img._zb()
    check(lib().zb_set_border_zero(a, int(img.pixfmt), C.c_uint32(max(0, rect.l)), C.c_uint32(max(0, rect.t)), C.c_uint32(max(0, rect.r)),
                                   C.c_uint32(max(0, rect.b)), current_stream()))


def letterbox_rect(rows: int, cols: int, out_rows: int, out_cols: int):
    """The content rectangle Image.letterbox computes (transforms.zig:69-97, f32 throughout), or None when both scale factors are
    exactly equal (the aspect ratios match and the whole output is content)."""
    rows_scale = np.float32(out_rows) / np.float32(rows)
    cols_scale = np.float32(out_cols) / np.float32(cols)
    if rows_scale == cols_scale:
        return None
    aspect = min(rows_scale, cols_scale)

    def round_half_away(v: np.float32) -> int:                                   # @round on a non-negative f32
        return int(np.floor(np.float64(v) + 0.5))
    scaled_rows = round_half_away(np.float32(aspect * np.float32(rows)))        # :82-83
    scaled_cols = round_half_away(np.float32(aspect * np.float32(cols)))
    off_r = max(0, out_rows - scaled_rows) // 2                                  # -| is a saturating subtraction, :86-87
    off_c = max(0, out_cols - scaled_cols) // 2
    return Rectangle(off_c, off_r, off_c + scaled_cols, off_r + scaled_rows)


def letterbox(image: Image, out: Image, method: Interpolation = Interpolation.BILINEAR) -> Rectangle:
    """Image.letterbox (transforms.zig:46-108): aspect-preserving resize into the centre of `out`, zero padding around it.
    Returns the rectangle of `out` that holds the image content.  One resize into a view plus four strided memsets."""
    if out.rows == 0 or out.cols == 0:
        return Rectangle(0, 0, 0, 0)
    if image.rows == 0 or image.cols == 0:                                       # :57-60
        set_border_zero(out, Rectangle(0, 0, 0, 0))
        return Rectangle(0, 0, 0, 0)
    full = Rectangle(0, 0, out.cols, out.rows)
    if image.rows == out.rows and image.cols == out.cols:                        # :63-66
        image.copy(out)
        return full
    rect = letterbox_rect(image.rows, image.cols, out.rows, out.cols)
    if rect is None:                                                             # equal scale factors: plain resize, :73-76
        image.resize(out, method)
        return full
    view = out.view(rect)
    image.resize(view, method)                                                   # :101
    set_border_zero(out, rect)                                                   # :104
    return rect
