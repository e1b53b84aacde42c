"""Host-side mirror of zignal's `Image(T)` for the hot path (reference src/image.zig:97-1249).

Same method names, argument meaning and error behaviour as the reference's Zig API (snake_case as in
the reference's own Python binding, bindings/python/src/image/*.zig); every method body is a call
through the C ABI in include/zignal_b200.h.  Pixel storage lives in device memory (a torch CUDA
tensor is used purely as the owner of that memory and of the stream); `Image.from_numpy` /
`to_numpy` move data across PCIe, and the module-level `host_*` functions are the literal drop-in
for host-resident `Image.data` (H2D + op + D2H inside one C call).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from enum import IntEnum
from typing import Optional, Tuple

import numpy as np

from . import _ffi
from ._ffi import ZbImage, check, lib


class PixFmt(IntEnum):
    U8 = 0        # Image(u8)
    F32 = 1       # Image(f32)
    RGB8 = 2      # Image(Rgb(u8))
    RGBA8 = 3     # Image(Rgba(u8))
    RGBAF32 = 4   # Image(Rgba(f32))


class BorderMode(IntEnum):  # reference border.zig:10-19
    ZERO = 0
    REPLICATE = 1
    MIRROR = 2
    WRAP = 3


class Blending(IntEnum):  # reference blending.zig:8-22
    NONE = 0
    NORMAL = 1
    MULTIPLY = 2
    SCREEN = 3
    OVERLAY = 4
    SOFT_LIGHT = 5
    HARD_LIGHT = 6
    COLOR_DODGE = 7
    COLOR_BURN = 8
    DARKEN = 9
    LIGHTEN = 10
    DIFFERENCE = 11
    EXCLUSION = 12


class Interpolation(IntEnum):  # reference interpolation.zig:53-68
    NEAREST = 0
    BILINEAR = 1
    BICUBIC = 2
    CATMULL_ROM = 3
    MITCHELL = 4
    LANCZOS = 5


_CH = {PixFmt.U8: 1, PixFmt.F32: 1, PixFmt.RGB8: 3, PixFmt.RGBA8: 4, PixFmt.RGBAF32: 4}
_NP = {PixFmt.U8: np.uint8, PixFmt.F32: np.float32, PixFmt.RGB8: np.uint8, PixFmt.RGBA8: np.uint8, PixFmt.RGBAF32: np.float32}


def pixfmt_of_array(a) -> PixFmt:
    dt = np.dtype(str(a.dtype).replace("torch.", ""))
    nd = a.ndim
    ch = a.shape[2] if nd == 3 else 1
    if dt == np.uint8 and nd == 2:
        return PixFmt.U8
    if dt == np.uint8 and ch == 3:
        return PixFmt.RGB8
    if dt == np.uint8 and ch == 4:
        return PixFmt.RGBA8
    if dt == np.float32 and nd == 2:
        return PixFmt.F32
    if dt == np.float32 and ch == 4:
        return PixFmt.RGBAF32
    raise TypeError(f"unsupported pixel array: dtype={a.dtype} shape={tuple(a.shape)}")


def _np_image(a: np.ndarray) -> ZbImage:
    """zb_image over a (possibly row-strided) numpy array."""
    px_bytes = a.dtype.itemsize * (a.shape[2] if a.ndim == 3 else 1)
    if a.ndim == 3:
        assert a.strides[2] == a.dtype.itemsize and a.strides[1] == px_bytes, "pixels must be packed"
    elif a.shape[1] > 1:
        assert a.strides[1] == a.dtype.itemsize
    row = a.strides[0] if a.shape[0] > 1 else a.shape[1] * px_bytes
    assert row % px_bytes == 0
    return ZbImage(a.ctypes.data, a.shape[0], a.shape[1], row // px_bytes)


def _fptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _mitchell(method, b, c):
    return int(method), C.c_float(b), C.c_float(c)


def _torch():
    import torch
    return torch


def current_stream() -> int:
    torch = _torch()
    return torch.cuda.current_stream().cuda_stream


@dataclass
class Rectangle:  # reference geometry/Rectangle.zig: l, t, r, b (exclusive r, b)
    l: int
    t: int
    r: int
    b: int


class Image:
    """Device-resident mirror of zignal's Image(T) {rows, cols, data, stride} (image.zig:97-102)."""

    def __init__(self, tensor, pixfmt: PixFmt, rows: int, cols: int, stride: int, offset_px: int = 0):
        self._t = tensor          # torch tensor that owns the storage (flat, channel elements)
        self.pixfmt = PixFmt(pixfmt)
        self.rows, self.cols, self.stride = int(rows), int(cols), int(stride)
        self._off = int(offset_px)

    # ---- construction (image.zig:124-184) -------------------------------------------------------
    @classmethod
    def init(cls, rows: int, cols: int, pixfmt: PixFmt, device=None) -> "Image":
        torch = _torch()
        lib()  # fail loudly if the CUDA library is missing
        dt = torch.uint8 if _NP[PixFmt(pixfmt)] == np.uint8 else torch.float32
        t = torch.empty(int(rows) * int(cols) * _CH[PixFmt(pixfmt)], dtype=dt, device=device or "cuda")
        return cls(t, pixfmt, rows, cols, cols)

    @classmethod
    def init_like(cls, other: "Image") -> "Image":
        return cls.init(other.rows, other.cols, other.pixfmt, other._t.device)

    @classmethod
    def from_numpy(cls, a: np.ndarray, device=None) -> "Image":
        torch = _torch()
        lib()
        fmt = pixfmt_of_array(a)
        t = torch.from_numpy(np.ascontiguousarray(a)).to(device or "cuda").reshape(-1)
        return cls(t, fmt, a.shape[0], a.shape[1], a.shape[1])

    @classmethod
    def from_tensor(cls, t) -> "Image":
        """Wrap a contiguous CUDA tensor of shape (rows, cols[, ch]) without copying."""
        fmt = pixfmt_of_array(t)
        assert t.is_cuda and t.is_contiguous()
        return cls(t.reshape(-1), fmt, t.shape[0], t.shape[1], t.shape[1])

    def to_numpy(self) -> np.ndarray:
        ch = _CH[self.pixfmt]
        full = self._t.cpu().numpy()
        if self.rows == 0 or self.cols == 0:
            return np.zeros((self.rows, self.cols) + ((ch,) if ch > 1 else ()), _NP[self.pixfmt])
        idx = (self._off + np.arange(self.rows)[:, None] * self.stride + np.arange(self.cols)[None, :])
        if ch == 1:
            return full[idx]
        return full.reshape(-1, ch)[idx]

    def tensor(self):
        """The (rows, cols[, ch]) CUDA tensor of a contiguous, non-view image."""
        assert self.is_contiguous() and self._off == 0
        ch = _CH[self.pixfmt]
        return self._t.view(self.rows, self.cols, ch) if ch > 1 else self._t.view(self.rows, self.cols)

    # ---- views (image.zig:332-357) --------------------------------------------------------------
    def view(self, rect: Rectangle) -> "Image":
        l, t = max(0, rect.l), max(0, rect.t)
        r, b = min(self.cols, rect.r), min(self.rows, rect.b)
        if r <= l or b <= t:
            return Image(self._t, self.pixfmt, 0, 0, 0, 0)
        return Image(self._t, self.pixfmt, b - t, r - l, self.stride, self._off + t * self.stride + l)

    def is_contiguous(self) -> bool:
        return self.cols == self.stride

    def has_same_shape(self, other) -> bool:
        return self.rows == other.rows and self.cols == other.cols

    def get_center(self) -> Tuple[float, float]:  # image.zig:322-327
        return (np.float32(self.cols) / np.float32(2), np.float32(self.rows) / np.float32(2))

    def _zb(self) -> ZbImage:
        esz = self._t.element_size() * _CH[self.pixfmt]
        return ZbImage(self._t.data_ptr() + self._off * esz, self.rows, self.cols, self.stride)

    def copy(self, dst: "Image") -> None:  # image.zig:375-392
        assert self.has_same_shape(dst)
        self._peer(dst, what="dst")
        a, d = self._zb(), dst._zb()
        self._run(lib().zb_copy, a, d, int(self.pixfmt))

    def dupe(self) -> "Image":  # image.zig:367-371
        out = Image.init_like(self)
        self.copy(out)
        return out

    def _peer(self, other: "Image", fmt: Optional[PixFmt] = None, what: str = "out") -> "Image":
        """Zig's Image(T) typing makes a mismatched `out` a compile error; the C ABI takes one pixfmt for both buffers, so the
        mirror checks it here: same pixel type (or the op's fixed result type) and same CUDA device as self."""
        fmt = self.pixfmt if fmt is None else PixFmt(fmt)
        if not isinstance(other, Image):
            raise TypeError(f"{what} must be an Image")
        if other.pixfmt != fmt:
            raise TypeError(f"{what} must be an Image of pixel type {fmt.name}, got {other.pixfmt.name}")
        if other._t.device != self._t.device:
            raise ValueError(f"{what} lives on {other._t.device}, self on {self._t.device}")
        return other

    def _run(self, fn, *args) -> None:
        """Call a stream-taking entry point on self's device and on that device's current stream (not the stream of whatever
        device happens to be current)."""
        torch = _torch()
        dev = self._t.device
        with torch.cuda.device(dev):
            check(fn(*args, torch.cuda.current_stream(dev).cuda_stream))

    def _out(self, out: Optional["Image"], fmt: Optional[PixFmt] = None) -> "Image":
        if out is None:
            return Image.init(self.rows, self.cols, self.pixfmt if fmt is None else fmt, self._t.device)
        return self._peer(out, fmt)

    # ---- filters (image.zig:635-648, 785-799, 917-994) --------------------------------------------
    def box_blur(self, radius: int, out: Optional["Image"] = None) -> "Image":
        out = self._out(out)
        a, d = self._zb(), out._zb()
        self._run(lib().zb_box_blur, a, d, int(self.pixfmt), int(radius))
        return out

    def sharpen(self, radius: int, out: Optional["Image"] = None) -> "Image":
        out = self._out(out)
        a, d = self._zb(), out._zb()
        self._run(lib().zb_sharpen, a, d, int(self.pixfmt), int(radius))
        return out

    def convolve(self, kernel, border: BorderMode = BorderMode.MIRROR, out: Optional["Image"] = None) -> "Image":
        k = np.ascontiguousarray(kernel, dtype=np.float32)
        if k.ndim != 2:
            raise ValueError("Kernel must be a 2D array")  # convolution.zig:200
        out = self._out(out)
        a, d = self._zb(), out._zb()
        self._run(lib().zb_convolve, a, d, int(self.pixfmt), _fptr(k), k.shape[0], k.shape[1], int(border))
        return out

    def convolve_separable(self, kernel_x, kernel_y, border: BorderMode = BorderMode.MIRROR,
                           out: Optional["Image"] = None) -> "Image":
        kx = np.ascontiguousarray(kernel_x, dtype=np.float32)
        ky = np.ascontiguousarray(kernel_y, dtype=np.float32)
        out = self._out(out)
        a, d = self._zb(), out._zb()
        self._run(lib().zb_conv_separable, a, d, int(self.pixfmt), _fptr(kx), kx.size, _fptr(ky), ky.size, int(border))
        return out

    def gaussian_blur(self, sigma: float, out: Optional["Image"] = None) -> "Image":
        out = self._out(out)
        a, d = self._zb(), out._zb()
        self._run(lib().zb_gaussian_blur, a, d, int(self.pixfmt), C.c_float(sigma))
        return out

    # ---- resampling (image.zig:523-541) -----------------------------------------------------------
    def resize(self, out: "Image", method: Interpolation = Interpolation.BILINEAR, b: float = 1 / 3, c: float = 1 / 3) -> "Image":
        self._peer(out)
        a, d = self._zb(), out._zb()
        self._run(lib().zb_resize, a, d, int(self.pixfmt), int(method), C.c_float(b), C.c_float(c))
        return out

    def scale(self, factor: float, method: Interpolation = Interpolation.BILINEAR) -> "Image":
        if factor <= 0:
            raise _ffi.ZignalError(8, "InvalidScaleFactor")       # image.zig:531
        new_rows = int(_round_half_away(np.float32(self.rows) * np.float32(factor)))
        new_cols = int(_round_half_away(np.float32(self.cols) * np.float32(factor)))
        if new_rows == 0 or new_cols == 0:
            raise _ffi.ZignalError(9, "InvalidDimensions")        # image.zig:536
        return self.resize(Image.init(new_rows, new_cols, self.pixfmt, self._t.device), method)

    # ---- geometry (image.zig:558-623) -------------------------------------------------------------
    def rotate_bounds(self, angle: float) -> Tuple[int, int]:
        r, c = C.c_uint32(), C.c_uint32()
        check(lib().zb_rotate_bounds(self.rows, self.cols, C.c_float(angle), C.byref(r), C.byref(c)))
        return r.value, c.value

    def rotate_into(self, out: "Image", angle: float, method: Interpolation = Interpolation.BILINEAR,
                    border: BorderMode = BorderMode.ZERO, cos_sin=None, b: float = 1 / 3, c: float = 1 / 3) -> "Image":
        self._peer(out)
        a, d = self._zb(), out._zb()
        if cos_sin is None:
            self._run(lib().zb_rotate_into, a, d, int(self.pixfmt), C.c_float(angle), int(method), C.c_float(b), C.c_float(c), int(border))
        else:
            self._run(lib().zb_rotate_into_cs, a, d, int(self.pixfmt), C.c_float(angle), C.c_float(cos_sin[0]), C.c_float(cos_sin[1]),
                                          int(method), C.c_float(b), C.c_float(c), int(border))
        return out

    def rotate(self, angle: float, method: Interpolation = Interpolation.BILINEAR, border: BorderMode = BorderMode.ZERO,
               cos_sin=None) -> "Image":
        rows, cols = self.rotate_bounds(angle)
        out = Image.init(rows, cols, self.pixfmt, self._t.device)
        return self.rotate_into(out, angle, method, border, cos_sin)

    def sobel(self, out: Optional["Image"] = None) -> "Image":
        """Image.sobel (image.zig:999-1009): gradient magnitude into an Image(u8) of the same shape."""
        out = self._out(out, PixFmt.U8)
        a, d = self._zb(), out._zb()
        self._run(lib().zb_sobel, a, d, int(self.pixfmt))
        return out

    def convert(self, target: PixFmt, out: Optional["Image"] = None) -> "Image":
        """Image.convert(allocator, TargetType) / convertInto (image.zig:396-421): per-pixel convertColor into another pixel type."""
        target = PixFmt(target)
        out = self._out(out, target)
        a, d = self._zb(), out._zb()
        self._run(lib().zb_convert, a, int(self.pixfmt), d, int(target))
        return out

    # ---- quality metrics (image.zig:1105-1147, metrics.zig) ----
    def _metric(self, fn, other: "Image") -> float:
        self._peer(other, what="other")  # Image(T).psnr(other: Image(T))
        out = C.c_double(0.0)
        a, b = self._zb(), other._zb()
        self._run(fn, a, b, int(self.pixfmt), C.byref(out))
        return out.value

    def psnr(self, other: "Image") -> float:
        """Image.psnr (image.zig:1105, metrics.zig:10-54): dB, inf for identical images."""
        return self._metric(lib().zb_psnr, other)

    def ssim(self, other: "Image") -> float:
        """Image.ssim (image.zig:1126, metrics.zig:56-114): mean SSIM over the interior, 11x11 Gaussian window (sigma 1.5)."""
        return self._metric(lib().zb_ssim, other)

    def mean_pixel_error(self, other: "Image") -> float:
        """Image.meanPixelError (image.zig:1145, metrics.zig:116-165): mean absolute component difference / component range."""
        return self._metric(lib().zb_mean_pixel_error, other)

    # ---- order-statistic filters (image.zig:650-790, order_statistic_blur.zig) ----
    def _order(self, radius: int, mode: int, param: float, border: BorderMode, out: Optional["Image"]) -> "Image":
        out = self._out(out)
        a, d = self._zb(), out._zb()
        self._run(lib().zb_order_blur, a, d, int(self.pixfmt), int(radius), mode, C.c_double(param), int(border))
        return out

    def percentile_blur(self, radius: int, percentile: float, border: BorderMode = BorderMode.MIRROR, out: Optional["Image"] = None) -> "Image":
        """Image.percentileBlur (image.zig:672-684)."""
        return self._order(radius, 0, percentile, border, out)

    def median_blur(self, radius: int, out: Optional["Image"] = None) -> "Image":
        """Image.medianBlur (image.zig:650-658): percentile 0.5 with .mirror (order_statistic_blur.zig:28)."""
        return self._order(radius, 0, 0.5, BorderMode.MIRROR, out)

    def min_blur(self, radius: int, border: BorderMode = BorderMode.MIRROR, out: Optional["Image"] = None) -> "Image":
        """Image.minBlur (image.zig:696-707): erosion."""
        return self._order(radius, 0, 0.0, border, out)

    def max_blur(self, radius: int, border: BorderMode = BorderMode.MIRROR, out: Optional["Image"] = None) -> "Image":
        """Image.maxBlur (image.zig:719-730): dilation."""
        return self._order(radius, 0, 1.0, border, out)

    def midpoint_blur(self, radius: int, border: BorderMode = BorderMode.MIRROR, out: Optional["Image"] = None) -> "Image":
        """Image.midpointBlur (image.zig:742-753)."""
        return self._order(radius, 1, 0.0, border, out)

    def alpha_trimmed_mean_blur(self, radius: int, trim_fraction: float, border: BorderMode = BorderMode.MIRROR,
                                out: Optional["Image"] = None) -> "Image":
        """Image.alphaTrimmedMeanBlur (image.zig:767-779)."""
        return self._order(radius, 2, trim_fraction, border, out)

    def canny(self, sigma: float, low_threshold: float, high_threshold: float, out: Optional["Image"] = None) -> "Image":
        """Image.canny (image.zig:1041-1063, edges.zig:212-274): binary (0 / 255) edge map into an Image(u8) of the same shape."""
        out = self._out(out, PixFmt.U8)
        a, d = self._zb(), out._zb()
        self._run(lib().zb_canny, a, d, int(self.pixfmt), C.c_float(sigma), C.c_float(low_threshold), C.c_float(high_threshold))
        return out

    def extract(self, out: "Image", rect, angle: float = 0.0, method: Interpolation = Interpolation.BILINEAR,
                border: BorderMode = BorderMode.ZERO, b: float = 1 / 3, c: float = 1 / 3) -> "Image":
        """Image.extract (transforms.zig:232-283): rect = (l, t, r, b) floats in source coordinates, rotated by `angle` CCW."""
        self._peer(out)
        a32 = np.float32(angle)
        cos_a, sin_a = np.cos(a32, dtype=np.float32), np.sin(a32, dtype=np.float32)
        a, d = self._zb(), out._zb()
        self._run(lib().zb_extract, a, d, int(self.pixfmt), C.c_float(rect[0]), C.c_float(rect[1]), C.c_float(rect[2]), C.c_float(rect[3]),
                               C.c_float(a32), C.c_float(cos_a), C.c_float(sin_a), int(method), C.c_float(b), C.c_float(c), int(border))
        return out

    def insert(self, source: "Image", rect, angle: float = 0.0, method: Interpolation = Interpolation.BILINEAR,
               b: float = 1 / 3, c: float = 1 / 3, blend: "Blending" = Blending.NONE) -> "Image":
        """Image.insert(source, rect, angle, method, blend_mode) (transforms.zig:293-376): modifies self in place.  Rgba(u8) samples are
        composited under a blend mode (blending.zig:26-156); other pixel types are assigned (image.zig:67-95)."""
        a32 = np.float32(angle)
        cos_a, sin_a = np.cos(a32, dtype=np.float32), np.sin(a32, dtype=np.float32)
        self._peer(source, fmt=source.pixfmt, what="source")          # `source: anytype`: any pixel type, same device
        d, s = self._zb(), source._zb()
        self._run(lib().zb_insert_from, d, int(self.pixfmt), s, int(source.pixfmt), C.c_float(rect[0]), C.c_float(rect[1]), C.c_float(rect[2]),
                  C.c_float(rect[3]), C.c_float(a32), C.c_float(cos_a), C.c_float(sin_a), int(method), C.c_float(b), C.c_float(c), int(blend))
        return self

    def crop(self, rect) -> "Image":
        """Image.crop (transforms.zig:216-222): round(height) x round(width) chip, out-of-bounds pixels zero."""
        def rnd(v):  # @round on f32, half away from zero
            v = np.float32(v)
            return int(np.sign(v) * np.floor(np.abs(np.float64(v)) + 0.5))
        l, t, r, b = (np.float32(v) for v in rect)
        rows, cols = rnd(np.float32(0) if t >= b else b - t), rnd(np.float32(0) if l >= r else r - l)   # Rectangle.height / width
        chip = Image.init(max(rows, 0), max(cols, 0), self.pixfmt, device=self._t.device)
        return self.extract(chip, rect, 0.0, Interpolation.NEAREST, BorderMode.ZERO)

    def warp(self, out: "Image", transform, method: Interpolation = Interpolation.BILINEAR, b: float = 1 / 3, c: float = 1 / 3) -> "Image":
        self._peer(out)
        kind, m = transform.as_f32()
        a, d = self._zb(), out._zb()
        self._run(lib().zb_warp, a, d, int(self.pixfmt), kind, _fptr(m), int(method), C.c_float(b), C.c_float(c))
        return out


def _round_half_away(v) -> float:
    v = float(v)
    return math.floor(abs(v) + 0.5) * (1 if v >= 0 else -1)


# ---------------------------------------------------------------------------------------------------
# Host-resident drop-ins: numpy in, numpy out; H2D + kernel + D2H happen inside one C-ABI call.
# ---------------------------------------------------------------------------------------------------
def _host_call(fn_name, src: np.ndarray, out: np.ndarray, *args):
    a, d = _np_image(src), _np_image(out)
    check(getattr(lib(), fn_name)(a, d, int(pixfmt_of_array(src)), *args))
    return out


def host_gaussian_blur(src: np.ndarray, sigma: float, out: Optional[np.ndarray] = None) -> np.ndarray:
    out = np.empty_like(src) if out is None else out
    return _host_call("zb_host_gaussian_blur", src, out, C.c_float(sigma))


def host_conv_separable(src, kx, ky, border=BorderMode.MIRROR, out=None) -> np.ndarray:
    out = np.empty_like(src) if out is None else out
    kx = np.ascontiguousarray(kx, dtype=np.float32)
    ky = np.ascontiguousarray(ky, dtype=np.float32)
    return _host_call("zb_host_conv_separable", src, out, _fptr(kx), kx.size, _fptr(ky), ky.size, int(border))


def host_convolve(src, kernel, border=BorderMode.MIRROR, out=None) -> np.ndarray:
    out = np.empty_like(src) if out is None else out
    k = np.ascontiguousarray(kernel, dtype=np.float32)
    return _host_call("zb_host_convolve", src, out, _fptr(k), k.shape[0], k.shape[1], int(border))


def host_box_blur(src, radius, out=None) -> np.ndarray:
    out = np.empty_like(src) if out is None else out
    return _host_call("zb_host_box_blur", src, out, int(radius))


def host_sharpen(src, radius, out=None) -> np.ndarray:
    out = np.empty_like(src) if out is None else out
    return _host_call("zb_host_sharpen", src, out, int(radius))


def host_resize(src, out_shape, method=Interpolation.BILINEAR, b=1 / 3, c=1 / 3, out=None) -> np.ndarray:
    if out is None:
        out = np.zeros((out_shape[0], out_shape[1]) + tuple(src.shape[2:]), dtype=src.dtype)
    return _host_call("zb_host_resize", src, out, int(method), C.c_float(b), C.c_float(c))


def host_rotate(src, angle, method=Interpolation.BILINEAR, border=BorderMode.ZERO) -> np.ndarray:
    r, c = C.c_uint32(), C.c_uint32()
    check(lib().zb_rotate_bounds(src.shape[0], src.shape[1], C.c_float(angle), C.byref(r), C.byref(c)))
    out = np.zeros((r.value, c.value) + tuple(src.shape[2:]), dtype=src.dtype)
    return _host_call("zb_host_rotate_into", src, out, C.c_float(angle), int(method), C.c_float(1 / 3), C.c_float(1 / 3), int(border))


def host_warp(src, out, transform, method=Interpolation.BILINEAR) -> np.ndarray:
    kind, m = transform.as_f32()
    return _host_call("zb_host_warp", src, out, kind, _fptr(m), int(method), C.c_float(1 / 3), C.c_float(1 / 3))


def gaussian_taps(sigma: float) -> np.ndarray:
    """Host math of gaussianBlur (image.zig:972-990)."""
    n = C.c_int()
    buf = np.zeros(8192, np.float32)
    check(lib().zb_gaussian_taps(C.c_float(sigma), _fptr(buf), buf.size, C.byref(n)))
    return buf[: n.value].copy()
