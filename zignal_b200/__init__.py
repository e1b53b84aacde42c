"""zignal_b200 -- B200 (sm_100a) implementation of zignal's per-pixel image hot path behind a C ABI.

`zignal_b200.lib/libzignal_b200.so` is the product; this package is the thin host-side mirror of the
reference's Image(T)/Matrix API over that ABI (see include/zignal_b200.h and INTEGRATION.md).
"""
from ._ffi import LibraryMissing, ZignalError, ZbImage, declared_symbols, lib  # noqa: F401
from .image import (Blending, BorderMode, Image, Interpolation, PixFmt, Rectangle, gaussian_taps, host_box_blur,  # noqa: F401
                    host_conv_separable, host_convolve, host_gaussian_blur, host_resize, host_rotate, host_sharpen,
                    host_warp)

from . import fdm, matrix, pca  # noqa: F401,E402

__version__ = "0.1.0"
