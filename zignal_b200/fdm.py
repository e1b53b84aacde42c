"""Host-side mirror of zignal's FeatureDistributionMatching(T) (reference src/fdm.zig:19-275): same
method names, state machine and errors; statistics reduction and the per-pixel map run on the GPU."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._ffi import check, lib
from .image import Image, PixFmt, _np_image, current_stream, pixfmt_of_array


class FeatureDistributionMatching:
    def __init__(self, pixfmt: PixFmt):  # .init(allocator), fdm.zig:42
        self._h = C.c_void_p()
        self.pixfmt = PixFmt(pixfmt)
        check(lib().zb_fdm_create(C.byref(self._h), int(pixfmt)))
        self._source = None

    def deinit(self):  # fdm.zig:61
        if self._h:
            lib().zb_fdm_destroy(self._h)
            self._h = C.c_void_p()

    def set_target(self, target: Image):  # fdm.zig:68
        t = target._zb()
        check(lib().zb_fdm_set_target(self._h, t, current_stream()))

    def set_source(self, source: Image):  # fdm.zig:127
        self._source = source
        s = source._zb()
        check(lib().zb_fdm_set_source(self._h, s))

    def update(self):  # fdm.zig:141 (modifies the source image in place; the three kernels are queued, nothing waits)
        check(lib().zb_fdm_update(self._h, current_stream()))

    def status(self):
        """Wait for the stream and raise SvdFailed if the device-side 3x3 solve of the last update did not converge."""
        check(lib().zb_fdm_status(self._h, current_stream()))

    def match(self, source: Image, target: Image):  # fdm.zig:133
        self.set_target(target)
        self.set_source(source)
        self.update()

    # ---- multi-GPU: statistics as exact integer moments that one all-reduce can combine ----------
    @staticmethod
    def moments(img: Image, as_luma: bool = False) -> np.ndarray:
        sums = (C.c_uint64 * 11)()
        a = img._zb()
        check(lib().zb_fdm_moments(a, int(img.pixfmt), int(as_luma), sums, current_stream()))
        return np.array(list(sums), dtype=np.uint64)

    def set_target_moments(self, sums: np.ndarray):
        arr = (C.c_uint64 * 11)(*[int(v) for v in sums])
        check(lib().zb_fdm_set_target_moments(self._h, arr))

    def update_with_moments(self, sums: np.ndarray):
        arr = (C.c_uint64 * 11)(*[int(v) for v in sums])
        check(lib().zb_fdm_update_with_moments(self._h, arr, current_stream()))


def host_match(source: np.ndarray, target: np.ndarray) -> np.ndarray:
    """fdm.match on host arrays (the source copy is returned; the reference works in place)."""
    out = np.ascontiguousarray(source).copy()
    tgt = np.ascontiguousarray(target)
    s, t = _np_image(out), _np_image(tgt)
    check(lib().zb_host_fdm_match(s, t, int(pixfmt_of_array(out))))
    return out
