"""Sweep of the fused separable-convolution kernel's tuning knobs at the headline size (one gpurun call)."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import zignal_b200 as zb  # noqa: E402
from zignal_b200 import BorderMode, Image  # noqa: E402

L = zb.lib()
R = C_ = 8192
x = torch.rand(R, C_, 4, device="cuda", dtype=torch.float32)
src = Image.from_tensor(x)
dst = Image.init_like(src)
PEAK = 6490.5


def time_it(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


taps15 = zb.gaussian_taps(2.25)
print("device", torch.cuda.get_device_name(0))
for stages in (2, 3):
    for f2 in (0, 1):
        for band in (128, 256, 512, 1024):
            L.zb_tune(b"conv.stages", stages)
            L.zb_tune(b"conv.f32x2", f2)
            L.zb_tune(b"conv.band_rows", band)
            ms = time_it(lambda: src.convolve_separable(taps15, taps15, BorderMode.MIRROR, out=dst))
            gbs = 2 * R * C_ * 16 / (ms * 1e-3) / 1e9
            print(f"K=15 stages={stages} f32x2={f2} band={band}: {ms:.4f} ms  {R*C_/ms/1e3:.0f} Mpx/s  {gbs:.0f} GB/s  frac {gbs/PEAK:.3f}", flush=True)
L.zb_tune(b"conv.band_rows", 256)
for stages in (2, 3):
    for f2 in (0, 1):
        L.zb_tune(b"conv.stages", stages)
        L.zb_tune(b"conv.f32x2", f2)
        for half in (1, 3, 5, 7, 8):
            t = (np.random.default_rng(0).random(2 * half + 1) + 0.1).astype(np.float32)
            t /= t.sum()
            ms = time_it(lambda: src.convolve_separable(t, t, BorderMode.MIRROR, out=dst))
            print(f"K={2*half+1} stages={stages} f32x2={f2}: {ms:.4f} ms", flush=True)
for border in (BorderMode.ZERO, BorderMode.REPLICATE, BorderMode.WRAP):
    L.zb_tune(b"conv.stages", 3)
    L.zb_tune(b"conv.f32x2", 0)
    ms = time_it(lambda: src.convolve_separable(taps15, taps15, border, out=dst))
    print(f"K=15 border={border.name}: {ms:.4f} ms", flush=True)
y = torch.empty_like(x)
print(f"torch copy: {time_it(lambda: y.copy_(x)):.4f} ms")
# other ops, first numbers
im = Image.from_tensor(torch.randint(0, 256, (4096, 4096, 4), device="cuda", dtype=torch.uint8))
o = Image.init_like(im)
print(f"box_blur r=3 4096^2 RGBA8: {time_it(lambda: im.box_blur(3, out=o), n=5, warm=1):.3f} ms")
print(f"gaussian sigma=2.25 4096^2 RGBA8 (generic): {time_it(lambda: im.gaussian_blur(2.25, out=o), n=5, warm=1):.3f} ms")
big = Image.from_tensor(torch.randint(0, 256, (16384, 16384, 3), device="cuda", dtype=torch.uint8))
small = Image.init(4096, 4096, big.pixfmt)
ms = time_it(lambda: big.resize(small, zb.Interpolation.BICUBIC), n=5, warm=1)
print(f"C3 bicubic 16384^2->4096^2 RGB8: {ms:.3f} ms  {855638016/ms/1e6:.0f} GB/s")
import ctypes as C
n = 64
xs = torch.randint(0, 256, (n, 1080, 1920, 4), device="cuda", dtype=torch.uint8)
ys = torch.empty((n, 2122, 2122, 4), device="cuda", dtype=torch.uint8)
a = np.float32(np.pi / 4)
cs = (np.float32(np.cos(np.float64(a))), np.float32(np.sin(np.float64(a))))
s0, d0 = Image.from_tensor(xs[0])._zb(), Image.from_tensor(ys[0])._zb()
fn = lambda: L.zb_rotate_into_batch(s0, 1080 * 1920, d0, 2122 * 2122, n, 3, C.c_float(a), C.c_float(cs[0]), C.c_float(cs[1]), 1, C.c_float(1 / 3),
                                    C.c_float(1 / 3), 0, zb.image.current_stream())
ms = time_it(fn, n=5, warm=1)
print(f"C4 rotate45 bilinear batch {n} x 1080p RGBA8: {ms:.3f} ms  {n*26305936/ms/1e6:.0f} GB/s")
