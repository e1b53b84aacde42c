"""Runs ONE configuration a couple of times so that `ncu -k regex:<kernel> -c N` can capture it.
usage: python tools/gpu_profile_cfg.py rotate|boxblur|sharpen|bicubic|lanczos|bilinear|gemm|fdm|gauss8|conv3|canny|median|ssim|motion|convert"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import zignal_b200 as zb  # noqa: E402
from zignal_b200 import BorderMode, Image, Interpolation, PixFmt, matrix  # noqa: E402
from zignal_b200.fdm import FeatureDistributionMatching  # noqa: E402

L = zb.lib()
cfg = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
g = torch.Generator(device="cuda").manual_seed(1)
if cfg == "rotate":
    n = 8
    xs = torch.randint(0, 256, (n, 1080, 1920, 4), device="cuda", dtype=torch.uint8, generator=g)
    ys = torch.empty((n, 2122, 2122, 4), device="cuda", dtype=torch.uint8)
    a = np.float32(np.pi / 4)
    cs = (np.float32(np.cos(np.float64(a))), np.float32(np.sin(np.float64(a))))
    s0, d0 = Image.from_tensor(xs[0])._zb(), Image.from_tensor(ys[0])._zb()
    fn = lambda: L.zb_rotate_into_batch(s0, 1080 * 1920, d0, 2122 * 2122, n, 3, C.c_float(a), C.c_float(cs[0]), C.c_float(cs[1]), 1,
                                        C.c_float(1 / 3), C.c_float(1 / 3), 0, zb.image.current_stream())
elif cfg in ("boxblur", "sharpen", "gauss8", "conv3"):
    x = torch.randint(0, 256, (8192, 8192, 4), device="cuda", dtype=torch.uint8, generator=g)
    s, d = Image.from_tensor(x), Image.from_tensor(torch.empty_like(x))
    k3 = np.full((3, 3), 1 / 9, np.float32)
    fn = {"boxblur": lambda: s.box_blur(3, out=d), "sharpen": lambda: s.sharpen(3, out=d), "gauss8": lambda: s.gaussian_blur(2.25, out=d),
          "conv3": lambda: s.convolve(k3, BorderMode.MIRROR, out=d)}[cfg]
elif cfg == "sobelgray":
    x = torch.randint(0, 256, (8192, 8192), device="cuda", dtype=torch.uint8, generator=g)
    s, d = Image.from_tensor(x), Image.from_tensor(torch.empty_like(x))
    fn = lambda: s.sobel(out=d)
elif cfg == "svd256":
    xr = torch.randn(4096, 256, device="cuda", generator=g)
    cov = (xr.T @ xr / 4095).contiguous()
    fn = lambda: matrix.svd_device(cov.clone(), True, False)
elif cfg in ("gaussgray", "gaussrgb"):
    shape = (8192, 8192) if cfg == "gaussgray" else (8192, 8192, 3)
    x = torch.randint(0, 256, shape, device="cuda", dtype=torch.uint8, generator=g)
    s, d = Image.from_tensor(x), Image.from_tensor(torch.empty_like(x))
    fn = lambda: s.gaussian_blur(2.25, out=d)
elif cfg in ("bicubic", "lanczos", "bilinear"):
    x = torch.randint(0, 256, (16384, 16384, 3), device="cuda", dtype=torch.uint8, generator=g)
    big = Image.from_tensor(x)
    small = Image.init(4096, 4096, big.pixfmt)
    m = {"bicubic": Interpolation.BICUBIC, "lanczos": Interpolation.LANCZOS, "bilinear": Interpolation.BILINEAR}[cfg]
    fn = lambda: big.resize(small, m)
elif cfg == "gemm":
    X = torch.randn(1048576, 256, device="cuda", dtype=torch.float32, generator=g)
    fn = lambda: matrix.gemm_device(X, X, True, False, 1.0, 0.0, None)
elif cfg == "fdm":
    src = torch.randint(0, 256, (4096, 4096, 3), device="cuda", dtype=torch.uint8, generator=g)
    tgt = (torch.randint(0, 256, (4096, 4096, 3), device="cuda", dtype=torch.uint8, generator=g) // 2 + 40).to(torch.uint8)
    f = FeatureDistributionMatching(PixFmt.RGB8)
    si, ti = Image.from_tensor(src), Image.from_tensor(tgt)
    f.set_target(ti)
    f.set_source(si)
    fn = lambda: f.update()
elif cfg in ("canny", "median", "ssim", "motion", "convert"):             # the 8(f) additions at 4096 x 4096
    from zignal_b200.compose import motion_blur_linear
    gray = Image.from_tensor(torch.randint(0, 256, (4096, 4096), device="cuda", dtype=torch.uint8, generator=g))
    rgba = Image.from_tensor(torch.randint(0, 256, (4096, 4096, 4), device="cuda", dtype=torch.uint8, generator=g))
    rgba2 = Image.from_tensor(torch.randint(0, 256, (4096, 4096, 4), device="cuda", dtype=torch.uint8, generator=g))
    out8, outc = Image.init_like(gray), Image.init_like(rgba)
    outf = Image.init(4096, 4096, PixFmt.RGBAF32, device="cuda")
    fn = {"canny": lambda: gray.canny(1.4, 20.0, 60.0, out=out8), "median": lambda: rgba.median_blur(2, out=outc),
          "ssim": lambda: rgba.ssim(rgba2), "motion": lambda: motion_blur_linear(rgba, outc, 0.6, 15),
          "convert": lambda: rgba.convert(PixFmt.RGBAF32, out=outf)}[cfg]
else:
    raise SystemExit("unknown config " + cfg)
for _ in range(reps):
    fn()
torch.cuda.synchronize()
print(cfg, "kernel", L.zb_last_kernel().decode())
