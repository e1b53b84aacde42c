"""Times every BASELINE config / §8 row on one GPU and writes gpurun_out/all_configs.json (copied to profiles/)."""
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import zignal_b200 as zb  # noqa: E402
from zignal_b200 import BorderMode, Image, Interpolation, PixFmt  # noqa: E402
from zignal_b200.fdm import FeatureDistributionMatching  # noqa: E402
from zignal_b200 import matrix  # noqa: E402

L = zb.lib()
PEAK = 6490.5
out = []


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


def rec(name, ms, algo_bytes, px, note=""):
    gbs = algo_bytes / (ms * 1e-3) / 1e9
    r = {"config": name, "ms": ms, "algorithmic_bytes": algo_bytes, "GB_per_s": gbs, "frac_of_measured_hbm_peak": gbs / PEAK,
         "Mpixels_per_s": px / 1e6 / (ms * 1e-3), "kernel": L.zb_last_kernel().decode(), "note": note}
    out.append(r)
    print(json.dumps(r), flush=True)


g = torch.Generator(device="cuda").manual_seed(1)
# C1: 3x3 box blur 512x512 u8
x = torch.randint(0, 256, (512, 512), device="cuda", dtype=torch.uint8, generator=g)
s, d = Image.from_tensor(x), Image.from_tensor(torch.empty_like(x))
rec("C1 boxBlur r=1 512x512 u8", time_it(lambda: s.box_blur(1, out=d), n=50), 2 * 512 * 512, 512 * 512, "latency-bound (3 tiny kernels)")
# C2
x = torch.rand(8192, 8192, 4, device="cuda", dtype=torch.float32, generator=g)
s, d = Image.from_tensor(x), Image.from_tensor(torch.empty_like(x))
rec("C2 gaussian 15x15 8192^2 RGBA f32", time_it(lambda: s.gaussian_blur(2.25, out=d)), 2 * x.numel() * 4, 8192 * 8192)
del x, s, d
# RGBA8 blur (the reference's most common real type)
x = torch.randint(0, 256, (8192, 8192, 4), device="cuda", dtype=torch.uint8, generator=g)
s, d = Image.from_tensor(x), Image.from_tensor(torch.empty_like(x))
rec("gaussian 15x15 8192^2 RGBA u8", time_it(lambda: s.gaussian_blur(2.25, out=d)), 2 * x.numel(), 8192 * 8192, "issue bound (120 MAC/px on the FP32 pipe)")
rec("boxBlur r=3 8192^2 RGBA u8", time_it(lambda: s.box_blur(3, out=d), n=5, warm=1), 2 * x.numel(), 8192 * 8192)
rec("sharpen r=3 8192^2 RGBA u8", time_it(lambda: s.sharpen(3, out=d), n=5, warm=1), 2 * x.numel(), 8192 * 8192)
k3 = np.full((3, 3), 1 / 9, np.float32)
rec("convolve 3x3 8192^2 RGBA u8", time_it(lambda: s.convolve(k3, BorderMode.MIRROR, out=d), n=5, warm=1), 2 * x.numel(), 8192 * 8192)
del x, s, d
# Rgb / gray u8 blur (shared-memory tile kernel) and Sobel
for shp, name in (((8192, 8192, 3), "RGB u8"), ((8192, 8192), "gray u8")):
    x = torch.randint(0, 256, shp, device="cuda", dtype=torch.uint8, generator=g)
    s, d = Image.from_tensor(x), Image.from_tensor(torch.empty_like(x))
    rec(f"gaussian 15x15 8192^2 {name}", time_it(lambda: s.gaussian_blur(2.25, out=d), n=5, warm=1), 2 * x.numel(), 8192 * 8192)
    e = Image.from_tensor(torch.empty((8192, 8192), device="cuda", dtype=torch.uint8))
    rec(f"sobel 8192^2 {name}", time_it(lambda: s.sobel(out=e), n=5, warm=1), x.numel() + 8192 * 8192, 8192 * 8192, "byte tiles: luma on the way in, integer gradients, f32 magnitude")
    del e
    del x, s, d
# C3
x = torch.randint(0, 256, (16384, 16384, 3), device="cuda", dtype=torch.uint8, generator=g)
big = Image.from_tensor(x)
small = Image.init(4096, 4096, big.pixfmt)
for m in (Interpolation.BICUBIC, Interpolation.BILINEAR, Interpolation.LANCZOS):
    rec(f"C3 resize {m.name} 16384^2->4096^2 RGB u8", time_it(lambda: big.resize(small, m), n=5, warm=1), 855638016, 4096 * 4096)
del x, big, small
# C4
n = 128
xs = torch.randint(0, 256, (n, 1080, 1920, 4), device="cuda", dtype=torch.uint8, generator=g)
ys = torch.empty((n, 2122, 2122, 4), device="cuda", dtype=torch.uint8)
a = np.float32(np.pi / 4)
cs = (np.float32(np.cos(np.float64(a))), np.float32(np.sin(np.float64(a))))
s0, d0 = Image.from_tensor(xs[0])._zb(), Image.from_tensor(ys[0])._zb()
fn = lambda: L.zb_rotate_into_batch(s0, 1080 * 1920, d0, 2122 * 2122, n, 3, C.c_float(a), C.c_float(cs[0]), C.c_float(cs[1]), 1, C.c_float(1 / 3),
                                    C.c_float(1 / 3), 0, zb.image.current_stream())
rec(f"C4 rotate45 bilinear zero, batch {n} x 1920x1080 RGBA u8", time_it(fn, n=5, warm=1), n * 26305936, n * 1080 * 1920)
del xs, ys
# C5 FDM 4096x4096 Rgb u8
src = torch.randint(0, 256, (4096, 4096, 3), device="cuda", dtype=torch.uint8, generator=g)
tgt = (torch.randint(0, 256, (4096, 4096, 3), device="cuda", dtype=torch.uint8, generator=g) // 2 + 40).to(torch.uint8)
f = FeatureDistributionMatching(PixFmt.RGB8)
si, ti = Image.from_tensor(src), Image.from_tensor(tgt)
f.set_target(ti)
f.set_source(si)
rec("C5 fdm.update 4096^2 Rgb u8 (statistics + in-kernel 3x3 solve, map)", time_it(lambda: f.update(), n=20, warm=3), 150994944, 4096 * 4096,
    "two launches, nothing leaves the device")
rec("C5 fdm.match (target stats too)", time_it(lambda: f.match(si, ti), n=10, warm=2), 150994944 + 50331648, 4096 * 4096)
f.deinit()
del src, tgt
# PCA GEMM: X^T X, n = 1,048,576 x 256 f32
X = torch.randn(1048576, 256, device="cuda", dtype=torch.float32, generator=g)
ms = time_it(lambda: matrix.gemm_device(X, X, True, False, 1.0 / (X.shape[0] - 1), 0.0, None), n=3, warm=1)
flops = 2.0 * X.shape[0] * 256 * 256
r = {"config": "PCA GEMM X^T X, n=1048576 x dim=256 f32 (tcgen05 3xTF32, fp32-accurate)", "ms": ms, "TFLOP_per_s": flops / (ms * 1e-3) / 1e12,
     "kernel": L.zb_last_kernel().decode(), "algorithmic_bytes": 1048576 * 256 * 4, "GB_per_s": 1048576 * 256 * 4 / (ms * 1e-3) / 1e9,
     "note": "useful fp32-accurate flops; 2.25x that is issued as TF32 MMAs (3 products on the upper-triangle tiles); shared-memory operand bandwidth bound"}
out.append(r)
print(json.dumps(r), flush=True)
# the SVD step of Pca.fit: 256 x 256 covariance, one-sided Jacobi on the device; and the end-to-end core of Pca.fit
mean = torch.empty(256, device="cuda")
cen = torch.empty_like(X)
matrix.center_columns(X, mean, True, cen)
cov = matrix.gemm_device(cen, cen, True, False, 1.0 / (X.shape[0] - 1), 0.0, None)
ms_svd = time_it(lambda: matrix.svd_device(cov, True, False), n=3, warm=1)
r = {"config": "SVD 256x256 f32 (one-sided Jacobi, persistent cooperative kernel; PCA covariance)", "ms": ms_svd, "kernel": L.zb_last_kernel().decode(),
     "note": "latency-bound: 255 rounds x ~8 sweeps, one grid barrier per round"}
out.append(r)
print(json.dumps(r), flush=True)


def fit():
    matrix.center_columns(X, mean, True, cen)
    matrix.svd_device(matrix.gemm_device(cen, cen, True, False, 1.0 / (X.shape[0] - 1), 0.0, None), True, False)


r = {"config": "Pca.fit device core (column means + centring + X^T X + SVD), n=1048576 x dim=256 f32", "ms": time_it(fit, n=3, warm=1),
     "kernel": L.zb_last_kernel().decode()}
out.append(r)
print(json.dumps(r), flush=True)
a64 = np.random.default_rng(0).standard_normal((200, 200))
a64 = (a64 + a64.T) / 2
import time
t0 = time.perf_counter()
matrix.eigh(a64)
r = {"config": "Matrix.eigh 200x200 f64 (two-sided Jacobi, host pointers in and out)", "ms": (time.perf_counter() - t0) * 1e3, "kernel": L.zb_last_kernel().decode()}
out.append(r)
print(json.dumps(r), flush=True)
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "all_configs.json").write_text(json.dumps(out, indent=1))
