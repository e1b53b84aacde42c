"""Bring-up of the warp-specialised fused kernel: parity first (small, under a short timeout), then timing."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib as zo  # noqa: E402
import zignal_b200 as zb  # noqa: E402
from zignal_b200 import BorderMode, Image  # noqa: E402

L = zb.lib()
rng = np.random.default_rng(0)
BN = {0: "zero", 1: "replicate", 2: "mirror", 3: "wrap"}
L.zb_tune(b"conv.variant", 1)
fails = 0
for (rows, cols) in [(64, 64), (96, 520), (300, 777), (1200, 1030)]:
    for border in range(4):
        for half in (1, 7, 8):
            img = rng.random((rows, cols, 4), dtype=np.float32)
            k = rng.random(2 * half + 1).astype(np.float32) + 0.05
            k /= k.sum()
            want = zo.conv_separable(img, k, k, BN[border])
            L.zb_set_exact_f32(1)
            got = Image.from_numpy(img).convolve_separable(k, k, BorderMode(border)).to_numpy()
            torch.cuda.synchronize()
            if not np.array_equal(got, want):
                fails += 1
                print("WS MISMATCH", rows, cols, BN[border], half, float(np.abs(got - want).max()), flush=True)
    print("ws parity", rows, cols, "fails", fails, flush=True)
L.zb_set_exact_f32(0)
R = C_ = 8192
x = torch.rand(R, C_, 4, device="cuda", dtype=torch.float32)
src = Image.from_tensor(x)
dst = Image.init_like(src)


def time_it(fn, n=30, warm=3):
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


for variant in (0, 1):
    L.zb_tune(b"conv.variant", variant)
    for border in (BorderMode.ZERO, BorderMode.MIRROR):
        for half in (3, 5, 6, 7, 8):
            t = (np.random.default_rng(0).random(2 * half + 1) + 0.1).astype(np.float32)
            t /= t.sum()
            ms = time_it(lambda: src.convolve_separable(t, t, border, out=dst))
            gbs = 2 * R * C_ * 16 / (ms * 1e-3) / 1e9
            print(f"variant={variant} border={border.name} K={2*half+1}: {ms:.4f} ms {gbs:.0f} GB/s frac {gbs/6490.5:.3f}", flush=True)
L.zb_tune(b"conv.variant", 1)
for band in (128, 192, 256, 384, 512):
    L.zb_tune(b"conv.band_rows", band)
    t = zb.gaussian_taps(2.25)
    ms = time_it(lambda: src.convolve_separable(t, t, BorderMode.MIRROR, out=dst))
    print(f"variant=1 mirror K=15 band={band}: {ms:.4f} ms frac {2*R*C_*16/(ms*1e-3)/1e9/6490.5:.3f}", flush=True)
print("TOTAL FAILS", fails)
