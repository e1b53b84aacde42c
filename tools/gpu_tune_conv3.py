"""Third sweep: FFMA vs FFMA2 after the addressing rewrite."""
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


for stages in (2, 3):
    for f2 in (0, 1):
        L.zb_tune(b"conv.stages", stages)
        L.zb_tune(b"conv.f32x2", f2)
        for border in (BorderMode.ZERO, BorderMode.MIRROR):
            for half in (3, 5, 6, 7, 8):
                t = (np.random.default_rng(0).random(2 * half + 1) + 0.1).astype(np.float32)
                t /= t.sum()
                ms = time_it(lambda: src.convolve_separable(t, t, border, out=dst))
                gbs = 2 * R * C_ * 16 / (ms * 1e-3) / 1e9
                print(f"stages={stages} f32x2={f2} border={border.name} K={2*half+1}: {ms:.4f} ms {gbs:.0f} GB/s frac {gbs/PEAK:.3f}", flush=True)
L.zb_set_exact_f32(1)
t = zb.gaussian_taps(2.25)
print(f"exact K=15 mirror: {time_it(lambda: src.convolve_separable(t, t, BorderMode.MIRROR, out=dst)):.4f} ms")
L.zb_set_exact_f32(0)
