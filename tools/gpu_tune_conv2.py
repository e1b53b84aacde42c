"""Second sweep: borders x tap counts for the fused kernel after the fix-up rewrite."""
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


L.zb_tune(b"conv.stages", 2)
L.zb_tune(b"conv.f32x2", 0)
for border in (BorderMode.ZERO, BorderMode.MIRROR, BorderMode.REPLICATE, BorderMode.WRAP):
    for half in (1, 3, 5, 7, 8):
        t = (np.random.default_rng(0).random(2 * half + 1) + 0.1).astype(np.float32)
        t /= t.sum()
        ms = time_it(lambda: src.convolve_separable(t, t, border, out=dst))
        gbs = 2 * R * C_ * 16 / (ms * 1e-3) / 1e9
        print(f"border={border.name} K={2*half+1}: {ms:.4f} ms {gbs:.0f} GB/s frac {gbs/PEAK:.3f}", flush=True)
taps = zb.gaussian_taps(2.25)
for band in (192, 224, 256, 320, 384):
    L.zb_tune(b"conv.band_rows", band)
    ms = time_it(lambda: src.convolve_separable(taps, taps, BorderMode.MIRROR, out=dst))
    print(f"mirror K=15 band={band}: {ms:.4f} ms frac {2*R*C_*16/(ms*1e-3)/1e9/PEAK:.3f}", flush=True)
L.zb_tune(b"conv.band_rows", 256)
for stages in (2, 3):
    L.zb_tune(b"conv.stages", stages)
    ms = time_it(lambda: src.convolve_separable(taps, taps, BorderMode.MIRROR, out=dst))
    print(f"mirror K=15 stages={stages}: {ms:.4f} ms", flush=True)
# non-square / ragged sizes
for (r, c) in [(4320, 7680), (2160, 3840), (1080, 1920), (8192, 8191), (3000, 5000)]:
    xx = torch.rand(r, c, 4, device="cuda", dtype=torch.float32)
    s2 = Image.from_tensor(xx)
    d2 = Image.init_like(s2)
    L.zb_tune(b"conv.stages", 2)
    ms = time_it(lambda: s2.convolve_separable(taps, taps, BorderMode.MIRROR, out=d2))
    print(f"{r}x{c} mirror K=15: {ms:.4f} ms {2*r*c*16/(ms*1e-3)/1e9:.0f} GB/s  kernel={L.zb_last_kernel().decode()}", flush=True)
