"""GPU bring-up script for the convolution kernels: parity vs the oracle + first timings."""
import sys
import time
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
fails = 0


def rel(a, b):
    return float(np.max(np.abs(a.astype(np.float64) - b) / np.maximum(np.abs(b), 1e-6)))


print("device:", torch.cuda.get_device_name(0))
# ---- generic paths, bit exact
for shape, dt in [((37, 53), np.uint8), ((37, 53, 3), np.uint8), ((40, 41, 4), np.uint8), ((37, 53), np.float32), ((33, 29, 4), np.float32)]:
    for border in range(4):
        for (nx, ny) in [(1, 1), (3, 5), (7, 7), (4, 6), (15, 15)]:
            img = rng.integers(0, 256, shape).astype(dt) if dt == np.uint8 else rng.random(shape, dtype=np.float32)
            kx = rng.random(nx).astype(np.float32); kx /= kx.sum()
            ky = rng.random(ny).astype(np.float32); ky /= ky.sum()
            L.zb_set_force_generic(1)
            got = Image.from_numpy(img).convolve_separable(kx, ky, BorderMode(border)).to_numpy()
            L.zb_set_force_generic(0)
            want = zo.conv_separable(img, kx, ky, BN[border])
            if not np.array_equal(got, want):
                fails += 1
                print("GENERIC MISMATCH", shape, dt, border, nx, ny, np.abs(got.astype(np.float64) - want).max())
print("generic separable done, fails =", fails)

# ---- fused kernel
for (rows, cols) in [(64, 64), (96, 520), (300, 777), (513, 1030), (40, 16)]:
    for border in range(4):
        for half in [1, 3, 7, 8]:
            n = 2 * half + 1
            img = rng.random((rows, cols, 4), dtype=np.float32)
            k = rng.random(n).astype(np.float32) + 0.05; k /= k.sum()
            want = zo.conv_separable(img, k, k, BN[border])
            for exact in (1, 0):
                L.zb_set_exact_f32(exact)
                got = Image.from_numpy(img).convolve_separable(k, k, BorderMode(border)).to_numpy()
                kern = L.zb_last_kernel().decode()
                ok = np.array_equal(got, want) if exact else rel(got, want) <= 1e-5
                if not ok:
                    fails += 1
                    bad = np.argwhere(np.abs(got - want) > 1e-5 * np.maximum(np.abs(want), 1e-6))
                    print("FUSED MISMATCH", rows, cols, BN[border], half, "exact" if exact else "fma", kern, rel(got, want), "nbad", len(bad), bad[:4].tolist())
L.zb_set_exact_f32(0)
print("fused done, fails =", fails, "last kernel", L.zb_last_kernel().decode())
# views / even kernels through fused
img = rng.random((200, 300, 4), dtype=np.float32)
big = Image.from_numpy(img)
v = big.view(zb.Rectangle(8, 5, 290, 190))
k6 = rng.random(6).astype(np.float32); k6 /= k6.sum()
got = v.convolve_separable(k6, k6[:5] / k6[:5].sum(), BorderMode.MIRROR).to_numpy()
want = zo.conv_separable(np.ascontiguousarray(img[5:190, 8:290]), k6, k6[:5] / k6[:5].sum(), "mirror")
print("view+even kernel:", L.zb_last_kernel().decode(), rel(got, want))

# ---- timing at the headline size
def time_it(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))

R = C_ = 8192
x = torch.rand(R, C_, 4, device="cuda", dtype=torch.float32)
src = Image.from_tensor(x)
dst = Image.init_like(src)
taps = zb.gaussian_taps(2.25)
peak = 6490.5
for exact in (0, 1):
    L.zb_set_exact_f32(exact)
    med, mn = time_it(lambda: src.convolve_separable(taps, taps, BorderMode.MIRROR, out=dst))
    gbs = 2 * R * C_ * 16 / (med * 1e-3) / 1e9
    print(f"fused exact={exact} {L.zb_last_kernel().decode()}: median {med:.3f} ms min {mn:.3f} ms  {R*C_/med/1e3:.0f} Mpx/s  {gbs:.0f} GB/s  frac {gbs/peak:.3f}")
L.zb_set_exact_f32(0)
for half in (1, 3, 5, 8):
    t = zb.gaussian_taps(half / 3.0 - 0.01)
    med, mn = time_it(lambda: src.convolve_separable(t, t, BorderMode.MIRROR, out=dst))
    print(f"fused K={len(t)}: median {med:.3f} ms min {mn:.3f}")
L.zb_set_force_generic(1)
med, mn = time_it(lambda: src.convolve_separable(taps, taps, BorderMode.MIRROR, out=dst), n=5, warm=1)
print(f"generic two-pass: median {med:.3f} ms")
L.zb_set_force_generic(0)
# copy roofline on this box
y = torch.empty_like(x)
med, mn = time_it(lambda: y.copy_(x))
print(f"torch copy 1 GiB: {med:.3f} ms -> {2*x.numel()*4/med/1e6:.0f} GB/s")
print("TOTAL FAILS", fails)
