"""Bring-up of the tcgen05 3xTF32 X^T X kernel: accuracy vs f64, then timing at the PCA size."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import zignal_b200 as zb  # noqa: E402
from zignal_b200 import matrix  # noqa: E402

L = zb.lib()
g = torch.Generator(device="cuda").manual_seed(7)
for (n, dim) in [(4096, 256), (4096, 128), (10000, 256), (65536, 256), (100003, 128), (50000, 384), (40000, 512)]:
    X = torch.randn(n, dim, device="cuda", dtype=torch.float32, generator=g)
    X[:, 3] *= 100.0
    X[:, 7] += 5.0
    C = matrix.gemm_device(X, X, True, False, 1.0, 0.0, None)
    torch.cuda.synchronize()
    kern = L.zb_last_kernel().decode()
    ref = (X.double().T @ X.double())
    err = (C.double() - ref).abs()
    rel_max = float(err.max() / ref.abs().max())
    rel_el = float((err / ref.abs().clamp_min(1e-30)).max())
    big = (err / ref.abs().clamp_min(1.0)).max()
    print(f"n={n} dim={dim} kernel={kern}: max_abs_err/max|C| = {rel_max:.3e}  max elementwise rel = {rel_el:.3e}  err/(max(|ref|,1)) = {float(big):.3e}", flush=True)
    if rel_max > 1e-3:
        bad = (err / ref.abs().max() > 1e-3).nonzero()
        print("  BAD entries:", bad[:8].tolist(), "C", C[bad[0][0], bad[0][1]].item(), "ref", ref[bad[0][0], bad[0][1]].item())
# alpha / beta
X = torch.randn(8192, 256, device="cuda", dtype=torch.float32, generator=g)
C0 = torch.randn(256, 256, device="cuda", dtype=torch.float32, generator=g)
C = matrix.gemm_device(X, X, True, False, 0.5, 2.0, C0)
ref = 0.5 * (X.double().T @ X.double()) + 2.0 * C0.double()
print("alpha/beta:", float((C.double() - ref).abs().max() / ref.abs().max()))


def time_it(fn, n=10, warm=2):
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


X = torch.randn(1048576, 256, device="cuda", dtype=torch.float32, generator=g)
ms = time_it(lambda: matrix.gemm_device(X, X, True, False, 1.0 / (X.shape[0] - 1), 0.0, None))
fl = 2.0 * X.shape[0] * 256 * 256
print(f"PCA X^T X 1048576x256: {ms:.4f} ms  {fl/(ms*1e-3)/1e12:.1f} TFLOP/s (fp32-accurate)  {3*fl/(ms*1e-3)/1e12:.1f} TF32 TFLOP/s issued   kernel={L.zb_last_kernel().decode()}")
C = matrix.gemm_device(X, X, True, False, 1.0, 0.0, None)
ref = X.double().T @ X.double()
print("accuracy at n=1M:", float((C.double() - ref).abs().max() / ref.abs().max()), float(((C.double() - ref).abs() / ref.abs().clamp_min(1e-30)).max()))
Cb = torch.matmul(X.T, X)
print("cuBLAS fp32 accuracy at n=1M:", float((Cb.double() - ref).abs().max() / ref.abs().max()))
Xp = torch.rand(1048576, 256, device="cuda", dtype=torch.float32, generator=g) * 255.0   # uncentred pixel-like data: every sum is coherent
Cp = matrix.gemm_device(Xp, Xp, True, False, 1.0, 0.0, None)
refp = Xp.double().T @ Xp.double()
print("positive data n=1M: ours", float(((Cp.double() - refp).abs() / refp.abs()).max()), "cuBLAS fp32", float((((Xp.T @ Xp).double() - refp).abs() / refp.abs()).max()))
del Xp, Cp, refp
L.zb_set_force_generic(1)
ms2 = time_it(lambda: matrix.gemm_device(X, X, True, False, 1.0, 0.0, None), n=2, warm=1)
print(f"CUDA-core f64-accumulate path: {ms2:.3f} ms")
L.zb_set_force_generic(0)
ms3 = time_it(lambda: torch.matmul(X.T, X), n=5, warm=2)
print(f"torch (cuBLAS, default fp32 = no tf32): {ms3:.3f} ms")
