"""Debug driver for the tcgen05 X^T X kernel: runs each shared-memory layout hypothesis in its own process."""
import os, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def child(layout: str) -> None:
    os.environ["ZB_TC_LAYOUT"] = layout
    sys.path.insert(0, str(ROOT))
    import torch
    import zignal_b200 as zb
    from zignal_b200 import matrix

    n, dim = 4096, 256
    X = torch.zeros(n, dim, device="cuda", dtype=torch.float32)
    X[:, :] = torch.arange(dim, device="cuda", dtype=torch.float32)[None, :] + 1.0  # C[i][j] = n (i+1)(j+1)
    os.environ["ZB_TC_DEBUG"] = "1"
    C = matrix.gemm_device(X, X, True, False, 1.0, 0.0, None)
    torch.cuda.synchronize()
    os.environ.pop("ZB_TC_DEBUG")
    print("layout", layout, "kernel", zb.lib().zb_last_kernel().decode())
    print("C[0,0:4]", C[0, :4].tolist(), "expect", [n * 1.0 * (j + 1) for j in range(4)])
    print("C[1,0:4]", C[1, :4].tolist(), "C[128,0:4]", C[128, :4].tolist(), "C[255,255]", C[255, 255].item(), "expect", n * 256.0 * 256.0)
    print("nonzero", int((C != 0).sum()), "nan", int(torch.isnan(C).sum()))
    for dim in (256, 128):
        g = torch.Generator(device="cuda").manual_seed(3)
        X = torch.randn(65536 + 40, dim, device="cuda", generator=g)
        C = matrix.gemm_device(X, X, True, False, 1.0, 0.0, None)
        ref = (X.double().T @ X.double())
        err = ((C.double() - ref).abs().max() / ref.abs().max()).item()
        print(f"random dim={dim}: kernel {zb.lib().zb_last_kernel().decode()} max_abs_err/max|C| = {err:.3e}")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(sys.argv[1])
    else:
        for lay in ("1", "0"):
            r = subprocess.run([sys.executable, __file__, lay], capture_output=True, text=True, timeout=90)
            print(r.stdout[-3000:])
            print(r.stderr[-1500:])
