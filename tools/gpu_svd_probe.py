"""One Pca-sized SVD (256 x 256 f32 covariance, device pointers) per Jacobi kernel, for `ncu --metrics gpu__time_duration.sum`; prints sweeps."""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import zignal_b200 as zb
from zignal_b200 import matrix
rng = np.random.default_rng(3)
x = rng.standard_normal((4096, 256)).astype(np.float32)
cov = torch.from_numpy((x.T @ x / 4095).astype(np.float32)).cuda()
for knob in (1, 0):
    zb.lib().zb_tune(b"jacobi.cluster", knob)
    u, s, v, conv = matrix.svd_device(cov.clone(), True, False)
    torch.cuda.synchronize()
    print("cluster" if knob else "cooperative", zb.lib().zb_last_kernel().decode(), "sweeps", zb.lib().zb_last_sweeps(), "conv", conv, flush=True)
zb.lib().zb_tune(b"jacobi.cluster", 1)
