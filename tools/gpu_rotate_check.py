"""C4 (rotate 45 deg, bilinear, .zero, batch of 1080p Rgba frames): times the shared-memory tile kernel against the gather kernel and
checks that the two produce the same bits on the whole batch.  usage: python tools/gpu_rotate_check.py [frames] [angle]"""
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import zignal_b200 as zb  # noqa: E402
from zignal_b200 import Image  # noqa: E402

L = zb.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
angle = np.float32(float(sys.argv[2]) if len(sys.argv) > 2 else np.pi / 4)
g = torch.Generator(device="cuda").manual_seed(1)
xs = torch.randint(0, 256, (n, 1080, 1920, 4), device="cuda", dtype=torch.uint8, generator=g)
orows, ocols = Image.from_tensor(xs[0]).rotate_bounds(angle)
ys = torch.empty((n, orows, ocols, 4), device="cuda", dtype=torch.uint8)
ys2 = torch.empty_like(ys)
cs = (np.float32(np.cos(np.float64(angle))), np.float32(np.sin(np.float64(angle))))
s0 = Image.from_tensor(xs[0])._zb()


def run(dst):
    d0 = Image.from_tensor(dst[0])._zb()
    zb._ffi.check(L.zb_rotate_into_batch(s0, 1080 * 1920, d0, orows * ocols, n, 3, C.c_float(angle), C.c_float(cs[0]), C.c_float(cs[1]), 1,
                                         C.c_float(1 / 3), C.c_float(1 / 3), 0, zb.image.current_stream()))


def time_it(dst, reps=10):
    for _ in range(2):
        run(dst)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        run(dst)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, L.zb_last_kernel().decode()


algo = n * (1080 * 1920 * 4 + orows * ocols * 4)
out = {"frames": n, "angle": float(angle), "dst": [orows, ocols], "algorithmic_bytes": algo}
for name, knob, dst in (("tile", 1, ys), ("gather", 0, ys2)):
    zb._ffi.check(L.zb_tune(b"rotate.tile", knob))
    ms, k = time_it(dst)
    out[name] = {"ms": ms, "kernel": k, "us_per_frame": ms * 1e3 / n, "GB_per_s": algo / ms / 1e6}
zb._ffi.check(L.zb_tune(b"rotate.tile", 1))
out["bit_identical"] = bool(torch.equal(ys, ys2))
print(json.dumps(out))
sys.exit(0 if out["bit_identical"] else 1)
