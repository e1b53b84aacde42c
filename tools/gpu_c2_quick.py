"""C2 at a glance: the headline kernel for every border mode (CUDA events, 50 launches each)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import zignal_b200 as zb
from zignal_b200 import BorderMode, Image
x = torch.rand(8192, 8192, 4, device="cuda")
src = Image.from_tensor(x); dst = Image.init_like(src)
taps = zb.gaussian_taps(2.25)
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for knob in (1, 0, 1):
    zb.lib().zb_tune(b"conv.edge_fast", knob)
    for border in (BorderMode.MIRROR, BorderMode.REPLICATE, BorderMode.ZERO, BorderMode.WRAP):
        ms = t(lambda: src.convolve_separable(taps, taps, border, out=dst))
        print(f"edge_fast={knob} {border.name}: {ms:.4f} ms  {2 * 8192 * 8192 * 16 / ms / 1e6:.0f} GB/s  kernel {zb.lib().zb_last_kernel().decode()}", flush=True)
zb.lib().zb_tune(b"conv.edge_fast", 1)
y = torch.empty_like(x)
print(f"torch copy 1 GiB: {t(lambda: y.copy_(x)):.4f} ms")
