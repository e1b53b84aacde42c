"""torchrun --nproc-per-node N tools/gpu_shard_diag.py : where does the step time of the sharded convolution go?
  a. the single-GPU kernel on this rank's (IPC-exported, peer-mapped) block, back to back
  b. the sharded kernel, back to back (what bench.py times)
  c. the sharded kernel, one launch at a time between device synchronisations (kernel duration without queueing effects)
  d. host time to enqueue 100 sharded steps"""
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import zignal_b200 as zb  # noqa: E402
from zignal_b200 import shard  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
comm = shard.ShardComm.from_torch_distributed()
taps = zb.gaussian_taps(2.25)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
src, dst = comm.image(rows, 8192, zb.PixFmt.RGBAF32, 8), comm.image(rows, 8192, zb.PixFmt.RGBAF32, 8)
src.interior_tensor().copy_(torch.rand(rows, 8192, 4, device=dev))


def loop(fn, n=100):
    for _ in range(5):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    host = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, host


def single():
    src.block_image().convolve_separable(taps, taps, zb.BorderMode.MIRROR, out=dst.block_image())


def sharded():
    src.conv_separable(dst, taps, taps, zb.BorderMode.MIRROR)


a, ha = loop(single)
b, hb = loop(sharded)
alt = {}
for name, path in (("pull kernel + plain kernel", 2), ("NCCL exchange + plain kernel", 1)):   # the other two paths, back to back
    zb._ffi.check(zb.lib().zb_shard_tune_path(path))
    alt[name] = loop(sharded)[0]
zb._ffi.check(zb.lib().zb_shard_tune_path(0))
print(f"rank {rank}: " + " | ".join(f"{k} {v:.4f} ms" for k, v in alt.items()), flush=True)
per = []
for _ in range(30):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sharded()
    e1.record()
    torch.cuda.synchronize()
    per.append(e0.elapsed_time(e1))
per.sort()
comm.status()
# where inside the kernel the time goes: stamps of the last launch (one at a time, ranks released together by the barrier)
import ctypes as C  # noqa: E402
stamps = (C.c_uint64 * 8)()
zb._ffi.check(zb.lib().zb_shard_debug_times(comm._h, stamps, zb.image.current_stream()))
t = [int(x) for x in stamps]
print(f"rank {rank}: kernel stamps (us after the first CTA started): neighbours ready {(t[1] - t[0]) / 1e3:.1f} | halo copied {(t[2] - t[0]) / 1e3:.1f} | "
      f"last CTA done {(t[3] - t[0]) / 1e3:.1f} | neighbours done reading {(t[4] - t[0]) / 1e3:.1f}", flush=True)
print(f"rank {rank}: single-GPU kernel on shard memory {a:.4f} ms (host {ha:.4f}) | sharded back-to-back {b:.4f} ms (host {hb:.4f}) | "
      f"sharded one at a time median {per[len(per) // 2]:.4f} min {per[0]:.4f} | kernel {zb.lib().zb_last_kernel().decode()}", flush=True)
comm.destroy()
if world > 1:
    dist.destroy_process_group()
