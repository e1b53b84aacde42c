"""torchrun --nproc-per-node N tools/gpu_shard_check.py : the overlapped row-block convolution over NCCL must reproduce the
single-GPU result of the whole image bit for bit (every border mode, f32 and u8)."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import zignal_b200 as zb  # noqa: E402
from zignal_b200 import shard  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rows, cols, halo = 512, 1000, 8
taps = zb.gaussian_taps(2.25)
ok = True
for pixfmt, dtype in ((zb.PixFmt.RGBAF32, torch.float32), (zb.PixFmt.RGBA8, torch.uint8)):
    g = torch.Generator(device=dev).manual_seed(11)     # same seed everywhere: every rank can build the global image
    if dtype == torch.float32:
        full = torch.rand(rows * world, cols, 4, device=dev, generator=g)
    else:
        full = torch.randint(0, 256, (rows * world, cols, 4), device=dev, dtype=torch.uint8, generator=g)
    for border in (zb.BorderMode.MIRROR, zb.BorderMode.ZERO, zb.BorderMode.REPLICATE, zb.BorderMode.WRAP):
        want = zb.Image.from_tensor(full).convolve_separable(taps, taps, border).tensor()[rank * rows:(rank + 1) * rows]
        sb = shard.RowBlock(rows, cols, pixfmt, halo, dev, rank, world)
        ob = shard.RowBlock(rows, cols, pixfmt, halo, dev, rank, world)
        sb.interior_tensor().copy_(full[rank * rows:(rank + 1) * rows])
        for _ in range(3):
            sb.conv_separable(ob, taps, taps, border)
        torch.cuda.synchronize()
        same = bool(torch.equal(ob.interior_tensor(), want))
        ok &= same
        if not same:
            print(f"rank {rank}: MISMATCH {pixfmt.name} {border.name}", flush=True)
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("shard check:", "OK" if int(t.item()) == 1 else "FAILED", f"(world {world})", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
