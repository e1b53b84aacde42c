"""torchrun --nproc-per-node N tools/gpu_shard_check.py [--quick]

Multi-GPU parity through the C ABI's zb_shard_* entry points (one process per GPU):
  * convolution of a row-sharded image must reproduce the single-GPU result of the whole image BIT FOR BIT -- every border mode,
    RGBA f32 (the fused kernel that fetches the neighbours' rows over NVLink in its prologue) and u8 formats (halo pull kernel + the ordinary
    kernels), ragged block heights, a ping-pong chain of steps (the flags must order reads against the neighbours' next writes),
    and again with the NCCL send/recv exchange forced;
  * box blur through zb_shard_halo_exchange + zb_box_blur on the extended view;
  * fdm.update on a row-sharded image (11 moments all-gathered over peer memory inside the statistics kernel) == single GPU;
  * batch split of rotate (C4) == the same frames of the single-GPU batch.
Prints one line per group and "shard check: OK" on rank 0; exit code 1 on any mismatch.
"""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import zignal_b200 as zb  # noqa: E402
from zignal_b200 import shard  # noqa: E402
from zignal_b200.fdm import FeatureDistributionMatching  # noqa: E402

quick = "--quick" in sys.argv
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
comm = shard.ShardComm.from_torch_distributed()
L = zb.lib()
ok = True


def report(name, same):
    global ok
    t = torch.tensor([1 if same else 0], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    good = int(t.item()) == 1
    ok &= good
    if rank == 0:
        print(f"  {name}: {'ok' if good else 'MISMATCH'}", flush=True)


def block_rows(total_like, ragged):
    """Block heights per rank: equal, or ragged (multiples of 8 for the fused path are NOT required of the ragged case)."""
    if not ragged:
        return [total_like] * world
    return [total_like + 8 * ((r * 3) % 4) for r in range(world)]


def make_full(pixfmt, rows_total, cols, seed):
    g = torch.Generator(device=dev).manual_seed(seed)   # same seed everywhere: every rank can build the global image
    ch = {zb.PixFmt.U8: 1, zb.PixFmt.RGB8: 3, zb.PixFmt.RGBA8: 4, zb.PixFmt.RGBAF32: 4, zb.PixFmt.F32: 1}[pixfmt]
    shape = (rows_total, cols) + ((ch,) if ch > 1 else ())
    if pixfmt in (zb.PixFmt.RGBAF32, zb.PixFmt.F32):
        return torch.rand(shape, device=dev, generator=g)
    return torch.randint(0, 256, shape, device=dev, dtype=torch.uint8, generator=g)


if rank == 0:
    print(f"world {world}, peer access {comm.peer_access}", flush=True)

taps = zb.gaussian_taps(2.25)
halo = 8
for path_name, path in (("auto", 0), ("nccl-exchange", 1), ("pull-kernel", 2)):
    if path != 0 and (world == 1 or (path == 2 and not comm.peer_access)):
        continue
    L.zb_shard_tune_path(path)
    for pixfmt in ((zb.PixFmt.RGBAF32, zb.PixFmt.RGBA8, zb.PixFmt.RGB8, zb.PixFmt.U8) if not quick else (zb.PixFmt.RGBAF32, zb.PixFmt.RGBA8)):
        for ragged in (False, True):
            if path == 1 and ragged and quick:
                continue
            rows_r = block_rows(256, ragged)
            cols = 1000 if pixfmt != zb.PixFmt.RGBAF32 else 1016
            full = make_full(pixfmt, sum(rows_r), cols, 11)
            r0 = sum(rows_r[:rank])
            a = comm.image(rows_r[rank], cols, pixfmt, halo)
            b = comm.image(rows_r[rank], cols, pixfmt, halo)
            same = True
            for border in (zb.BorderMode.MIRROR, zb.BorderMode.ZERO, zb.BorderMode.REPLICATE, zb.BorderMode.WRAP):
                # three chained steps a -> b -> a -> b: the single-GPU chain on the whole image is the reference
                ref = zb.Image.from_tensor(full)
                for _ in range(3):
                    ref = ref.convolve_separable(taps, taps, border)
                a.interior_tensor().copy_(full[r0:r0 + rows_r[rank]])
                src, dst = a, b
                for _ in range(3):
                    src.conv_separable(dst, taps, taps, border)
                    src, dst = dst, src
                torch.cuda.synchronize()
                comm.status()
                same &= bool(torch.equal(src.interior_tensor(), ref.tensor()[r0:r0 + rows_r[rank]]))
            report(f"conv {path_name:13s} {pixfmt.name:8s} {'ragged' if ragged else 'equal '} kernel={L.zb_last_kernel().decode()}", same)
            a.free()
            b.free()
L.zb_shard_tune_path(0)

# exact mode of the fused sharded kernel
L.zb_set_exact_f32(1)
full = make_full(zb.PixFmt.RGBAF32, 256 * world, 1016, 5)
a, b = comm.image(256, 1016, zb.PixFmt.RGBAF32, 8), comm.image(256, 1016, zb.PixFmt.RGBAF32, 8)
a.interior_tensor().copy_(full[256 * rank:256 * (rank + 1)])
want = zb.Image.from_tensor(full).convolve_separable(taps, taps, zb.BorderMode.MIRROR).tensor()[256 * rank:256 * (rank + 1)]
a.conv_separable(b, taps, taps, zb.BorderMode.MIRROR)
torch.cuda.synchronize()
report(f"conv exact-mode RGBAF32 kernel={L.zb_last_kernel().decode()}", bool(torch.equal(b.interior_tensor(), want)))
L.zb_set_exact_f32(0)
a.free()
b.free()

# other neighbourhood filters through zb_shard_halo_exchange + the ordinary entry point on the extended view.
# (box blur / sharpen evaluate an f32 summed-area table whose column prefix starts at the GLOBAL row 0 -- integral.zig:41-78 -- so
# a row block reproduces the single-GPU bits only while every prefix stays below 2^24, where f32 sums are exact: 48 columns here.)
cases = [("box blur r=3 (exact-sum regime)", 48, 4, lambda s_, d_: s_.box_blur(3, out=d_), lambda im: im.box_blur(3)),
         ("dense 3x3 convolve", 640, 1, lambda s_, d_: s_.convolve(np.array([[1, 2, 1], [2, 4, 2], [1, 2, 1]], np.float32) / 16, zb.BorderMode.MIRROR, out=d_),
          lambda im: im.convolve(np.array([[1, 2, 1], [2, 4, 2], [1, 2, 1]], np.float32) / 16, zb.BorderMode.MIRROR)),
         ("median blur r=2", 640, 2, lambda s_, d_: s_.median_blur(2, out=d_), lambda im: im.median_blur(2))]
for name, cols, reach, fn, whole_fn in cases:
    for pixfmt in (zb.PixFmt.RGBA8, zb.PixFmt.U8):
        full = make_full(pixfmt, 200 * world, cols, 3)
        a, b = comm.image(200, cols, pixfmt, 8), comm.image(200, cols, pixfmt, 8)
        a.interior_tensor().copy_(full[200 * rank:200 * (rank + 1)])
        a.apply_neighbourhood(b, fn, reach, zb.BorderMode.MIRROR)
        want = whole_fn(zb.Image.from_tensor(full)).tensor()[200 * rank:200 * (rank + 1)]
        torch.cuda.synchronize()
        report(f"{name} {pixfmt.name}", bool(torch.equal(b.interior_tensor(), want)))
        a.free()
        b.free()

# fdm on a row-sharded Rgb image
rows = 512
g = torch.Generator(device=dev).manual_seed(5)
src_full = torch.randint(0, 256, (rows * world, 1024, 3), device=dev, dtype=torch.uint8, generator=g)
base = torch.randint(0, 256, (rows * world, 1024, 3), device=dev, dtype=torch.uint8, generator=g).to(torch.float32)
mix = torch.tensor([[0.5, 0.2, 0.0], [0.1, 0.4, 0.1], [0.0, 0.2, 0.45]], device=dev)
tgt_full = (base @ mix.T * 0.6 + 50.0).clamp(0, 255).to(torch.uint8)
one = FeatureDistributionMatching(zb.PixFmt.RGB8)
whole = zb.Image.from_tensor(src_full.clone())
one.match(whole, zb.Image.from_tensor(tgt_full))
one.status()
want = whole.tensor()[rows * rank:rows * (rank + 1)]
f = FeatureDistributionMatching(zb.PixFmt.RGB8)
mine = zb.Image.from_tensor(src_full[rows * rank:rows * (rank + 1)].clone())
comm.fdm_set_target(f, zb.Image.from_tensor(tgt_full[rows * rank:rows * (rank + 1)].contiguous()))
f.set_source(mine)
for _ in range(1):
    comm.fdm_update(f)
f.status()
comm.status()
report("fdm update (sharded moments all-gather)", bool(torch.equal(mine.tensor(), want)))
f.deinit()
one.deinit()

# C4: batch split
n, R, Cc = 16, 270, 480
gen = torch.Generator(device=dev).manual_seed(4)
x = torch.randint(0, 256, (n, R, Cc, 4), device=dev, dtype=torch.uint8, generator=gen)
angle = np.float32(np.pi / 4)
cs = (np.float32(np.cos(np.float64(angle))), np.float32(np.sin(np.float64(angle))))
orows, ocols = zb.Image.from_tensor(x[0]).rotate_bounds(angle)


def rotate_batch(xs):
    y = torch.empty((xs.shape[0], orows, ocols, 4), device=dev, dtype=torch.uint8)
    if xs.shape[0]:
        zb._ffi.check(L.zb_rotate_into_batch(zb.Image.from_tensor(xs[0])._zb(), R * Cc, zb.Image.from_tensor(y[0])._zb(), orows * ocols, xs.shape[0],
                                             int(zb.PixFmt.RGBA8), C.c_float(angle), C.c_float(cs[0]), C.c_float(cs[1]), int(zb.Interpolation.BILINEAR),
                                             C.c_float(1 / 3), C.c_float(1 / 3), int(zb.BorderMode.ZERO), zb.image.current_stream()))
    return y


lo, hi = comm.split(n)
report("rotate batch split", bool(torch.equal(rotate_batch(x[lo:hi].contiguous()), rotate_batch(x)[lo:hi])))

if rank == 0:
    print("shard check:", "OK" if ok else "FAILED", f"(world {world})", flush=True)
comm.destroy()
if world > 1:
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
