#!/usr/bin/env python
"""bench.py -- headline benchmark of zignal_b200 (contract: see the task statement / DESIGN.md "Measurement").

Workload (BASELINE.json configs[1]): 15x15 separable Gaussian blur (sigma = 2.25, mirror border, the taps
`Image.gaussianBlur` computes) on an 8192 x 8192 RGBA f32 image per GPU.  A "step" is one blur of that image.
With N > 1 GPUs the job is one (N*8192) x 8192 image sharded into row blocks, one block per rank (weak
scaling); every step first exchanges the 7 halo rows with the row neighbours (one batched NCCL
send/recv pair per neighbour over NVLink), then blurs the block.

  value      : Mpixels/s, whole job, inputs resident in HBM, timed with CUDA events (max over ranks)
  e2e        : the same metric through the host-pointer C-ABI entry point (pinned host buffers; H2D of the
               input and D2H of the result inside the timed region)
  roofline   : algorithmic bytes (read once + write once = 32 B/px) / kernel time vs the measured HBM peak
  cpu_baseline: the CPU oracle port (the reference's algorithm, 4 planar Image(f32) convolutions) on rank 0

`--impl reference` times the reference's own CPU implementation of the path instead (the C++ port under
oracle/: the Zig reference cannot be built in this image) with all host threads.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

ROWS = COLS = 8192
SIGMA = 2.25
HALO = 7
METRIC = "Mpixels/sec Gaussian-blur 8K RGBA f32"
WORKLOAD = "gaussian_blur_15x15_sigma2.25_mirror_8192x8192_rgba_f32"
ALGO_BYTES_PER_PX = 32  # read 16 B + write 16 B


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi SM clocks / throttle reasons while the GPU is under load."""

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.proc = None
        self.thread = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 7:
                self.samples.append(parts)

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        loaded = [s for s in self.samples if s[6].isdigit() and int(s[6]) >= 50] or self.samples
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": len(self.samples), "samples_under_load": len(loaded)}
        try:
            out["sm_mhz"] = float(np.median([float(s[0]) for s in loaded]))
            out["sm_max_mhz"] = float(loaded[0][1])
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for i, n in enumerate(names):
                if any(s[2 + i].lower().startswith("active") for s in loaded):
                    out["reasons"].append(n)
        except Exception:
            pass
        return out


def cpu_reference_pass(zo, planes, outs, taps, threads):
    """One pass of the reference's CPU path for this workload: Image(f32).convolveSeparable on each of the 4 channel
    planes (the only form the reference accepts for float data, convolution.zig:322-436)."""
    zo.set_threads(threads)
    t0 = time.perf_counter()
    for p, o in zip(planes, outs):
        zo.conv_separable(p, taps, taps, "mirror", out=o)
    return time.perf_counter() - t0


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation (oracle port), all host threads, rank 0 only."""
    if rank != 0:
        return
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_lib as zo
    threads = zo.hw_threads()
    rng = np.random.default_rng(2)
    rows = ROWS // 4  # bounded sample: a quarter of the image height, full width (vertical-pass stride preserved)
    planes = [rng.random((rows, COLS), dtype=np.float32) for _ in range(4)]
    outs = [np.empty_like(p) for p in planes]
    taps = zo.gaussian_taps(SIGMA)
    for _ in range(max(1, min(args.warmup, 2))):
        cpu_reference_pass(zo, planes, outs, taps, threads)
    steps = max(1, min(args.steps, 10))
    t = 0.0
    for _ in range(steps):
        t += cpu_reference_pass(zo, planes, outs, taps, threads)
    ms = t / steps * 1e3
    mpx = rows * COLS / 1e6 / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": METRIC, "value": mpx, "unit": "Mpixels/s", "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{rows}x{COLS} rows crop (1/4 of the image) per step", "parallelism": f"cpu{threads}"},
        "cpu_baseline": {"value": mpx, "unit": "Mpixels/s", "cores": threads, "kind": "port",
                         "sample": f"{rows}x{COLS} RGBA f32 as 4 planar Image(f32) convolutions, {steps} passes, OpenMP row-parallel"},
        "e2e": {"value": mpx, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="zignal_b200", choices=["zignal_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import zignal_b200 as zb
    from zignal_b200 import shard

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = zb.lib()  # raises if the CUDA library is missing -- there is no fallback
    stream = torch.cuda.current_stream().cuda_stream

    taps = zb.gaussian_taps(SIGMA)
    assert taps.size == 15
    gen = torch.Generator(device=dev).manual_seed(2 + rank)
    # row block of the global image with HALO rows above and below (halo rows are filled by the exchange / the global border)
    sb = shard.RowBlock(ROWS, COLS, zb.PixFmt.RGBAF32, HALO if world > 1 else 0, dev, rank, world)
    sb.interior_tensor().copy_(torch.rand(ROWS, COLS, 4, device=dev, dtype=torch.float32, generator=gen))
    out = shard.RowBlock(ROWS, COLS, zb.PixFmt.RGBAF32, HALO if world > 1 else 0, dev, rank, world)

    def step():
        # one pass of the hot path over this rank's row block: halo exchange (N > 1) overlapped with the rows that do not
        # read a halo, then the 2 x 7 boundary rows (zb_conv_separable_rows); N = 1 is a single zb_conv_separable-equivalent call
        sb.conv_separable(out, taps, taps, zb.BorderMode.MIRROR, stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = L.zb_kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = L.zb_kernel_launch_count() - launches0
    kernel_name = L.zb_last_kernel().decode()
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps

    # kernel-only time for the roofline (same stream, events around the kernel launches alone)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    src, dst = sb.image(), out.image()
    a, d = src._zb(), dst._zb()
    fp = taps.ctypes.data_as(C.POINTER(C.c_float))
    ksteps = max(args.steps, 20)
    k0.record()
    for _ in range(ksteps):
        L.zb_conv_separable(a, d, int(zb.PixFmt.RGBAF32), fp, 15, fp, 15, int(zb.BorderMode.MIRROR), stream)
    k1.record()
    torch.cuda.synchronize()
    kernel_ms = k0.elapsed_time(k1) / ksteps
    # keep the GPU under the same load a little longer so the clock sampler sees it
    if rank == 0:
        t_end = time.time() + 1.2
        while time.time() < t_end:
            for _ in range(50):
                L.zb_conv_separable(a, d, int(zb.PixFmt.RGBAF32), fp, 15, fp, 15, int(zb.BorderMode.MIRROR), stream)
            torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the host-pointer C ABI (pinned host buffers) ------------------------------------------
    e2e = None
    if not args.no_e2e:
        nbytes = ROWS * COLS * 16
        hin, hout = C.c_void_p(), C.c_void_p()
        zb._ffi.check(L.zb_malloc_host(C.byref(hin), nbytes))
        zb._ffi.check(L.zb_malloc_host(C.byref(hout), nbytes))
        harr = np.ctypeslib.as_array(C.cast(hin, C.POINTER(C.c_float)), shape=(ROWS, COLS, 4))
        harr[:] = sb.interior_tensor().cpu().numpy()
        hi = zb.ZbImage(hin.value, ROWS, COLS, COLS)
        ho = zb.ZbImage(hout.value, ROWS, COLS, COLS)
        esteps = max(2, min(args.steps, 5))
        zb._ffi.check(L.zb_host_gaussian_blur(hi, ho, int(zb.PixFmt.RGBAF32), C.c_float(SIGMA)))  # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(esteps):
            zb._ffi.check(L.zb_host_gaussian_blur(hi, ho, int(zb.PixFmt.RGBAF32), C.c_float(SIGMA)))
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / esteps
        te = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_ms = float(te.item())
        e2e = {"value": world * ROWS * COLS / 1e6 / (e2e_ms * 1e-3), "unit": "Mpixels/s", "h2d_bytes_per_step": nbytes * world,
               "d2h_bytes_per_step": nbytes * world, "ms_per_step": e2e_ms, "steps": esteps,
               "api": "zb_host_gaussian_blur (host pointers, pinned; H2D + kernel + D2H per step)"}
        L.zb_free_host(hin)
        L.zb_free_host(hout)

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        px_per_launch = (ROWS + (2 * HALO if world > 1 else 0)) * COLS
        kernel_ms_after = kernel_ms
        if world == 1:
            # N = 1: the timed region IS `steps` launches of this one kernel on the launching stream, so its average launch
            # duration is the step time itself (the separate post-loop measurement is kept as `kernel_ms_after_timed_region`)
            kernel_ms = ms_per_step
        achieved = ALGO_BYTES_PER_PX * px_per_launch / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tp = ROOT / "profiles" / "traffic.json"
        if tp.exists():
            try:
                traffic = json.loads(tp.read_text()).get("fused_sep_rgbaf32_dram_bytes_per_launch")
            except Exception:
                traffic = None
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            sys.path.insert(0, str(ROOT / "tests"))
            import oracle_lib as zo
            x = sb.interior_tensor().cpu().numpy()
            planes = [np.ascontiguousarray(x[..., k]) for k in range(4)]
            outs = [np.empty_like(p) for p in planes]
            ztaps = zo.gaussian_taps(SIGMA)
            best = min(cpu_reference_pass(zo, planes, outs, ztaps, 1) for _ in range(2))
            cpu = {"value": ROWS * COLS / 1e6 / best, "unit": "Mpixels/s", "cores": 1, "kind": "port",
                   "sample": "full 8192x8192 RGBA f32 as 4 planar Image(f32) separable convolutions, best of 2, 1 thread "
                             "(the reference is single-threaded); Zig reference not buildable here -> C++ port under oracle/"}
        line = {
            "metric": METRIC, "value": world * ROWS * COLS / 1e6 / (ms_per_step * 1e-3), "unit": "Mpixels/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "image": [ROWS * world, COLS], "pixel": "RGBA f32 interleaved (16 B)", "taps": 15, "border": "mirror",
                       "parallelism": f"row-block x{world}" + (" + NCCL halo send/recv (7 rows)" if world > 1 else ""),
                       "l2": "input 1 GiB per GPU >> 126 MB L2 (no flush needed)", "kernel": kernel_name},
            "e2e": e2e, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "kernel_ms": kernel_ms, "kernel_ms_after_timed_region": kernel_ms_after,
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_PX * px_per_launch},
            "cpu_baseline": cpu, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
