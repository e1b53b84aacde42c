"""world_size-2 (and 3) gloo tests of the multi-GPU host logic on CPU: the halo exchange + global-border fill
must let every rank reproduce its rows of the single-image result, and the moment all-reduce must equal
the global moments.  The local operator here is the CPU oracle (the CUDA kernels cannot run without a GPU);
what is under test is zignal_b200.shard."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as zo


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, border, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from zignal_b200 import BorderMode, PixFmt, shard
        rows, cols, halo = 24, 19, 7
        rng = np.random.default_rng(42)
        full = rng.random((rows * world, cols, 4), dtype=np.float32)
        taps = zo.gaussian_taps(2.25)
        want = zo.conv_separable(full, taps, taps, border)
        blk = shard.RowBlock(rows, cols, PixFmt.RGBAF32, halo, "cpu", rank, world)
        blk.interior_tensor().copy_(torch.from_numpy(full[rank * rows:(rank + 1) * rows]))
        blk.exchange_halo(BorderMode[border.upper()])
        ext = blk.extended_tensor().numpy()
        out_ext = zo.conv_separable(np.ascontiguousarray(ext), taps, taps, border)
        got = out_ext[halo:halo + rows]
        ok_conv = bool(np.array_equal(got, want[rank * rows:(rank + 1) * rows]))
        # fdm moments: per-shard integer sums, one all-reduce
        img = np.random.default_rng(7).integers(0, 256, (rows * world, cols, 3), dtype=np.uint8)
        part = img[rank * rows:(rank + 1) * rows].reshape(-1, 3).astype(np.int64)
        r, g, b = part[:, 0], part[:, 1], part[:, 2]
        sums = np.array([len(r), r.sum(), g.sum(), b.sum(), (r * r).sum(), (r * g).sum(), (r * b).sum(), (g * g).sum(), (g * b).sum(),
                         (b * b).sum(), np.count_nonzero((r != g) | (g != b))], dtype=np.uint64)
        tot = shard.allreduce_moments(sums)
        allp = img.reshape(-1, 3).astype(np.int64)
        ok_mom = int(tot[0]) == allp.shape[0] and int(tot[5]) == int((allp[:, 0] * allp[:, 1]).sum()) and int(tot[9]) == int((allp[:, 2] ** 2).sum())
        lo, hi = shard.split_batch(10, rank, world)
        q.put((rank, ok_conv, ok_mom, (lo, hi)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("border", ["mirror", "zero", "replicate", "wrap"])
def test_row_block_halo_exchange_reproduces_single_image_result(world, border):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, border, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert all(ok for _, ok, _, _ in res), f"conv rows differ: {res}"
    assert all(ok for _, _, ok, _ in res), f"moment all-reduce differs: {res}"
    spans = [s for *_, s in res]
    assert spans[0][0] == 0 and spans[-1][1] == 10 and all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))


def test_single_rank_wrap_and_mirror_fill():
    from zignal_b200 import BorderMode, PixFmt, shard
    rng = np.random.default_rng(1)
    full = rng.random((20, 9, 4), dtype=np.float32)
    taps = zo.gaussian_taps(1.0)
    for border in ("mirror", "wrap", "replicate", "zero"):
        blk = shard.RowBlock(20, 9, PixFmt.RGBAF32, 3, "cpu", 0, 1)
        blk.interior_tensor().copy_(torch.from_numpy(full))
        blk.exchange_halo(BorderMode[border.upper()])
        out = zo.conv_separable(np.ascontiguousarray(blk.extended_tensor().numpy()), taps, taps, border)[3:23]
        assert np.array_equal(out, zo.conv_separable(full, taps, taps, border)), border


def _plan_worker(rank, world, port, border, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from zignal_b200 import BorderMode, PixFmt, shard
        rows, cols, halo = 24, 11, 8
        rng = np.random.default_rng(5)
        full = rng.random((rows * world, cols, 4), dtype=np.float32)
        taps = zo.gaussian_taps(2.25)           # 15 taps: half = 7 <= halo
        half = taps.size // 2
        want = zo.conv_separable(full, taps, taps, border)[rank * rows:(rank + 1) * rows]
        blk = shard.RowBlock(rows, cols, PixFmt.RGBAF32, halo, "cpu", rank, world)
        blk.interior_tensor().copy_(torch.from_numpy(full[rank * rows:(rank + 1) * rows]))
        b = BorderMode[border.upper()]
        lo, hi, steps = blk.conv_plan(half, b)
        ext = blk.extended_tensor()
        # 1. windows flagged halo-free must not read a halo row: poison the halos, convolve the view, compare those rows
        ext[:halo] = float("nan")
        ext[halo + rows:] = float("nan")
        view = np.ascontiguousarray(ext.numpy()[lo:hi])
        pre = zo.conv_separable(view, taps, taps, border)
        first = halo - lo
        ok_free = all(np.array_equal(pre[r0:r1], want[r0 - first:r1 - first]) for r0, r1, needs in steps if not needs)
        # 2. after the exchange the union of all windows is exactly the interior and reproduces the global result
        for r in blk.post_halo_exchange(b):
            r.wait()
        view = np.ascontiguousarray(ext.numpy()[lo:hi])
        post = zo.conv_separable(view, taps, taps, border)
        got = np.full_like(want, np.nan)
        for r0, r1, _ in steps:
            got[r0 - first:r1 - first] = post[r0:r1]
        covered = sorted((r0, r1) for r0, r1, _ in steps)
        ok_cover = covered[0][0] == first and covered[-1][1] == first + rows and all(covered[i][1] == covered[i + 1][0] for i in range(len(covered) - 1))
        q.put((rank, bool(ok_free), bool(ok_cover), bool(np.array_equal(got, want))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("border", ["mirror", "zero", "replicate", "wrap"])
def test_overlapped_conv_plan(world, border):
    """The overlapped schedule of RowBlock.conv_separable: edge blocks leave the outer halo out of the view (the kernel's own
    border handling makes the global edge), halo-free windows really are halo-free, and the windows tile the interior."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_plan_worker, args=(r, world, port, border, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(a and b and c for _, a, b, c in res), res


def _cov_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from zignal_b200 import shard
        rng = np.random.default_rng(9)
        x = (rng.standard_normal((301, 12)) * np.arange(1, 13) + 3.0).astype(np.float32)
        lo, hi = shard.split_batch(x.shape[0], rank, world)
        mean, cov = shard.sharded_covariance(torch.from_numpy(x[lo:hi]), gram_fn=lambda c: c.double().T @ c.double(),
                                             center_fn=lambda a, m: a - m)
        xc = x.astype(np.float64) - x.astype(np.float64).mean(axis=0)
        want = xc.T @ xc / (x.shape[0] - 1)
        ok_mean = bool(np.allclose(mean.numpy(), x.astype(np.float64).mean(axis=0), rtol=1e-6, atol=1e-6))
        ok_cov = bool(np.abs(cov.numpy() - want).max() <= 1e-5 * np.abs(want).max())
        q.put((rank, ok_mean, ok_cov))
    finally:
        dist.destroy_process_group()


def test_sharded_covariance_allreduce():
    """PCA covariance with samples split over ranks: all-reduce of column sums, then of the dim x dim partial products."""
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cov_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(a and b for _, a, b in res), res


def _nbhd_worker(rank, world, port, border, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from zignal_b200 import BorderMode, PixFmt, shard
        rows, cols, halo = 11, 17, 3
        full = np.random.default_rng(5).integers(0, 256, (rows * world, cols, 4), dtype=np.uint8)
        results = []
        for name, reach, op in [("median r2", 2, lambda a: zo.order_blur(a, 2, "percentile", 0.5, border)),
                                ("max r3", 3, lambda a: zo.order_blur(a, 3, "percentile", 1.0, border)),
                                ("dense 3x3", 1, lambda a: zo.convolve(a, np.full((3, 3), 1.0 / 9.0, np.float32), border))]:
            want = op(full)
            src = shard.RowBlock(rows, cols, PixFmt.RGBA8, halo, "cpu", rank, world)
            dst = shard.RowBlock(rows, cols, PixFmt.RGBA8, halo, "cpu", rank, world)
            src.extended_tensor().fill_(0xEE)                       # poison: a halo that is used without being exchanged shows up
            src.interior_tensor().copy_(torch.from_numpy(full[rank * rows:(rank + 1) * rows]))

            def fn(s, d, op=op):                                     # the local operator is the CPU oracle on the view the block hands out
                res = op(s.to_numpy())
                d._t[d._off * 4:(d._off + d.rows * d.stride) * 4] = torch.from_numpy(np.ascontiguousarray(res).reshape(-1))
            src.apply_neighbourhood(dst, fn, reach, BorderMode[border.upper()])
            results.append(bool(np.array_equal(dst.interior_tensor().numpy(), want[rank * rows:(rank + 1) * rows])))
        q.put((rank, results))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2, 3])
@pytest.mark.parametrize("border", ["mirror", "zero", "replicate", "wrap"])
def test_sharded_neighbourhood_filters_reproduce_the_single_image_result(world, border):
    """RowBlock.apply_neighbourhood: order-statistic and dense filters on a row-sharded image equal the filter of the whole image, for
    every border mode (global edges come from the filter's own border handling on a view that leaves the outer halo out)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nbhd_worker, args=(r, world, port, border, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(all(flags) for _, flags in res), res


def test_conv_plan_invariants_for_every_rank_up_to_eight_gpus():
    """conv_plan is pure bookkeeping, so it can be checked for every (world, rank) of the 1..8-GPU scaling run without a process group:
    the windows tile the interior exactly once, windows flagged halo-free never read a neighbour's halo row, halo windows exist only next
    to a neighbour, and edge blocks leave the outer halo out of the view."""
    from zignal_b200 import BorderMode, PixFmt, shard
    rng = np.random.default_rng(11)
    for world in range(1, 9):
        for rank in range(world):
            for border in (BorderMode.ZERO, BorderMode.REPLICATE, BorderMode.MIRROR, BorderMode.WRAP):
                for _ in range(6):
                    half = int(rng.integers(0, 9))
                    halo = half + int(rng.integers(0, 3)) if (world > 1 or border == BorderMode.WRAP) else int(rng.integers(0, 10))
                    rows = int(rng.integers(max(halo + 1, 1), 40))
                    blk = shard.RowBlock(rows, 5, PixFmt.U8, halo, "cpu", rank, world)
                    lo, hi, steps = blk.conv_plan(half, border)
                    if world == 1 and (border != BorderMode.WRAP or halo == 0):
                        assert (lo, hi) == (halo, halo + rows) and steps == [(0, rows, False)]
                        continue
                    up, down = blk.neighbours(border) if world > 1 else (0, 0)
                    assert lo == (0 if up is not None else halo) and hi == (rows + 2 * halo if down is not None else rows + halo)
                    first = halo - lo
                    last = first + rows
                    covered = sorted((a, b) for a, b, _ in steps)
                    assert covered[0][0] == first and covered[-1][1] == last, (world, rank, border, rows, halo, half, steps)
                    assert all(x[1] == y[0] for x, y in zip(covered, covered[1:])) and all(b > a for a, b in covered if rows > 0)
                    for a, b, needs in steps:
                        reads_up_halo = up is not None and a - half < first
                        reads_down_halo = down is not None and b - 1 + half >= last
                        if not needs:
                            assert not reads_up_halo and not reads_down_halo, (world, rank, border, rows, halo, half, steps)
                    assert all(not needs for _, _, needs in steps) or up is not None or down is not None


def test_c_abi_batch_split_matches_the_python_model():
    """zb_shard_split (pure host arithmetic behind the C ABI) == shard.split_batch for every rank of 1..8-GPU runs."""
    import ctypes as C

    from zignal_b200 import _ffi, shard
    L = _ffi.lib()
    for world in range(1, 9):
        for n in (0, 1, 7, 128, 1024, 1000003):
            covered = 0
            for rank in range(world):
                lo, hi = C.c_uint32(), C.c_uint32()
                assert L.zb_shard_split(n, rank, world, C.byref(lo), C.byref(hi)) == 0
                assert (lo.value, hi.value) == shard.split_batch(n, rank, world)
                assert lo.value == covered
                covered = hi.value
            assert covered == n
    lo, hi = C.c_uint32(), C.c_uint32()
    assert L.zb_shard_split(4, 3, 2, C.byref(lo), C.byref(hi)) != 0          # rank outside the world


def test_shard_entry_points_reject_bad_arguments_without_a_gpu():
    import ctypes as C

    from zignal_b200 import _ffi
    L = _ffi.lib()
    h = C.c_void_p()
    assert L.zb_shard_comm_create(C.byref(h), 0, 0, None) == 5                 # world < 1: InvalidArgument
    assert L.zb_shard_comm_create(C.byref(h), 2, 2, None) == 5                 # rank outside the world
    assert L.zb_shard_comm_create(C.byref(h), 0, 2, None) == 5                 # world > 1 needs the unique id
    assert L.zb_shard_conv_separable(None, None, None, None, 0, None, 0, 2, None) == 5
    assert L.zb_shard_halo_exchange(None, None, 1, 2, None) == 5
    assert L.zb_shard_fdm_update(None, None, None) == 5
    assert L.zb_shard_tune_path(7) == 5
