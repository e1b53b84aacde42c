"""Row-block sharding of one large image across the GPUs of a box (one process per GPU).

Two layers live here:
  * `ShardComm` / `ShardImage` -- the ctypes mirror of the C ABI's multi-GPU entry points (zb_shard_*, include/zignal_b200.h):
    NCCL bootstraps, NVLink peer memory carries the data (the RGBA f32 convolution kernel TMA-loads its neighbours' edge rows
    itself, everything else goes through one pull kernel), fdm's 11 moments are all-gathered inside the statistics kernel.
    This is the product path; bench.py and tools/gpu_shard_check.py run it under torchrun.
  * `RowBlock` -- the same partitioning written against torch.distributed (gloo on CPU tensors, NCCL on CUDA tensors).  It is
    the host-logic model the world-size-2/3 gloo tests exercise without a GPU, and the fallback when a host cannot use the
    C ABI's communicator.

The reference is single-process; this is the multi-GPU layer SURVEY.md 8(e) defines for the hot path:
  * convolution / blur: contiguous row blocks, each stored with `halo` extra rows above and below; one
    batched send/recv pair per row neighbour fills the halos (NCCL over NVLink on GPUs, gloo on CPU for
    the host-logic tests); the global top/bottom edges are filled locally per the BorderMode
    (reference border.zig:46-63).  The op then runs on the extended block and the interior rows are the
    result -- the outer `halo` output rows are discarded;
  * resize / rotate / warp: no exchange (output row blocks, or a batch split);
  * fdm / pca statistics: one all-reduce of the 11 exact integer moments (`allreduce_moments`).
"""
from __future__ import annotations

import numpy as np

from .image import BorderMode, Image, PixFmt, _CH, _NP


def _dist():
    import torch.distributed as dist
    return dist


class RowBlock:
    """Rows [rank*rows, (rank+1)*rows) of a (world*rows) x cols image, with `halo` rows of padding on both sides."""

    def __init__(self, rows: int, cols: int, pixfmt: PixFmt, halo: int, device, rank: int, world: int):
        import torch
        self.rows, self.cols, self.halo = int(rows), int(cols), int(halo)
        self.pixfmt = PixFmt(pixfmt)
        self.rank, self.world = int(rank), int(world)
        ch = _CH[self.pixfmt]
        dt = torch.uint8 if _NP[self.pixfmt] == np.uint8 else torch.float32
        shape = (self.rows + 2 * self.halo, self.cols) + ((ch,) if ch > 1 else ())
        self.t = torch.zeros(shape, dtype=dt, device=device)

    # -- views -------------------------------------------------------------------------------------
    def extended_tensor(self):
        return self.t

    def interior_tensor(self):
        return self.t[self.halo:self.halo + self.rows]

    def image(self) -> Image:
        """The extended block (interior + halos) as an Image for the C ABI."""
        return Image(self.t.reshape(-1), self.pixfmt, self.rows + 2 * self.halo, self.cols, self.cols)

    def interior_image(self) -> Image:
        return Image(self.t.reshape(-1), self.pixfmt, self.rows, self.cols, self.cols, self.halo * self.cols)

    # -- overlapped convolution: exchange in flight while the halo-independent rows are computed -------
    def neighbours(self, border: BorderMode):
        """(up, down) ranks of this block, None at a global image edge (wrap closes the ring)."""
        wrap = BorderMode(border) == BorderMode.WRAP
        up = self.rank - 1 if self.rank > 0 else (self.world - 1 if wrap else None)
        down = self.rank + 1 if self.rank < self.world - 1 else (0 if wrap else None)
        return up, down

    def conv_plan(self, half: int, border: BorderMode):
        """Row bookkeeping of the overlapped schedule.  Returns (view_lo, view_hi, steps): the op runs on the VIEW
        rows [view_lo, view_hi) of the extended block -- a halo that lies beyond a global image edge is left out of the
        view, so the kernel's own border handling produces the true edge (no halo fill needed) -- and `steps` is a list of
        (row_begin, row_end, needs_halo) output windows in VIEW coordinates: first the rows that read no halo row,
        then the (at most two) strips of `half` rows next to a neighbour."""
        h, n = self.halo, self.rows
        if self.world == 1 and (BorderMode(border) != BorderMode.WRAP or h == 0):
            return h, n + h, [(0, n, False)]           # the whole image is local: one plain call on the interior
        assert h >= half, "halo must cover the kernel half-width"
        if self.world > 1 or BorderMode(border) == BorderMode.WRAP:
            up, down = self.neighbours(border) if self.world > 1 else (0, 0)
        else:
            up, down = None, None
        lo = 0 if up is not None else h
        hi = n + 2 * h if down is not None else n + h
        first, last = h - lo, h - lo + n              # interior rows in view coordinates
        a = first + (half if up is not None else 0)
        b = last - (half if down is not None else 0)
        steps = []
        if b > a:
            steps.append((a, b, False))
        else:                                          # block shorter than the kernel: everything waits for the halo
            a = b = first
        if up is not None and a > first:
            steps.append((first, a, True))
        if down is not None and last > max(b, first):
            steps.append((max(b, a), last, True))
        if not steps:
            steps.append((first, last, up is not None or down is not None))
        return lo, hi, steps

    def post_halo_exchange(self, border: BorderMode = BorderMode.MIRROR):
        """Post the neighbour sends / receives and return the requests (empty without neighbours)."""
        h, n = self.halo, self.rows
        if h == 0 or self.world == 1:
            return []
        assert n > h, "row block must be taller than the halo"
        t = self.t
        dist = _dist()
        up, down = self.neighbours(border)
        ops = []
        if up is not None:
            ops.append(dist.P2POp(dist.isend, t[h:2 * h], up))
        if down is not None:
            ops.append(dist.P2POp(dist.irecv, t[n + h:n + 2 * h], down))
            ops.append(dist.P2POp(dist.isend, t[n:n + h], down))
        if up is not None:
            ops.append(dist.P2POp(dist.irecv, t[0:h], up))
        return dist.batch_isend_irecv(ops) if ops else []

    def conv_separable(self, out: "RowBlock", kx, ky, border: BorderMode = BorderMode.MIRROR):
        """Image.convolveSeparable on the global image this block belongs to (interior rows of `out` receive the result).
        One exchange (NCCL isend/irecv of `halo` rows per neighbour) overlapped with the convolution of every row that
        does not read a halo; the 2 x half boundary rows follow once the halos have landed."""
        from . import _ffi
        from .image import _fptr
        import ctypes as C
        import torch
        assert out.halo == self.halo and out.rows == self.rows and out.cols == self.cols and out.pixfmt == self.pixfmt
        kx = np.ascontiguousarray(kx, dtype=np.float32)
        ky = np.ascontiguousarray(ky, dtype=np.float32)
        half = max(kx.size, ky.size) // 2
        if self.world == 1 and BorderMode(border) == BorderMode.WRAP and self.halo:
            self.exchange_halo(border)                 # single rank, wrap: local copy into the halos
        lo, hi, steps = self.conv_plan(half, border)
        cols = self.cols
        src = Image(self.t.reshape(-1), self.pixfmt, hi - lo, cols, cols, lo * cols)
        dst = Image(out.t.reshape(-1), out.pixfmt, hi - lo, cols, cols, lo * cols)
        a, d = src._zb(), dst._zb()
        L = _ffi.lib()
        # torch.distributed orders the exchange against the CURRENT stream of the block's device (batch_isend_irecv synchronises
        # with it when posted, wait() blocks it), so the kernels must be launched on exactly that stream
        st = torch.cuda.current_stream(self.t.device).cuda_stream
        reqs = self.post_halo_exchange(border)
        waited = False
        for (r0, r1, needs_halo) in steps:
            if needs_halo and not waited:
                for r in reqs:
                    r.wait()                           # the compute stream waits for the exchange; the host does not block
                waited = True
            _ffi.check(L.zb_conv_separable_rows(a, d, int(self.pixfmt), _fptr(kx), kx.size, _fptr(ky), ky.size, int(border),
                                                C.c_uint32(r0), C.c_uint32(r1), st))
        if not waited:
            for r in reqs:
                r.wait()
        return out

    def apply_neighbourhood(self, out: "RowBlock", fn, reach: int, border: BorderMode = BorderMode.MIRROR):
        """Any same-shape filter whose output row r reads input rows [r - reach, r + reach] -- boxBlur, sharpen, dense convolve, sobel,
        the order-statistic filters, motionBlur.linear (reach = distance / 2 + 1) -- on the global image this block belongs to: one
        neighbour exchange, then `fn(src, dst)` once on the rows this block holds.  `src` / `dst` are Image views of the two
        blocks; a halo beyond a global image edge is left out of the views, so the filter's own border handling (`border`, which
        `fn` must use) makes the true edge, exactly as in `conv_plan`.  Rows of `dst` within `reach` of a neighbour-side end of the
        view are computed from an incomplete neighbourhood, but those are halo rows of `out`: its interior rows are exact."""
        assert self.halo >= reach and out.halo == self.halo and out.rows == self.rows and out.cols == self.cols
        if self.world == 1 and BorderMode(border) == BorderMode.WRAP and self.halo:
            self.exchange_halo(border)                 # single rank, wrap: local copy into the halos
        for r in self.post_halo_exchange(border):
            r.wait()
        lo, hi, _ = self.conv_plan(reach, border)
        cols = self.cols
        src = Image(self.t.reshape(-1), self.pixfmt, hi - lo, cols, cols, lo * cols)
        dst = Image(out.t.reshape(-1), out.pixfmt, hi - lo, cols, cols, lo * cols)
        fn(src, dst)
        return out

    # -- the one exchange step of the convolution path ------------------------------------------------
    def exchange_halo(self, border: BorderMode = BorderMode.MIRROR):
        h, n = self.halo, self.rows
        if h == 0:
            return
        assert n > h, "row block must be taller than the halo"
        t = self.t
        dist = _dist()
        wrap = BorderMode(border) == BorderMode.WRAP
        up = self.rank - 1 if self.rank > 0 else (self.world - 1 if wrap else None)
        down = self.rank + 1 if self.rank < self.world - 1 else (0 if wrap else None)
        if self.world > 1:
            # Order matters when up == down (wrap with 2 ranks): the k-th send to a peer pairs with that peer's k-th recv,
            # so every rank posts (send up, recv from down) first and (send down, recv from up) second.
            ops = []
            if up is not None:
                ops.append(dist.P2POp(dist.isend, t[h:2 * h], up))           # my first rows -> upper neighbour's bottom halo
            if down is not None:
                ops.append(dist.P2POp(dist.irecv, t[n + h:n + 2 * h], down))
                ops.append(dist.P2POp(dist.isend, t[n:n + h], down))         # my last rows -> lower neighbour's top halo
            if up is not None:
                ops.append(dist.P2POp(dist.irecv, t[0:h], up))
            if ops:
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
        elif wrap:
            t[0:h] = t[n:n + h].clone()
            t[n + h:] = t[h:2 * h].clone()
        # global image edges: resolveIndex(-j) / resolveIndex(rows_total - 1 + j), j = 1..h
        b = BorderMode(border)
        if up is None:
            if b == BorderMode.ZERO:
                t[0:h].zero_()
            elif b == BorderMode.REPLICATE:
                t[0:h] = t[h:h + 1]
            elif b == BorderMode.MIRROR:      # row -j <- row j
                t[0:h] = t[h + 1:2 * h + 1].flip(0)
        if down is None:
            if b == BorderMode.ZERO:
                t[n + h:].zero_()
            elif b == BorderMode.REPLICATE:
                t[n + h:] = t[n + h - 1:n + h]
            elif b == BorderMode.MIRROR:      # row (R-1)+j <- row (R-1)-j
                t[n + h:] = t[n - 1:n + h - 1].flip(0)


def allreduce_moments(sums: np.ndarray, device=None) -> np.ndarray:
    """Sum the 11 exact integer moments (zb_fdm_moments) over all ranks: the single collective of the fdm path."""
    import torch
    dist = _dist()
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(sums, dtype=np.uint64)
    t = torch.from_numpy(np.asarray(sums, dtype=np.uint64).astype(np.int64))
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy().astype(np.uint64)


def split_batch(n_items: int, rank: int, world: int):
    """Contiguous share [lo, hi) of a batch for this rank (rotate / resize batches: no exchange)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sharded_covariance(x_local, gram_fn=None, center_fn=None):
    """PCA's scaled covariance X^T X / (n - 1) (reference pca.zig:135-154,331-338) with the SAMPLES split across ranks
    (SURVEY 8(e)): one all-reduce of the column sums (dim f64 values + the count) gives the global mean, every rank centres
    its rows and contracts them on its own GPU (the tcgen05 X^T X kernel), and one all-reduce of the dim x dim partial
    products finishes the job.  Returns (mean, cov) as tensors on x_local's device, identical on every rank.

    gram_fn(centered) -> centered^T centered and center_fn(x, mean) -> x - mean default to the CUDA library; the host-logic
    tests pass numpy-backed stand-ins."""
    import torch
    dist = _dist()
    n_local, dim = x_local.shape
    world = dist.get_world_size() if dist.is_initialized() else 1
    head = torch.empty(dim + 1, dtype=torch.float64, device=x_local.device)
    head[:dim] = x_local.sum(dim=0, dtype=torch.float64)
    head[dim] = float(n_local)
    if world > 1:
        dist.all_reduce(head, op=dist.ReduceOp.SUM)
    n = int(round(float(head[dim].item())))
    if n < 2:
        from ._ffi import ZignalError
        raise ZignalError(12, "InsufficientData")
    mean = (head[:dim] / n).to(x_local.dtype)
    if center_fn is None:
        from . import matrix

        def center_fn(x, m):
            out = torch.empty_like(x)
            matrix.center_columns(x, m.contiguous(), False, out)
            return out
    if gram_fn is None:
        from . import matrix

        def gram_fn(c):
            return matrix.gemm_device(c, c, True, False, 1.0, 0.0, None)
    part = gram_fn(center_fn(x_local, mean)).to(torch.float64)
    if world > 1:
        dist.all_reduce(part, op=dist.ReduceOp.SUM)
    return mean, (part / float(n - 1)).to(x_local.dtype)


# =====================================================================================================================
# The C ABI's multi-GPU layer (zb_shard_*): NCCL bootstrap + NVLink peer memory
# =====================================================================================================================
class _DevMem:
    """A raw device allocation exposed through __cuda_array_interface__ so torch can wrap it without owning it."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class ShardComm:
    """zb_shard_comm: one per process / GPU.  `ShardComm.from_torch_distributed()` takes rank / world from the initialised
    process group and broadcasts the NCCL unique id through it (any backend); a host without torch passes the id itself."""

    def __init__(self, rank: int, world: int, unique_id: bytes = None):
        import ctypes as C
        from . import _ffi
        L = _ffi.lib()
        self._h = C.c_void_p()
        idbuf = (C.c_uint8 * 128).from_buffer_copy(unique_id) if unique_id is not None else None
        _ffi.check(L.zb_shard_comm_create(C.byref(self._h), int(rank), int(world), idbuf))
        self.rank, self.world = int(rank), int(world)
        pa = C.c_int(0)
        _ffi.check(L.zb_shard_comm_info(self._h, None, None, C.byref(pa)))
        self.peer_access = bool(pa.value)

    @staticmethod
    def unique_id() -> bytes:
        import ctypes as C
        from . import _ffi
        buf = (C.c_uint8 * 128)()
        _ffi.check(_ffi.lib().zb_shard_unique_id(buf))
        return bytes(buf)

    @classmethod
    def from_torch_distributed(cls) -> "ShardComm":
        dist = _dist()
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return cls(0, 1, None)
        rank, world = dist.get_rank(), dist.get_world_size()
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return cls(rank, world, box[0])

    def destroy(self):
        from . import _ffi
        if self._h:
            _ffi.lib().zb_shard_comm_destroy(self._h)
            self._h = None

    def status(self):
        """Waits for the current stream; raises DeviceFailure if a kernel gave up waiting for a neighbour."""
        from . import _ffi
        from .image import current_stream
        _ffi.check(_ffi.lib().zb_shard_status(self._h, current_stream()))

    def split(self, n_items: int):
        import ctypes as C
        from . import _ffi
        lo, hi = C.c_uint32(), C.c_uint32()
        _ffi.check(_ffi.lib().zb_shard_split(int(n_items), self.rank, self.world, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def allreduce(self, t):
        """In-place sum of a contiguous CUDA tensor (f32 / f64 / i64-as-u64) over all ranks (zb_shard_allreduce, NCCL)."""
        import torch
        from . import _ffi
        dt = {torch.float32: 0, torch.float64: 1, torch.int64: 2}[t.dtype]
        assert t.is_cuda and t.is_contiguous()
        _ffi.check(_ffi.lib().zb_shard_allreduce(self._h, t.data_ptr(), t.numel(), dt, torch.cuda.current_stream(t.device).cuda_stream))
        return t

    def image(self, rows: int, cols: int, pixfmt: PixFmt, halo: int = 0) -> "ShardImage":
        return ShardImage(self, rows, cols, pixfmt, halo)

    # ---- fdm on a row-sharded image ----
    def fdm_set_target(self, f, target_block: Image):
        from . import _ffi
        from .image import current_stream
        t = target_block._zb()
        _ffi.check(_ffi.lib().zb_shard_fdm_set_target(self._h, f._h, t, current_stream()))

    def fdm_update(self, f):
        from . import _ffi
        from .image import current_stream
        _ffi.check(_ffi.lib().zb_shard_fdm_update(self._h, f._h, current_stream()))


class ShardImage:
    """This rank's row block (rows x cols, `halo` spare rows above and below) of a global image of world blocks stacked in rank
    order: symmetric memory from zb_shard_alloc, described to the library by zb_shard_image_create."""

    def __init__(self, comm: ShardComm, rows: int, cols: int, pixfmt: PixFmt, halo: int = 0):
        import ctypes as C
        import torch
        from . import _ffi
        from ._ffi import ZbImage
        L = _ffi.lib()
        self.comm, self.rows, self.cols, self.halo = comm, int(rows), int(cols), int(halo)
        self.pixfmt = PixFmt(pixfmt)
        ch = _CH[self.pixfmt]
        self._np = _NP[self.pixfmt]
        esz = np.dtype(self._np).itemsize
        self._pb = ch * esz
        total_rows = self.rows + 2 * self.halo
        nbytes = max(16, total_rows * self.cols * self._pb)
        base = C.c_void_p()
        _ffi.check(L.zb_shard_alloc(comm._h, nbytes, C.byref(base)))
        self._base = base.value
        shape = (total_rows, self.cols) + ((ch,) if ch > 1 else ())
        self._mem = _DevMem(self._base, shape, "|u1" if self._np == np.uint8 else "<f4")
        self.t = torch.as_tensor(self._mem, device=torch.device("cuda", torch.cuda.current_device()))
        self.t.zero_()
        torch.cuda.synchronize()
        blk = ZbImage(self._base + self.halo * self.cols * self._pb, self.rows, self.cols, self.cols)
        self._h = C.c_void_p()
        _ffi.check(L.zb_shard_image_create(comm._h, blk, self.halo, int(self.pixfmt), C.byref(self._h)))

    def interior_tensor(self):
        return self.t[self.halo:self.halo + self.rows]

    def block_image(self) -> Image:
        return Image(self.t.reshape(-1), self.pixfmt, self.rows, self.cols, self.cols, self.halo * self.cols)

    def free(self):
        from . import _ffi
        L = _ffi.lib()
        if self._h:
            L.zb_shard_image_destroy(self._h)
            self._h = None
        if self._base:
            self.t = None
            _ffi.check(L.zb_shard_free(self.comm._h, self._base))
            self._base = None

    def _stream(self):
        import torch
        return torch.cuda.current_stream(self.t.device).cuda_stream

    def conv_separable(self, out: "ShardImage", kx, ky, border: BorderMode = BorderMode.MIRROR) -> "ShardImage":
        from . import _ffi
        from .image import _fptr
        kx = np.ascontiguousarray(kx, dtype=np.float32)
        ky = np.ascontiguousarray(ky, dtype=np.float32)
        _ffi.check(_ffi.lib().zb_shard_conv_separable(self.comm._h, self._h, out._h, _fptr(kx), kx.size, _fptr(ky), ky.size, int(border),
                                                      self._stream()))
        return out

    def gaussian_blur(self, out: "ShardImage", sigma: float) -> "ShardImage":
        import ctypes as C
        from . import _ffi
        _ffi.check(_ffi.lib().zb_shard_gaussian_blur(self.comm._h, self._h, out._h, C.c_float(sigma), self._stream()))
        return out

    def halo_exchange(self, reach: int, border: BorderMode = BorderMode.MIRROR):
        from . import _ffi
        _ffi.check(_ffi.lib().zb_shard_halo_exchange(self.comm._h, self._h, int(reach), int(border), self._stream()))

    def view(self, reach: int, border: BorderMode = BorderMode.MIRROR):
        """(Image of the block plus the exchanged halo rows, index of the block's first row inside it)."""
        import ctypes as C
        from . import _ffi
        from ._ffi import ZbImage
        v = ZbImage()
        first = C.c_uint32()
        _ffi.check(_ffi.lib().zb_shard_view(self._h, int(reach), int(border), C.byref(v), C.byref(first)))
        off_rows = self.halo - first.value
        return Image(self.t.reshape(-1), self.pixfmt, v.rows, self.cols, self.cols, off_rows * self.cols), first.value

    def apply_neighbourhood(self, out: "ShardImage", fn, reach: int, border: BorderMode = BorderMode.MIRROR) -> "ShardImage":
        """Any same-shape filter whose output row r reads input rows [r - reach, r + reach] (boxBlur, sharpen, dense convolve, sobel,
        order statistics, motionBlur.linear) on the global image: one halo exchange, then `fn(src_view, dst_view)` once."""
        assert out.halo == self.halo and out.rows == self.rows and out.cols == self.cols and self.halo >= reach
        self.halo_exchange(reach, border)
        sv, first = self.view(reach, border)
        dv, first2 = out.view(reach, border)
        assert first == first2
        fn(sv, dv)
        return out
