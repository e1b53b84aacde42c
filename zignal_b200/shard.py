"""Row-block sharding of one large image across the GPUs of a box (one process per GPU).

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

    def conv_separable(self, out: "RowBlock", kx, ky, border: BorderMode = BorderMode.MIRROR, stream=None):
        """Image.convolveSeparable on the global image this block belongs to (interior rows of `out` receive the result).
        One exchange (NCCL isend/irecv of `halo` rows per neighbour) overlapped with the convolution of every row that
        does not read a halo; the 2 x half boundary rows follow once the halos have landed."""
        from . import _ffi
        from .image import _fptr, current_stream
        import ctypes as C
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
        st = stream if stream is not None else current_stream()
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
