"""K4 — one very long (seq_a, seq_b) pair split into column stripes (BASELINE config[4]).

The recurrence is the reference's (dtw.py:32-40); striping has no reference equivalent.
Stripe g owns columns [c0_g, c1_g) of seq_b (reference orientation dtw.py:7-8: rows = seq_a,
columns = seq_b) and all M rows.  For every band of 128 rows it needs acc_cost[i, c0_g - 1] from
stripe g-1 and hands acc_cost[i, c1_g - 1] to stripe g+1.

* ``dtw_striped_local``       — all stripes on one GPU, one after the other (tests, small GPUs).
* ``dtw_striped_distributed`` — one stripe per rank (one process per GPU).  Every rank allocates an
  exchange block (boundary column + per-band flags) with CUDA IPC; its left neighbour maps it and
  the stripe kernel stores boundary values and raises flags straight into it over NVLink
  (system-scope release/acquire) while the right neighbour's kernel is already running: compute and
  hand-off are one kernel, band by band.  The backtrack walks rank G-1 -> 0; the row at which the
  path crosses a stripe edge is passed with torch.distributed send/recv (NCCL on GPU tensors, gloo
  on CPU tensors in the tests).
"""
import ctypes as C

import numpy as np
import torch

try:
    from . import _native as nat
    from .dtw import DtwPlan
except ImportError:
    import _native as nat
    from dtw import DtwPlan

BAND_ROWS = 128


def stripe_bounds(n_cols, n_stripes):
    """Contiguous column ranges, as even as possible, every stripe non-empty."""
    n_stripes = max(1, min(int(n_stripes), int(n_cols)))
    base, extra = divmod(int(n_cols), n_stripes)
    bounds, c = [], 0
    for g in range(n_stripes):
        w = base + (1 if g < extra else 0)
        bounds.append((c, c + w))
        c += w
    return bounds


def stitch_segments(segments):
    """segments[g] = stripe g's part of the path (start-to-end order, global columns) -> full path."""
    parts = [np.asarray(s, dtype=np.int64).reshape(-1, 2) for s in segments if len(s)]
    return np.concatenate(parts, axis=0) if parts else np.empty((0, 2), dtype=np.int64)


class _Stripe(object):
    """Device objects of one stripe: single-pair fp64 plan of (M x width)."""

    def __init__(self, M, width, col0, first):
        self.M, self.width, self.col0, self.first = int(M), int(width), int(col0), bool(first)
        self.plan = DtwPlan([M], [width], dtype="fp64")
        self.out3 = torch.zeros(3, dtype=torch.int32, device=self.plan.device)

    def accumulate(self, d_a, d_b_stripe, leftb, in_flag, rightb_ptr, out_flag_ptr):
        p = self.plan
        p.acc_end.zero_()          # NaN afterwards == "a wait on the left stripe timed out" (see check())
        nat.check(nat.lib().afs_dtw_accumulate_stripe(
            p._h, nat.ptr(d_a), nat.ptr(d_b_stripe), nat.ptr(p.workspace), nat.ptr(p.acc_end),
            _vp(leftb), _vp(in_flag), _vp(rightb_ptr), _vp(out_flag_ptr), nat.stream_ptr()))

    def backtrack(self, start_i, start_j):
        """-> (segment (P,2) int64 in start-to-end order, exit_row or -1)."""
        p = self.plan
        nat.check(nat.lib().afs_dtw_backtrack_stripe(p._h, nat.ptr(p.workspace), int(start_i), int(start_j), self.col0,
                                                     1 if self.first else 0, nat.ptr(p.path), nat.ptr(self.out3),
                                                     nat.stream_ptr()))
        pos, cnt, exit_i = (int(v) for v in self.out3.cpu().numpy())
        seg = p.path[pos : pos + cnt].cpu().numpy().astype(np.int64)
        return seg, exit_i

    def close(self):
        self.plan.close()


def _vp(x):
    """tensor | raw device address | None -> c_void_p"""
    if x is None:
        return C.c_void_p(0)
    if isinstance(x, torch.Tensor):
        return C.c_void_p(x.data_ptr())
    return C.c_void_p(int(x))


def dtw_striped_local(seq_a, seq_b, n_stripes):
    """All stripes on the current GPU, sequentially.  Returns (acc_end, path (P,2) int64)."""
    nat.require_cuda()
    a = np.ascontiguousarray(seq_a, dtype=np.float64)
    b = np.ascontiguousarray(seq_b, dtype=np.float64)
    assert a.shape[0] == 12 and b.shape[0] == 12
    M, N = a.shape[1], b.shape[1]
    bounds = stripe_bounds(N, n_stripes)
    dev = nat.device()
    d_a = torch.from_numpy(a).to(dev)
    stripes, rightbs = [], []
    leftb = None
    for g, (c0, c1) in enumerate(bounds):
        st = _Stripe(M, c1 - c0, c0, g == 0)
        d_b = torch.from_numpy(np.ascontiguousarray(b[:, c0:c1])).to(dev)
        rightb = torch.empty(M, dtype=torch.float64, device=dev)
        st.accumulate(d_a, d_b, leftb, None, rightb, None)
        stripes.append(st)
        rightbs.append(rightb)
        leftb = rightb
    acc_end = float(stripes[-1].plan.acc_end.cpu()[0])
    segs = [None] * len(bounds)
    i, g = M - 1, len(bounds) - 1
    while g >= 0:
        seg, exit_i = stripes[g].backtrack(i, bounds[g][1] - bounds[g][0] - 1)
        segs[g] = seg
        if exit_i < 0:
            break
        i, g = exit_i, g - 1
    for g2 in range(g):           # stripes the path never reached (cannot happen: it ends at column 0)
        segs[g2] = np.empty((0, 2), dtype=np.int64)
    for st in stripes:
        st.close()
    return acc_end, stitch_segments([s for s in segs if s is not None])


def handoff_backtrack(rank, world, local_backtrack, last_start_row, dist, device="cpu", group=None):
    """Host protocol of the distributed backtrack.  `local_backtrack(start_row) -> (segment, exit_row)`
    runs this rank's stripe; the entry row travels right-to-left with send/recv.  Returns this rank's
    segment (possibly empty).  `rank` / `world` are relative to `group`."""
    def peer(r):
        return r if group is None else dist.get_global_rank(group, r)

    buf = torch.zeros(1, dtype=torch.int64, device=device)
    if rank == world - 1:
        start = int(last_start_row)
    else:
        dist.recv(buf, src=peer(rank + 1), group=group)
        start = int(buf.item())
    if start < 0:
        seg, exit_i = np.empty((0, 2), dtype=np.int64), -1      # the path ended in a stripe to the right
    else:
        seg, exit_i = local_backtrack(start)
    if rank > 0:
        buf[0] = exit_i
        dist.send(buf, dst=peer(rank - 1), group=group)
    return seg


def gather_segments(seg, rank, world, dist, device="cpu", group=None):
    """Path segments of all ranks -> list on rank 0 (None elsewhere), as tensors over the process group's own
    transport (NCCL send/recv on GPU tensors: a few hundred microseconds for a 200k-point path; the pickling
    gather_object it replaces took 30 ms)."""
    def peer(r):
        return r if group is None else dist.get_global_rank(group, r)

    n = torch.tensor([len(seg)], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(t.item()) for t in sizes]
    if rank == 0:
        out = [np.asarray(seg, dtype=np.int64).reshape(-1, 2)]
        for r in range(1, world):
            if sizes[r] == 0:
                out.append(np.empty((0, 2), dtype=np.int64))
                continue
            buf = torch.empty((sizes[r], 2), dtype=torch.int64, device=device)
            dist.recv(buf, src=peer(r), group=group)
            out.append(buf.cpu().numpy())
        return out
    if sizes[rank] > 0:
        dist.send(torch.from_numpy(np.ascontiguousarray(seg, dtype=np.int64)).to(device), dst=peer(0), group=group)
    return None


class StripedDtwDistributed(object):
    """Rank g of a process group aligns stripe g.  Reusable across runs of the same shapes."""

    def __init__(self, M, N, dist, group=None):
        nat.require_cuda()
        self.dist = dist
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.group = group
        self.M, self.N = int(M), int(N)
        self.bounds = stripe_bounds(N, self.world)
        assert len(self.bounds) == self.world, "more ranks than columns"
        c0, c1 = self.bounds[self.rank]
        self.stripe = _Stripe(M, c1 - c0, c0, self.rank == 0)
        self.nbands = (self.M + BAND_ROWS - 1) // BAND_ROWS
        self.flag_off = (self.M * 8 + 255) // 256 * 256
        self.inbox_bytes = self.flag_off + self.nbands * 4
        L = nat.lib()
        handle = (C.c_ubyte * 64)()
        ptr = C.c_void_p()
        nat.check(L.afs_ipc_alloc(self.inbox_bytes, C.byref(ptr), handle))
        self.inbox = ptr.value
        dev = self.stripe.plan.device
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
        allh = [torch.empty(64, dtype=torch.uint8, device=dev) for _ in range(self.world)]
        dist.all_gather(allh, mine, group=group)
        self.peer = None
        if self.rank + 1 < self.world:
            raw = bytes(allh[self.rank + 1].cpu().numpy().tolist())
            hb = (C.c_ubyte * 64).from_buffer_copy(raw)
            pp = C.c_void_p()
            nat.check(L.afs_ipc_open(hb, C.byref(pp)))
            self.peer = pp.value
        dist.barrier(group=group)
        self._armed = False          # reset() arms one accumulate(): the flags of the exchange block are a 0/1 latch

    def _clear_inbox(self):
        nat.check(nat.lib().afs_ipc_clear(C.c_void_p(self.inbox), self.inbox_bytes, nat.stream_ptr()))

    def accumulate(self, d_a, d_b_stripe):
        """Asynchronous launch of this rank's stripe (call on every rank; kernels overlap across GPUs).
        Every run needs a reset() on all ranks first: the band flags in the exchange block stay raised after a run,
        and a second run would consume the previous run's boundary column."""
        if not self._armed:
            raise nat.AfsError("StripedDtwDistributed.accumulate: call reset() on every rank before each run")
        self._armed = False
        left = self.inbox if self.rank > 0 else None
        inflag = self.inbox + self.flag_off if self.rank > 0 else None
        right = self.peer if self.peer is not None else None
        outflag = self.peer + self.flag_off if self.peer is not None else None
        self.stripe.accumulate(d_a, d_b_stripe, left, inflag, right, outflag)

    def reset(self):
        """Zero the flags of the exchange block; all ranks must have done so before anyone launches."""
        torch.cuda.synchronize()
        self._clear_inbox()
        torch.cuda.synchronize()
        self.dist.barrier(group=self.group)
        self._armed = True

    def check(self):
        """Collective: raise AfsError on EVERY rank if any rank's stripe kernel gave up waiting for its left
        neighbour (the kernel bounds that wait, so a crashed rank cannot hang the other GPUs)."""
        dev = self.stripe.plan.device
        bad = torch.isnan(self.stripe.plan.acc_end[:1]).to(torch.int32)
        self.dist.all_reduce(bad, op=self.dist.ReduceOp.MAX, group=self.group)
        if int(bad.cpu()[0]) != 0:
            raise nat.AfsError("striped DTW: a stripe timed out waiting for the stripe on its left; results are invalid")

    def backtrack(self):
        """Returns the full path on rank 0 (None elsewhere) and acc_end on the last rank."""
        self.check()
        dev = self.stripe.plan.device
        width = self.bounds[self.rank][1] - self.bounds[self.rank][0]
        seg = handoff_backtrack(self.rank, self.world, lambda i: self.stripe.backtrack(i, width - 1), self.M - 1,
                                self.dist, device=dev, group=self.group)
        gathered = gather_segments(seg, self.rank, self.world, self.dist, device=dev, group=self.group)
        if self.rank == 0:
            return stitch_segments(gathered)
        return None

    def acc_end(self):
        return float(self.stripe.plan.acc_end.cpu()[0]) if self.rank == self.world - 1 else None

    def close(self):
        L = nat.lib()
        if self.peer is not None:
            L.afs_ipc_close(C.c_void_p(self.peer))
            self.peer = None
        self.dist.barrier(group=self.group)
        if self.inbox:
            L.afs_ipc_free(C.c_void_p(self.inbox))
            self.inbox = None
        self.stripe.close()
