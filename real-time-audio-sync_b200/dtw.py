"""Drop-in for the reference's ``dtw.py``: ``DTW(seq_a, seq_b) -> (cost, acc_cost, path)``.

Reference: dtw.py:5-53 (cost = 1 - seq_a.T @ seq_b; steps left/up/diag with
weights 1,1,2; first minimum wins in that order; backtrack from the end).
The arithmetic runs in kernels K2/K3 of libafsync (csrc/dtw.cu) through the
C ABI declared in include/afsync.h; there is no CPU path.

Deviation kept small and explicit: the reference always materialises the dense
``cost`` and ``acc_cost`` matrices (24 B/cell, 9.6 GB at 20k x 20k).  Here they are
produced only when ``M*N <= dense_limit`` (default 2**26 cells); above that the
first two tuple members are ``None`` and ``DTW.last_acc_end`` holds acc_cost[-1,-1].
"""
import ctypes as C

import numpy as np
import torch

try:
    from . import _native as nat
except ImportError:  # flat import (directory on sys.path, like the reference's modules)
    import _native as nat

DENSE_LIMIT = 1 << 26


class DtwPlan(object):
    """A batch of independent (seq_a, seq_b) pairs with fixed lengths: owns the
    libafsync plan, its device workspace and the path buffers; reusable."""

    def __init__(self, lens_a, lens_b, dtype="fp64", offs_a=None, offs_b=None, device=None):
        nat.require_cuda()
        self.device = nat.device() if device is None else torch.device(device)
        self.dtype = nat.AFS_F64 if dtype in ("fp64", "f64", torch.float64, np.float64) else nat.AFS_F32
        self.torch_dtype = torch.float64 if self.dtype == nat.AFS_F64 else torch.float32
        self.lens_a = np.ascontiguousarray(lens_a, dtype=np.int64).reshape(-1)
        self.lens_b = np.ascontiguousarray(lens_b, dtype=np.int64).reshape(-1)
        self.n_pairs = int(self.lens_a.shape[0])
        assert self.lens_b.shape[0] == self.n_pairs
        if offs_a is None:
            offs_a = np.concatenate(([0], np.cumsum(self.lens_a * 12)[:-1]))
        if offs_b is None:
            offs_b = np.concatenate(([0], np.cumsum(self.lens_b * 12)[:-1]))
        self.offs_a = np.ascontiguousarray(offs_a, dtype=np.int64)
        self.offs_b = np.ascontiguousarray(offs_b, dtype=np.int64)
        L = nat.lib()
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            nat.check(L.afs_dtw_plan_create(C.byref(h), self.n_pairs,
                                            self.lens_a.ctypes.data_as(nat._i64p), self.lens_b.ctypes.data_as(nat._i64p),
                                            self.offs_a.ctypes.data_as(nat._i64p), self.offs_b.ctypes.data_as(nat._i64p),
                                            12, self.dtype))
        self._h = h
        nbytes = C.c_size_t()
        nat.check(L.afs_dtw_plan_workspace_bytes(self._h, C.byref(nbytes)))
        self.workspace_bytes = int(nbytes.value)
        self.workspace = torch.empty(self.workspace_bytes, dtype=torch.uint8, device=self.device)
        off, cap = C.c_int64(), C.c_int64()
        nat.check(L.afs_dtw_plan_path_layout(self._h, -1, C.byref(off), C.byref(cap)))
        self.path_total = int(cap.value)
        self.path = torch.empty((self.path_total, 2), dtype=torch.int32, device=self.device)
        self.path_start = torch.empty(self.n_pairs, dtype=torch.int32, device=self.device)
        self.path_len = torch.empty(self.n_pairs, dtype=torch.int32, device=self.device)
        self.acc_end = torch.empty(self.n_pairs, dtype=torch.float64, device=self.device)
        self.path_off = np.concatenate(([0], np.cumsum(self.lens_a + self.lens_b)[:-1])).astype(np.int64)

    def close(self):
        if getattr(self, "_h", None):
            nat.lib().afs_dtw_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def cells(self):
        return int((self.lens_a * self.lens_b).sum())

    def accumulate(self, d_a, d_b, dense_cost=None, dense_acc=None):
        """K2 on the current stream.  d_a / d_b: device tensors holding every pair's
        (12, len) block at the plan's offsets, in the plan dtype."""
        assert d_a.dtype == self.torch_dtype and d_b.dtype == self.torch_dtype
        assert d_a.is_cuda and d_b.is_cuda
        with torch.cuda.device(self.device):      # the launch goes to the plan's GPU whatever the caller's current device is
            nat.check(nat.lib().afs_dtw_accumulate(self._h, nat.ptr(d_a), nat.ptr(d_b), nat.ptr(self.workspace),
                                                    nat.ptr(self.acc_end), nat.ptr(dense_cost), nat.ptr(dense_acc),
                                                    nat.stream_ptr()))

    def backtrack(self):
        """K3 on the current stream."""
        with torch.cuda.device(self.device):
            nat.check(nat.lib().afs_dtw_backtrack(self._h, nat.ptr(self.workspace), nat.ptr(self.path),
                                                   nat.ptr(self.path_start), nat.ptr(self.path_len), nat.stream_ptr()))

    def run(self, d_a, d_b):
        self.accumulate(d_a, d_b)
        self.backtrack()

    def paths(self):
        """Host copies of all paths as int64 (P,2) arrays (synchronises)."""
        start = self.path_start.cpu().numpy()
        length = self.path_len.cpu().numpy()
        flat = self.path.cpu().numpy()
        out = []
        for p in range(self.n_pairs):
            s = int(self.path_off[p] + start[p])
            out.append(flat[s : s + int(length[p])].astype(np.int64))
        return out


class DtwPipeline(object):
    """A stream of same-shape batches with ``depth`` batches in flight (default 2).

    Each slot owns a plan (workspace, path buffers), a compute stream, device input buffers and pinned host result
    buffers.  ``submit(h_a, h_b)`` uploads one batch from (preferably pinned) host tensors on an upload stream and launches
    K2 + K3 on the slot's compute stream; the read-back of its paths is queued behind them on a download stream.  It
    returns the results of the batch that previously used the slot (``None`` while the pipeline fills); ``drain()``
    returns what is still in flight.  Uploads, kernels and read-backs of neighbouring batches overlap, and the
    ramp-down of one launch overlaps the ramp-up of the next.  Results: ``(paths, acc_end)`` as ``dtw_batch`` gives."""

    def __init__(self, lens_a, lens_b, dtype="fp64", depth=2, device=None):
        nat.require_cuda()
        self.depth = int(depth)
        assert self.depth >= 1
        self.plans = [DtwPlan(lens_a, lens_b, dtype=dtype, device=device) for _ in range(self.depth)]
        p0 = self.plans[0]
        self.device = p0.device
        n_a, n_b = int((p0.lens_a * 12).sum()), int((p0.lens_b * 12).sum())
        with torch.cuda.device(self.device):
            self.copy_stream = torch.cuda.Stream()       # host -> device
            self.down_stream = torch.cuda.Stream()       # device -> host (its own stream: a read-back waits for its
                                                         # batch's kernels and must not hold up the next upload)
            self.streams = [torch.cuda.Stream() for _ in range(self.depth)]
        self.d_a = [torch.empty(n_a, dtype=p0.torch_dtype, device=self.device) for _ in range(self.depth)]
        self.d_b = [torch.empty(n_b, dtype=p0.torch_dtype, device=self.device) for _ in range(self.depth)]
        self.h_path = [torch.empty((p0.path_total, 2), dtype=torch.int32).pin_memory() for _ in range(self.depth)]
        self.h_start = [torch.empty(p0.n_pairs, dtype=torch.int32).pin_memory() for _ in range(self.depth)]
        self.h_len = [torch.empty(p0.n_pairs, dtype=torch.int32).pin_memory() for _ in range(self.depth)]
        self.h_end = [torch.empty(p0.n_pairs, dtype=torch.float64).pin_memory() for _ in range(self.depth)]
        self.done = [None] * self.depth          # read-back complete (copy stream)
        self.computed = [None] * self.depth      # K2 + K3 complete (compute stream)
        self.k = 0
        self.h2d_bytes = (n_a + n_b) * self.d_a[0].element_size()
        self.d2h_bytes = p0.path_total * 8 + p0.n_pairs * 16

    def _collect(self, q, as_arrays=True):
        self.done[q].synchronize()
        self.done[q] = None
        if not as_arrays:
            return self.h_path[q], self.h_start[q], self.h_len[q], self.h_end[q]
        plan, flat = self.plans[q], self.h_path[q].numpy()
        start, length = self.h_start[q].numpy(), self.h_len[q].numpy()
        paths = []
        for p in range(plan.n_pairs):
            s0 = int(plan.path_off[p] + start[p])
            paths.append(flat[s0 : s0 + int(length[p])].astype(np.int64))
        return paths, self.h_end[q].numpy().copy()

    def submit(self, h_a, h_b, as_arrays=True):
        q = self.k % self.depth
        self.k += 1
        prev = self._collect(q, as_arrays) if self.done[q] is not None else None
        plan, cs, st = self.plans[q], self.copy_stream, self.streams[q]
        with torch.cuda.stream(cs):
            if self.computed[q] is not None:
                cs.wait_event(self.computed[q])              # the slot's previous kernels no longer read its inputs
            self.d_a[q].copy_(h_a.reshape(-1), non_blocking=True)
            self.d_b[q].copy_(h_b.reshape(-1), non_blocking=True)
            up = torch.cuda.Event()
            up.record(cs)
        with torch.cuda.stream(st):
            st.wait_event(up)
            plan.accumulate(self.d_a[q], self.d_b[q])
            plan.backtrack()                                 # the slot's previous read-back was collected above
            self.computed[q] = torch.cuda.Event()
            self.computed[q].record(st)
        ds = self.down_stream
        with torch.cuda.stream(ds):
            ds.wait_event(self.computed[q])
            self.h_start[q].copy_(plan.path_start, non_blocking=True)
            self.h_len[q].copy_(plan.path_len, non_blocking=True)
            self.h_path[q].copy_(plan.path, non_blocking=True)
            self.h_end[q].copy_(plan.acc_end, non_blocking=True)
            self.done[q] = torch.cuda.Event()
            self.done[q].record(ds)
        return prev

    def drain(self, as_arrays=True):
        out = []
        for i in range(self.depth):
            q = (self.k + i) % self.depth
            if self.done[q] is not None:
                out.append(self._collect(q, as_arrays))
        return out

    def close(self):
        for p in self.plans:
            p.close()


def dtw_batch(seqs_a, seqs_b, dtype="fp64"):
    """Align many pairs in one launch.  seqs_a/seqs_b: lists of (12, len) arrays.
    Returns (paths, acc_end) with paths a list of int64 (P,2) arrays."""
    nat.require_cuda()
    lens_a = [np.shape(a)[1] for a in seqs_a]
    lens_b = [np.shape(b)[1] for b in seqs_b]
    plan = DtwPlan(lens_a, lens_b, dtype=dtype)
    npdt = np.float64 if plan.dtype == nat.AFS_F64 else np.float32
    flat_a = np.concatenate([np.ascontiguousarray(a, dtype=npdt).reshape(-1) for a in seqs_a])
    flat_b = np.concatenate([np.ascontiguousarray(b, dtype=npdt).reshape(-1) for b in seqs_b])
    d_a = torch.from_numpy(flat_a).to(plan.device)
    d_b = torch.from_numpy(flat_b).to(plan.device)
    plan.run(d_a, d_b)
    paths = plan.paths()
    acc_end = plan.acc_end.cpu().numpy()
    plan.close()
    return paths, acc_end


def DTW(seq_a, seq_b, dtype="fp64", dense_limit=None):
    """dtw.py:5-53.  seq_a (12, M), seq_b (12, N) -> (cost (M,N), acc_cost (M,N), path (P,2) int64)."""
    nat.require_cuda()
    a = np.asarray(seq_a)
    b = np.asarray(seq_b)
    assert a.ndim == 2 and b.ndim == 2 and a.shape[0] == b.shape[0], "sequences must be (features, frames)"
    if a.shape[0] != 12:
        raise nat.AfsError("the CUDA path is specialised for 12 chroma features (got %d)" % a.shape[0])
    M, N = a.shape[1], b.shape[1]
    plan = DtwPlan([M], [N], dtype=dtype)
    npdt = np.float64 if plan.dtype == nat.AFS_F64 else np.float32
    d_a = torch.from_numpy(np.ascontiguousarray(a, dtype=npdt)).to(plan.device)
    d_b = torch.from_numpy(np.ascontiguousarray(b, dtype=npdt)).to(plan.device)
    limit = DENSE_LIMIT if dense_limit is None else dense_limit
    cost = acc = None
    if M * N <= limit:
        d_cost = torch.empty((M, N), dtype=plan.torch_dtype, device=plan.device)
        d_acc = torch.empty((M, N), dtype=plan.torch_dtype, device=plan.device)
        plan.accumulate(d_a, d_b, d_cost, d_acc)
    else:
        d_cost = d_acc = None
        plan.accumulate(d_a, d_b)
    plan.backtrack()
    path = plan.paths()[0]
    DTW.last_acc_end = float(plan.acc_end.cpu()[0])
    if d_cost is not None:
        cost = d_cost.cpu().numpy().astype(np.float64, copy=False)
        acc = d_acc.cpu().numpy().astype(np.float64, copy=False)
    plan.close()
    return cost, acc, path


DTW.last_acc_end = None
