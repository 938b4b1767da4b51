"""Batched front-ends: many independent streams / pairs / tracks per launch.

``OtwBatch`` advances n independent OnlineTimeWarping / LiveNoteV2 / LiveNote
objects (kernel K5, csrc/otw.cu) with one launch per live frame (or one launch
for a block of frames).  The single-stream classes in otw_eran.py / livenote_v2.py
/ livenote.py are this class with n = 1.
"""
import ctypes as C

import numpy as np
import torch

try:
    from . import _native as nat
except ImportError:
    import _native as nat


def _pack_refs(refs, device):
    """refs: list of (12, N_s) arrays or one (n, 12, N) array -> (flat device tensor, lens, offs)."""
    if isinstance(refs, torch.Tensor):
        assert refs.dim() == 3 and refs.shape[1] == 12
        n, _, N = refs.shape
        flat = refs.to(device=device, dtype=torch.float64).contiguous().reshape(-1)
        lens = np.full(n, N, dtype=np.int64)
    else:
        arrs = [np.ascontiguousarray(r, dtype=np.float64) for r in refs]
        for r in arrs:
            if r.ndim != 2 or r.shape[0] != 12:
                raise nat.AfsError("reference sequences must be (12, frames) (got %r)" % (r.shape,))
        lens = np.array([r.shape[1] for r in arrs], dtype=np.int64)
        flat = torch.from_numpy(np.concatenate([r.reshape(-1) for r in arrs])).to(device)
    offs = np.concatenate(([0], np.cumsum(lens * 12)[:-1])).astype(np.int64)
    return flat, lens, offs


class OtwBatch(object):
    """n independent online aligners sharing (c, max_run_count, kind, metric)."""

    def __init__(self, refs, c, max_run_count, kind="otw", chroma_diff=False, device=None):
        nat.require_cuda()
        self.device = nat.device() if device is None else torch.device(device)
        self.kind = {"otw": nat.AFS_OTW, "livenote_v2": nat.AFS_LIVENOTE_V2, "livenote": nat.AFS_LIVENOTE_V1}[kind]
        self.c = int(c)
        self.chroma_diff = bool(chroma_diff)
        self.max_run_count = int(max_run_count)
        with torch.cuda.device(self.device):
            self.d_ref, self.ref_lens, self.ref_offs = _pack_refs(refs, self.device)
            self.n = int(self.ref_lens.shape[0])
            L = nat.lib()
            h = C.c_void_p()
            nat.check(L.afs_otw_create(C.byref(h), self.kind, self.n, nat.ptr(self.d_ref),
                                       self.ref_lens.ctypes.data_as(nat._i64p), self.ref_offs.ctypes.data_as(nat._i64p),
                                       12, self.c, self.max_run_count,
                                       nat.AFS_COST_EUCLID if chroma_diff else nat.AFS_COST_COSINE))
            self._h = h
            nbytes = C.c_size_t()
            nat.check(L.afs_otw_state_bytes(self._h, C.byref(nbytes)))
            self.state_bytes = int(nbytes.value)
            self.state = torch.empty(self.state_bytes, dtype=torch.uint8, device=self.device)
            self.pts = int(L.afs_otw_points_per_step(self._h))
            self.reset()
        off, cap = C.c_int64(), C.c_int64()
        self.path_off = np.empty(self.n, dtype=np.int64)
        self.path_cap = np.empty(self.n, dtype=np.int64)
        for s in range(self.n):
            nat.check(nat.lib().afs_otw_path_layout(self._h, s, C.byref(off), C.byref(cap)))
            self.path_off[s], self.path_cap[s] = off.value, cap.value
        nat.check(nat.lib().afs_otw_path_layout(self._h, -1, C.byref(off), C.byref(cap)))
        self.path_total = int(cap.value)
        self._out = {}
        self._flat = {}

    def reset(self):
        nat.check(nat.lib().afs_otw_reset(self._h, nat.ptr(self.state), nat.stream_ptr()))

    def seed_set_live(self):
        """State after set_live()'s loop-top best-point call at (0,0) (otw_eran.py:100-108)."""
        nat.check(nat.lib().afs_otw_seed_set_live(self._h, nat.stream_ptr()))

    def close(self):
        if getattr(self, "_h", None):
            nat.lib().afs_otw_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _outputs(self, n_frames):
        if n_frames not in self._out:
            # status | npoints | points are views of ONE buffer, so a host that wants all three needs a
            # single device-to-host copy per step (`outputs_flat`)
            k = n_frames * self.n
            flat = torch.empty(k * (2 + 2 * self.pts), dtype=torch.int32, device=self.device)
            self._out[n_frames] = (
                flat[:k].view(n_frames, self.n),
                flat[k : 2 * k].view(n_frames, self.n),
                flat[2 * k :].view(n_frames, self.n, self.pts, 2),
            )
            self._flat[n_frames] = flat
        return self._out[n_frames]

    def outputs_flat(self, n_frames=1):
        """The buffer behind (status, npoints, points) of `step_device` for this n_frames."""
        self._outputs(n_frames)
        return self._flat[n_frames]

    def step_device(self, d_frames, active=None, want_points=True):
        """d_frames: device tensor (n_frames, n, 12) or (n, 12) float64.  Asynchronous.
        Returns device tensors (status, npoints, points)."""
        if d_frames.dim() == 2:
            d_frames = d_frames.unsqueeze(0)
        assert d_frames.is_cuda and d_frames.dtype == torch.float64 and d_frames.is_contiguous()
        assert d_frames.shape[1] == self.n and d_frames.shape[2] == 12
        nf = int(d_frames.shape[0])
        st, npts, pts = self._outputs(nf)
        with torch.cuda.device(self.device):      # the launch goes to the batch's GPU whatever the caller's current device is
            nat.check(nat.lib().afs_otw_step(self._h, nat.ptr(d_frames), nf, nat.ptr(active), nat.ptr(st), nat.ptr(npts),
                                             nat.ptr(pts) if want_points else C.c_void_p(0), nat.stream_ptr()))
        return st, npts, pts

    def insert(self, frames):
        """Host API: frames (n, 12) array-like -> (status (n,), list of appended points per stream)."""
        fr = torch.from_numpy(np.ascontiguousarray(frames, dtype=np.float64).reshape(1, self.n, 12)).to(self.device)
        st, npts, pts = self.step_device(fr)
        st = st.cpu().numpy()[0]
        npts = npts.cpu().numpy()[0]
        pts = pts.cpu().numpy()[0]
        return st, [[(int(x), int(y)) for x, y in pts[s, : npts[s]]] for s in range(self.n)]

    # ---- checkpoint / resume: every bit of stream state (rings, scalars, paths) lives in the one device block
    # `self.state`, addressed by offsets only, so a host copy of it is a complete snapshot (the reference keeps the
    # same information in in-object numpy arrays, otw_eran.py:17-36) ----
    def _state_meta(self):
        return {"class": type(self).__name__, "kind": int(self.kind), "c": self.c, "max_run_count": self.max_run_count,
                "ref_lens": self.ref_lens.tolist(), "state_bytes": self.state_bytes}

    def export_state(self):
        """Snapshot of all streams (synchronises): {"meta": ..., "state": uint8 CPU tensor}."""
        return {"meta": self._state_meta(), "state": self.state.cpu().clone()}

    def import_state(self, snap):
        """Resume from `export_state()` of a batch built with the same references and parameters."""
        if snap["meta"] != self._state_meta():
            raise nat.AfsError("snapshot does not match this batch: %r vs %r" % (snap["meta"], self._state_meta()))
        self.state.copy_(snap["state"].to(self.device))

    def _dev_view(self, address, count, dtype):
        """Wrap `count` int32s at a device address inside the state block as a tensor view."""
        base = self.state.data_ptr()
        off = address - base
        assert 0 <= off and off + count * 4 <= self.state_bytes
        return self.state[off : off + count * 4].view(dtype)

    def positions(self):
        p = C.c_void_p()
        nat.check(nat.lib().afs_otw_positions_ptr(self._h, C.byref(p)))
        return self._dev_view(p.value, 2 * self.n, torch.int32).cpu().numpy().reshape(self.n, 2)

    def window(self, stream=0):
        """On-demand view of one stream's state in the reference's terms (otw_eran.py:23-35, livenote_v2.py:22-38):
        dict with the scalars t, j, previous, run_count, direction (as the reference's strings) and the two live lines
        of acc_cost — ``acc_row`` = acc_cost[t, j-c .. j], ``acc_col`` = acc_cost[t-c .. t, j] — with the column / row
        indices they belong to, plus the last c+1 live frames.  Indices before the matrix start are dropped."""
        c = int(self.c)
        scal = np.zeros(8, dtype=np.int32)
        row = np.empty(c + 1, dtype=np.float64)
        col = np.empty(c + 1, dtype=np.float64)
        live = np.empty((12, c + 1), dtype=np.float64)
        with torch.cuda.device(self.device):
            nat.check(nat.lib().afs_otw_read_window(self._h, int(stream), scal.ctypes.data_as(C.c_void_p), row.ctypes.data_as(C.c_void_p),
                                                    col.ctypes.data_as(C.c_void_p), live.ctypes.data_as(C.c_void_p), nat.stream_ptr()))
        t, j = int(scal[0]), int(scal[1])
        names = {0: "Both", 1: "Row", 2: "Column"}
        cols = np.arange(j - c, j + 1)
        rows = np.arange(t - c, t + 1)
        return {"t": t, "j": j, "previous": None if scal[2] == 0 else names[int(scal[2])], "run_count": int(scal[3]),
                "direction": names[int(scal[4])], "first_insert": bool(scal[5]), "status": int(scal[6]), "path_len": int(scal[7]),
                "cols": cols[cols >= 0], "acc_row": row[cols >= 0], "rows": rows[rows >= 0], "acc_col": col[rows >= 0],
                "live": live[:, rows >= 0]}

    def paths(self):
        """Full device-resident paths as a list of int64 (P,2) arrays."""
        pp, pl = C.c_void_p(), C.c_void_p()
        nat.check(nat.lib().afs_otw_path_ptr(self._h, C.byref(pp), C.byref(pl)))
        lens = self._dev_view(pl.value, self.n, torch.int32).cpu().numpy()
        flat = self._dev_view(pp.value, 2 * self.path_total, torch.int32).cpu().numpy().reshape(-1, 2)
        out = []
        for s in range(self.n):
            n = int(min(lens[s], self.path_cap[s]))
            out.append(flat[self.path_off[s] : self.path_off[s] + n].astype(np.int64))
        return out


class WtwBatch(object):
    """n independent windowed-time-warping aligners fed live chroma columns (kernel K6,
    csrc/wtw.cu; reference wtw.py:94-128).  W = dtw_win_size/hop_size, h = dtw_hop_size/hop_size."""

    def __init__(self, refs, W, h, device=None):
        nat.require_cuda()
        self.device = nat.device() if device is None else torch.device(device)
        self.W, self.h = int(W), int(h)
        with torch.cuda.device(self.device):
            self.d_ref, self.ref_lens, self.ref_offs = _pack_refs(refs, self.device)
            self.n = int(self.ref_lens.shape[0])
            L = nat.lib()
            hd = C.c_void_p()
            nat.check(L.afs_wtw_create(C.byref(hd), self.n, nat.ptr(self.d_ref), self.ref_lens.ctypes.data_as(nat._i64p),
                                       self.ref_offs.ctypes.data_as(nat._i64p), 12, self.W, self.h))
            self._h = hd
            nbytes = C.c_size_t()
            nat.check(L.afs_wtw_state_bytes(self._h, C.byref(nbytes)))
            self.state_bytes = int(nbytes.value)
            self.state = torch.empty(self.state_bytes, dtype=torch.uint8, device=self.device)
            self.reset()
        off, cap = C.c_int64(), C.c_int64()
        self.path_off = np.empty(self.n, dtype=np.int64)
        self.path_cap = np.empty(self.n, dtype=np.int64)
        for s in range(self.n):
            nat.check(nat.lib().afs_wtw_path_layout(self._h, s, C.byref(off), C.byref(cap)))
            self.path_off[s], self.path_cap[s] = off.value, cap.value
        nat.check(nat.lib().afs_wtw_path_layout(self._h, -1, C.byref(off), C.byref(cap)))
        self.path_total = int(cap.value)

    def reset(self):
        nat.check(nat.lib().afs_wtw_reset(self._h, nat.ptr(self.state), nat.stream_ptr()))

    def close(self):
        if getattr(self, "_h", None):
            nat.lib().afs_wtw_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def push_device(self, d_cols, active=None):
        """d_cols: device tensor (n_frames, n, 12) float64.  Returns the device status tensor (n_frames, n)."""
        if d_cols.dim() == 2:
            d_cols = d_cols.unsqueeze(0)
        assert d_cols.is_cuda and d_cols.dtype == torch.float64 and d_cols.is_contiguous()
        assert d_cols.shape[1] == self.n and d_cols.shape[2] == 12
        nf = int(d_cols.shape[0])
        st = torch.empty((nf, self.n), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            nat.check(nat.lib().afs_wtw_push(self._h, nat.ptr(d_cols), nf, nat.ptr(active), nat.ptr(st), nat.stream_ptr()))
        return st

    def push_audio_device(self, plan, d_audio, offsets, n_frames, compute="fp64", active=None):
        """Audio in, alignment out on the device (afs_wtw_push_audio: K1 -> reorder -> K6 on the current stream).
        d_audio: float32 device tensor holding one run of samples per stream, offsets (n + 1) int64 (even), every run
        exactly n_frames un-padded frames (wtw.py:81-83).  Returns the device status tensor (n_frames, n)."""
        assert d_audio.is_cuda and d_audio.dtype == torch.float32
        offs = np.ascontiguousarray(offsets, dtype=np.int64)
        assert offs.shape[0] == self.n + 1
        nf = int(n_frames)
        st = torch.empty((nf, self.n), dtype=torch.int32, device=self.device)
        scratch = torch.empty(2 * nf * self.n * 12, dtype=torch.float64, device=self.device)
        code = {"fp64": nat.AFS_F64, "f64": nat.AFS_F64, "fp32": nat.AFS_F32, "f32": nat.AFS_F32, "tc": nat.AFS_BF16X3}[compute]
        with torch.cuda.device(self.device):
            nat.check(nat.lib().afs_wtw_push_audio(self._h, plan._h, nat.ptr(d_audio), offs.ctypes.data_as(nat._i64p), nf, nat.ptr(scratch),
                                                   nat.ptr(active), nat.ptr(st), code, nat.stream_ptr()))
        return st

    def push(self, cols):
        """Host API: cols (n_frames, n, 12) or (n, 12) array -> status array (n_frames, n)."""
        c = np.ascontiguousarray(cols, dtype=np.float64)
        if c.ndim == 2:
            c = c.reshape(1, self.n, 12)
        return self.push_device(torch.from_numpy(c).to(self.device)).cpu().numpy()

    _dev_view = OtwBatch._dev_view
    export_state = OtwBatch.export_state
    import_state = OtwBatch.import_state

    def _state_meta(self):
        return {"class": type(self).__name__, "W": self.W, "h": self.h, "ref_lens": self.ref_lens.tolist(),
                "state_bytes": self.state_bytes}

    def positions(self):
        """(n, 3) array of (chroma_ptr, live_ptr, ref_ptr)."""
        p = C.c_void_p()
        nat.check(nat.lib().afs_wtw_positions_ptr(self._h, C.byref(p)))
        return self._dev_view(p.value, 3 * self.n, torch.int32).cpu().numpy().reshape(self.n, 3)

    def paths(self):
        pp, pl = C.c_void_p(), C.c_void_p()
        nat.check(nat.lib().afs_wtw_path_ptr(self._h, C.byref(pp), C.byref(pl)))
        lens = self._dev_view(pl.value, self.n, torch.int32).cpu().numpy()
        flat = self._dev_view(pp.value, 2 * self.path_total, torch.int32).cpu().numpy().reshape(-1, 2)
        out = []
        for s in range(self.n):
            n = int(min(lens[s], self.path_cap[s]))
            out.append(flat[self.path_off[s] : self.path_off[s] + n].astype(np.int64))
        return out
