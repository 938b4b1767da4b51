"""Single-stream wrapper shared by otw_eran.OnlineTimeWarping, livenote_v2.LiveNoteV2
and livenote.LiveNote: one stream of an OtwBatch (kernel K5), with the reference's
attribute names.  State lives on the GPU; `.path` is a host list kept in step."""
import numpy as np
import torch

try:
    from . import _native as nat
    from .batch import OtwBatch
except ImportError:
    import _native as nat
    from batch import OtwBatch


class SingleStream(object):
    def __init__(self, kind, ref, band, max_run_count, chroma_diff=False):
        ref = np.asarray(ref)
        assert ref.ndim == 2, "ref must be (features, frames)"
        if ref.shape[0] != 12:
            raise nat.AfsError("the CUDA path is specialised for 12 chroma features (got %d)" % ref.shape[0])
        self._kind = kind
        self._batch = OtwBatch([ref], band, max_run_count, kind=kind, chroma_diff=chroma_diff)
        self._ref = ref
        self._stopped = False
        self.path = []

    # --- reference entry point: otw_eran.py:38 / livenote_v2.py:43 ---
    def insert(self, live_sample):
        col = np.asarray(live_sample, dtype=np.float64).reshape(-1)
        assert col.shape[0] == 12
        status, pts = self._batch.insert(col.reshape(1, 12))
        self.path.extend(pts[0])
        if status[0] == nat.AFS_STEP_STOP:
            self._stopped = True
            return "stop"
        if status[0] == nat.AFS_STEP_FULL:
            print("Done. Ran out of room in pre-allocated live-sequence")
        return None

    def _run_all(self, live, from_start=False):
        """Feed a whole (12, M) live sequence in ONE launch; returns the appended points
        (or the whole device-resident path when from_start)."""
        live = np.ascontiguousarray(np.asarray(live, dtype=np.float64))
        assert live.ndim == 2 and live.shape[0] == 12
        frames = torch.from_numpy(np.ascontiguousarray(live.T).reshape(live.shape[1], 1, 12)).to(self._batch.device)
        before = 0 if from_start else len(self._batch.paths()[0])
        st, _, _ = self._batch.step_device(frames, want_points=False)
        self._stopped = bool((st == nat.AFS_STEP_STOP).any().item())
        return [(int(x), int(y)) for x, y in self._batch.paths()[0][before:]]

    def _positions(self):
        return self._batch.positions()[0]

    # --- the reference's state attributes, read from the device on demand (otw_eran.py:23-35, livenote_v2.py:22-38) ---
    @property
    def direction(self):
        return self._batch.window(0)["direction"]

    @property
    def previous(self):
        return self._batch.window(0)["previous"]

    @property
    def run_count(self):
        return self._batch.window(0)["run_count"]

    def acc_cost_window(self):
        """The part of the reference's dense ``acc_cost`` the algorithm can still read: row t over columns j-c..j and
        column j over rows t-c..t, as ``{(x, y): value}`` (never-evaluated cells hold the reference's fill value)."""
        w = self._batch.window(0)
        out = {(w["t"], int(y)): float(v) for y, v in zip(w["cols"], w["acc_row"])}
        out.update({(int(x), w["j"]): float(v) for x, v in zip(w["rows"], w["acc_col"])})
        return out

    def cost_window(self):
        """Local costs of the same cells (the reference's dense ``cost``), recomputed from the live frames the device
        keeps: 1 - <live, ref> for chroma (otw_eran.py:216, livenote_v2.py:170), Euclidean distance for chroma_diff."""
        w = self._batch.window(0)
        ref = np.asarray(self._ref, dtype=np.float64)
        euclid = getattr(self._batch, "chroma_diff", False)

        def cost(lv, rf):
            return float(np.sqrt(np.sum((lv - rf) ** 2))) if euclid else float(1 - np.dot(lv, rf))

        cur = w["live"][:, -1]
        out = {(w["t"], int(y)): cost(cur, ref[:, int(y)]) for y in w["cols"]}
        out.update({(int(x), w["j"]): cost(w["live"][:, k], ref[:, w["j"]]) for k, x in enumerate(w["rows"])})
        return out

    def close(self):
        self._batch.close()
