"""Drop-in for the reference's ``wtw.py``: ``WTW(ref_recording, params, debug_params)``.

Reference: wtw.py:19-240 (windowed time warping).  ``insert(live_audio_buf)`` takes a list
of samples, frames it (fft_len / hop_size, no zero pad), turns each frame into a chroma
column with kernel K1 (wtw.py:82-90) and feeds the columns to kernel K6, which runs the
W x W windows (get_cost_matrix / run_dtw / find_path / stitching, wtw.py:100-128).
Returns None or "stop" exactly where the reference does; ``.path`` is the list of
(live, ref) tuples, seam duplicates included.

Chroma for this class is computed in float64 by default (``compute="fp64"``) so that the
reference's own golden path (Songs/chopin/tests/wtw_test_20b.txt) is reproduced exactly.
"""
import numpy as np
import torch

try:
    from . import _native as nat
    from . import chroma as _chroma
    from .batch import WtwBatch
except ImportError:
    import _native as nat
    import chroma as _chroma
    from batch import WtwBatch


class WTW():

    def __init__(self, ref_recording, params, debug_params=None, compute="fp64"):
        # reference audio, fs = 22050 (wtw.py:23-24); an ndarray of samples is accepted too
        if isinstance(ref_recording, str):
            self.ref, self.fs = _chroma.load_wav(ref_recording)
        else:
            self.ref, self.fs = np.asarray(ref_recording, dtype=np.float32), 22050
        assert(self.fs == 22050)

        # params (wtw.py:27-30)
        self.fft_len = params['fft_len']
        self.hop_size = params['hop_size']
        self.dtw_win_size = params['dtw_win_size']
        self.dtw_hop_size = params['dtw_hop_size']
        if self.fft_len != _chroma.fft_len:
            raise nat.AfsError("the CUDA path implements fft_len = 4096 only")
        self._compute = compute
        self._plan = _chroma.default_plan() if self.hop_size == _chroma.hop_size else _chroma.ChromaPlan(self.fft_len, self.hop_size)
        self._W = self.dtw_win_size // self.hop_size
        self._h = self.dtw_hop_size // self.hop_size

        # reference chromagram (wtw.py:37-41)
        self.chroma_ref = self._chroma_of(self.ref, center=True)
        self.N = self.chroma_ref.shape[1] * 2   # rows are live
        self.M = self.chroma_ref.shape[1]       # cols are ref
        self._batch = WtwBatch([self.chroma_ref], self._W, self._h)
        self.buf = []
        self._pos = (0, 0, 0)

    def _chroma_of(self, samples, center):
        x = np.ascontiguousarray(samples, dtype=np.float32)
        if len(x) & 1:
            x = np.concatenate((x, np.zeros(1, dtype=np.float32)))
            n_true = len(x) - 1
        else:
            n_true = len(x)
        d_audio = torch.from_numpy(x).to(self._plan.device)
        d_out, foffs = self._plan.run(d_audio, [0, n_true], center=center, out_dtype=torch.float64, compute=self._compute)
        return np.ascontiguousarray(d_out.cpu().numpy().reshape(12, -1))

    def insert(self, live_audio_buf):
        # store incoming music (wtw.py:73)
        self.buf += live_audio_buf
        chroma_ptr, live_ptr, ref_ptr = self._pos
        if ref_ptr >= self.M - 1 or live_ptr >= self.N - 1:          # wtw.py:76-77
            return "stop"
        if len(self.buf) < self.fft_len:
            return None
        # every complete frame of this call: frame q = buf[q*hop : q*hop + fft_len]  (wtw.py:81-83)
        n_frames = (len(self.buf) - self.fft_len) // self.hop_size + 1
        used = (n_frames - 1) * self.hop_size + self.fft_len
        # chroma of every frame (wtw.py:84-90) and the windowed DTW steps they trigger, chained on the device
        x = np.zeros(used + (used & 1), dtype=np.float32)
        x[:used] = self.buf[:used]
        d_audio = torch.from_numpy(x).to(self._batch.device)
        status = self._batch.push_audio_device(self._plan, d_audio, [0, used], n_frames, compute=self._compute).cpu().numpy()[:, 0]
        if (status == nat.AFS_STEP_FULL).any():
            raise IndexError("WTW.insert: the live chroma buffer (2 x reference frames) is full (wtw.py:92 raises here too)")
        stops = np.nonzero(status == nat.AFS_STEP_STOP)[0]
        self._pos = tuple(int(v) for v in self._batch.positions()[0])
        if len(stops):
            # the reference returns at the first stopping frame and leaves the rest of the buffer unread
            self.buf = self.buf[(int(stops[0]) + 1) * self.hop_size:]
            return "stop"
        self.buf = self.buf[n_frames * self.hop_size:]
        return None

    def insert_chroma(self, col):
        """Feed one ready chroma column (what each trip of the loop at wtw.py:81-93 produces)."""
        st = self._batch.push(np.asarray(col, dtype=np.float64).reshape(1, 1, 12))[0, 0]
        self._pos = tuple(int(v) for v in self._batch.positions()[0])
        return "stop" if st == nat.AFS_STEP_STOP else None

    @property
    def path(self):
        return [(int(x), int(y)) for x, y in self._batch.paths()[0]]

    @property
    def chroma_ptr(self):
        return self._pos[0]

    @property
    def live_ptr(self):
        return self._pos[1]

    @property
    def ref_ptr(self):
        return self._pos[2]
