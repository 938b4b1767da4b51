"""Streaming front door: many microphones / files at once.

Replaces, for n concurrent streams, the per-stream loop of the reference's live app
(``livenote_live.py:161-209``: ``receive_audio`` accumulates samples; ``_process_input`` takes
4096-sample frames, hop 2048, through ``wav_to_chroma_col`` and ``OnlineTimeWarping.insert``,
then reads ``path[-1]`` as the current position).  Audio capture itself stays on the CPU and
only feeds samples (BASELINE.json north_star); everything from the frame on runs on the GPU:
one K1 launch turns every stream's pending frames into chroma columns (un-padded frames,
chroma.py:35-42) and one K5 launch per frame index advances all streams that have a frame.
"""
import numpy as np
import torch

try:
    from . import _native as nat
    from . import chroma as _chroma
    from .batch import OtwBatch
except ImportError:
    import _native as nat
    import chroma as _chroma
    from batch import OtwBatch


class StreamFrontDoor(object):
    def __init__(self, refs, c, max_run_count, kind="otw", compute="fp32"):
        self.batch = OtwBatch(refs, c, max_run_count, kind=kind)
        self.n = self.batch.n
        self.plan = _chroma.default_plan()
        self.compute = compute
        self.buf = [np.zeros(0, dtype=np.float32) for _ in range(self.n)]
        self.position = [None] * self.n          # path[-1] per stream, as the live app reads it
        self.stopped = np.zeros(self.n, dtype=bool)
        self.fft_len, self.hop = _chroma.fft_len, _chroma.hop_size

    def feed(self, chunks):
        """chunks: list of n 1-D sample arrays (possibly empty).  Returns the list of streams that
        reported "stop" during this call.  Frames are consumed exactly like livenote_live.py:186-208:
        while len(data) >= 4096: process data[:4096]; data = data[2048:]."""
        assert len(chunks) == self.n
        n_frames = np.zeros(self.n, dtype=np.int64)
        for s, ch in enumerate(chunks):
            if len(ch):
                self.buf[s] = np.concatenate((self.buf[s], np.asarray(ch, dtype=np.float32)))
            if len(self.buf[s]) >= self.fft_len and not self.stopped[s]:
                n_frames[s] = (len(self.buf[s]) - self.fft_len) // self.hop + 1
        total = int(n_frames.sum())
        if total == 0:
            return []
        # ---- K1: all pending frames of all streams in one launch (center=False: no zero pad) ----
        lens, parts = [], []
        for s in range(self.n):
            used = (n_frames[s] - 1) * self.hop + self.fft_len if n_frames[s] else 0
            parts.append(self.buf[s][:used])
            lens.append(used)                      # used is even (hop and fft_len are even)
        offs = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
        d_audio = torch.from_numpy(np.concatenate(parts) if total else np.zeros(0, np.float32)).to(self.batch.device)
        d_chroma, foffs = self.plan.run(d_audio, offs, center=False, out_dtype=torch.float64, compute=self.compute)
        # ---- K5: frame index q of every stream that has one, one launch per q ----
        stops = []
        max_f = int(n_frames.max())
        cols = torch.zeros((max_f, self.n, 12), dtype=torch.float64, device=self.batch.device)
        active_h = np.zeros((max_f, self.n), dtype=np.uint8)          # host copy: the loop below never indexes a device tensor
        for s in range(self.n):
            k = int(n_frames[s])
            if k:
                blk = d_chroma[12 * foffs[s] : 12 * foffs[s + 1]].view(12, k)
                cols[:k, s, :] = blk.t()
                active_h[:k, s] = 1
        active = torch.from_numpy(active_h).to(self.batch.device)
        # all launches first (asynchronous; the batch reuses its output block, so each step's outputs are copied aside on
        # the device), one synchronising read-back at the end
        outs = []
        for q in range(max_f):
            st, npts, pts = self.batch.step_device(cols[q].contiguous(), active=active[q].contiguous())
            outs.append((st.clone(), npts.clone(), pts.clone()))
        host = [(st.cpu().numpy()[0], npts.cpu().numpy()[0], pts.cpu().numpy()[0]) for st, npts, pts in outs]
        for q in range(max_f):
            st_h, np_h, pts_h = host[q]
            for s in np.nonzero(active_h[q])[0]:
                if not self.stopped[s]:
                    if np_h[s] > 0:
                        self.position[s] = (int(pts_h[s, np_h[s] - 1, 0]), int(pts_h[s, np_h[s] - 1, 1]))
                    if st_h[s] == nat.AFS_STEP_STOP:
                        self.stopped[s] = True
                        stops.append(s)
        for s in range(self.n):
            if n_frames[s]:
                self.buf[s] = self.buf[s][int(n_frames[s]) * self.hop :]
        return stops

    def paths(self):
        return self.batch.paths()

    def close(self):
        self.batch.close()
