"""Drop-in for the reference's ``livenote_v2.py``: ``LiveNoteV2(ref, params, debug_params, chroma_diff=False)``.

Reference: livenote_v2.py:3-236 — the OTW schedule with +inf fill, forward-only path
filter (:198-199) and optional Euclidean cost for chroma-difference features
(:167-170).  Runs in kernel K5 (csrc/otw.cu) as one stream of ``batch.OtwBatch``.
Kept: ``insert`` -> None | "stop", ``set_live``, ``.path``, ``.live_ptr``, ``.ref_ptr``.
"""
try:
    from ._stream import SingleStream
except ImportError:
    from _stream import SingleStream


class LiveNoteV2(SingleStream):
    def __init__(self, ref, params, debug_params=None, chroma_diff=False):
        self.search_band_width = params['search_band_width']
        self.max_run_count = params['max_run_count']
        self.seq_ref = ref
        self.chroma_diff = chroma_diff
        SingleStream.__init__(self, "livenote_v2", ref, self.search_band_width, self.max_run_count, chroma_diff=chroma_diff)

    @property
    def live_ptr(self):
        return int(self._positions()[0])

    @property
    def ref_ptr(self):
        return int(self._positions()[1])

    def set_live(self, live):
        """livenote_v2.py:108-155 on a fresh object (the reference does not reset state)."""
        if not self.path and self._positions()[0] == 0:
            self._batch.seed_set_live()
        self.path = self._run_all(live, from_start=True)
