"""Drop-in for the reference's ``livenote.py`` (v1): ``LiveNote(ref, params, debug_params)``.

Same kernel K5 as LiveNoteV2 without the forward-only path filter and with the
cosine cost only."""
try:
    from ._stream import SingleStream
except ImportError:
    from _stream import SingleStream


class LiveNote(SingleStream):
    def __init__(self, ref, params, debug_params=None):
        self.search_band_width = params['search_band_width']
        self.max_run_count = params['max_run_count']
        self.seq_ref = ref
        SingleStream.__init__(self, "livenote", ref, self.search_band_width, self.max_run_count)

    @property
    def live_ptr(self):
        return int(self._positions()[0])

    @property
    def ref_ptr(self):
        return int(self._positions()[1])

    def set_live(self, live):
        if not self.path and self._positions()[0] == 0:
            self._batch.seed_set_live()
        self.path = self._run_all(live, from_start=True)
