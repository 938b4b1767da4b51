"""Drop-in for the reference's ``otw_eran.py``: ``OnlineTimeWarping(ref, params)``.

Reference: otw_eran.py:5-239 (Dixon 2005 on-line time warping).  The per-frame
search-window update runs in kernel K5 (csrc/otw.cu); this class is one stream of
``batch.OtwBatch``.  Kept: ``insert`` -> None | "stop", ``set_live``, ``.path`` (list of
(live, ref) tuples; ndarray after set_live), ``.t``, ``.j``, ``.c``, ``.max_run_count``.
Not kept: the dense ``.cost`` / ``.acc_cost`` (2N x N) matrices — the device state is
the moving (c+1)-wide row/column window only.
"""
import numpy as np

try:
    from ._stream import SingleStream
except ImportError:
    from _stream import SingleStream


class OnlineTimeWarping(SingleStream):
    def __init__(self, ref, params):
        self.c = params['c']
        self.max_run_count = params['max_run_count']
        self.ref = ref
        SingleStream.__init__(self, "otw", ref, self.c, self.max_run_count)

    @property
    def t(self):
        return int(self._positions()[0])

    @property
    def j(self):
        return int(self._positions()[1])

    def set_live(self, live):
        """otw_eran.py:91-142: whole live sequence at once.  Same sweeps as the insert
        loop; the only difference is one extra best-point call at (0,0) before the first
        step (afs_otw_seed_set_live), so the path starts with (0,0)."""
        self._batch.reset()
        self._batch.seed_set_live()
        self.path = np.array(self._run_all(live, from_start=True))
