"""TEST INFRASTRUCTURE — ctypes front-end of the C oracle (oracle/afs_oracle.c)
plus the numpy restatement of chroma.py.  Mirrors the reference's call
signatures so parity tests read like the reference's own scripts.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
import this module; the product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import librosa_restated as _lr

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libafs_oracle.so")
_lib = None

_f64p = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)


def build(force=False):
    src = os.path.join(_HERE, "afs_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libafs_oracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_dtw.restype = C.c_int64
        L.orc_dtw.argtypes = [_f64p, _f64p, C.c_int, C.c_int64, C.c_int64, _f64p, _f64p, _i64p, _f64p]
        L.orc_dtw_many.restype = C.c_int64
        L.orc_dtw_many.argtypes = [_f64p, _f64p, C.c_int, C.c_int64, C.c_int64, C.c_int, _f64p]
        L.orc_otw_create.restype = C.c_void_p
        L.orc_otw_create.argtypes = [C.c_int, _f64p, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int]
        L.orc_otw_destroy.argtypes = [C.c_void_p]
        L.orc_otw_insert.restype = C.c_int
        L.orc_otw_insert.argtypes = [C.c_void_p, _f64p]
        for name in ("orc_otw_path_len", "orc_otw_t", "orc_otw_j", "orc_otw_evals"):
            getattr(L, name).restype = C.c_int64
            getattr(L, name).argtypes = [C.c_void_p]
        L.orc_otw_path.restype = _i64p
        L.orc_otw_path.argtypes = [C.c_void_p]
        L.orc_otw_acc.restype = C.c_double
        L.orc_otw_acc.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
        L.orc_wtw_create.restype = C.c_void_p
        L.orc_wtw_create.argtypes = [_f64p, C.c_int, C.c_int64, C.c_int, C.c_int]
        L.orc_wtw_destroy.argtypes = [C.c_void_p]
        L.orc_wtw_should_stop.restype = C.c_int
        L.orc_wtw_should_stop.argtypes = [C.c_void_p]
        L.orc_wtw_push_chroma.restype = C.c_int
        L.orc_wtw_push_chroma.argtypes = [C.c_void_p, _f64p]
        for name in ("orc_wtw_path_len", "orc_wtw_live_ptr", "orc_wtw_ref_ptr"):
            getattr(L, name).restype = C.c_int64
            getattr(L, name).argtypes = [C.c_void_p]
        L.orc_wtw_path.restype = _i64p
        L.orc_wtw_path.argtypes = [C.c_void_p]
        for name in ("orc_dot_gemm", "orc_dot_strided", "orc_dot_contig"):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [_f64p, _f64p, C.c_int]
        L.orc_np_sum.restype = C.c_double
        L.orc_np_sum.argtypes = [_f64p, C.c_int]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_f64p)


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# ---------------------------------------------------------------- chroma.py
fft_len = 4096
hop_size = 2048
fs = 22050

_fb_cache = {}


def chroma_fb(sr=fs, n_fft=fft_len):
    key = (sr, n_fft)
    if key not in _fb_cache:
        _fb_cache[key] = _lr.filters_chroma(sr, n_fft)
    return _fb_cache[key]


def create_stft(wav, L=fft_len, H=hop_size):
    """chroma.py:44-65 (== wtw.py:137-160): left zero-pad L/2, symmetric Hann,
    tail dropped; numpy's own rfft, one frame per row here, transposed on return."""
    x = np.concatenate((np.zeros(L // 2), np.asarray(wav, dtype=np.float64)))
    n = len(x)
    m = int(((n - L) // H) + 1) if n >= L else 0
    out = np.empty((1 + L // 2, max(m, 0)), dtype=complex)
    win = np.hanning(L)
    for k in range(m):
        out[:, k] = np.fft.rfft(x[k * H : k * H + L] * win)
    return out


def create_chroma(ft, normalize=True):
    """chroma.py:67-75."""
    spec = np.abs(ft) ** 2
    raw = np.dot(chroma_fb(), spec)
    if not normalize:
        return raw
    return _lr.util_normalize(raw, norm=2, axis=0)


def wav_samples_to_chroma(wav):
    return create_chroma(create_stft(wav))


def wav_to_chroma(path):
    """chroma.py:25-33."""
    wav, sr = _lr.load(path)
    assert sr == 22050
    return wav_samples_to_chroma(wav)


def wav_to_chroma_col(buf):
    """chroma.py:35-42 — one un-padded frame."""
    assert len(buf) == fft_len
    section = np.array(buf)
    return create_chroma(np.fft.rfft(section * np.hanning(len(section))))


def chroma_diff(chroma):
    """chroma.py:77-90 on an already computed chromagram."""
    return np.clip(np.diff(chroma), 0, float("inf"))


# ---------------------------------------------------------------- dtw.py
def DTW(seq_a, seq_b, dense=True):
    """dtw.py:5-53 -> (cost, acc_cost, path); dense=False returns (None, acc_end, path)."""
    a = _c64(seq_a)
    b = _c64(seq_b)
    F, M = a.shape
    N = b.shape[1]
    assert b.shape[0] == F
    path = np.empty((M + N, 2), dtype=np.int64)
    end = C.c_double(0.0)
    if dense:
        cost = np.empty((M, N))
        acc = np.empty((M, N))
        n = lib().orc_dtw(_p(a), _p(b), F, M, N, _p(cost), _p(acc), path.ctypes.data_as(_i64p), C.byref(end))
        return cost, acc, path[:n].copy()
    n = lib().orc_dtw(_p(a), _p(b), F, M, N, None, None, path.ctypes.data_as(_i64p), C.byref(end))
    return None, end.value, path[:n].copy()


# ---------------------------------------------------------------- otw family
KIND_OTW, KIND_LN2, KIND_LN1 = 0, 1, 2


class _Stepper(object):
    def __init__(self, kind, ref, c, max_run_count, metric=0):
        ref = _c64(ref)
        self._ref = ref
        self._h = lib().orc_otw_create(kind, _p(ref), ref.shape[0], ref.shape[1], int(c), int(max_run_count), metric)
        self._F = ref.shape[0]

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_otw_destroy(self._h)
            self._h = None

    def insert(self, live_sample):
        col = _c64(live_sample).reshape(-1)
        assert col.shape[0] == self._F
        r = lib().orc_otw_insert(self._h, _p(col))
        return "stop" if r == 1 else None

    @property
    def path(self):
        n = lib().orc_otw_path_len(self._h)
        ptr = lib().orc_otw_path(self._h)
        arr = np.ctypeslib.as_array(ptr, shape=(n, 2)).copy() if n else np.empty((0, 2), dtype=np.int64)
        return [(int(x), int(y)) for x, y in arr]

    def path_array(self):
        n = lib().orc_otw_path_len(self._h)
        if not n:
            return np.empty((0, 2), dtype=np.int64)
        return np.ctypeslib.as_array(lib().orc_otw_path(self._h), shape=(n, 2)).copy()

    @property
    def t(self):
        return lib().orc_otw_t(self._h)

    @property
    def j(self):
        return lib().orc_otw_j(self._h)

    @property
    def n_evals(self):
        return lib().orc_otw_evals(self._h)

    def acc(self, x, y):
        return lib().orc_otw_acc(self._h, x, y)


class OnlineTimeWarping(_Stepper):
    """otw_eran.py:5-239 (insert path)."""

    def __init__(self, ref, params):
        _Stepper.__init__(self, KIND_OTW, ref, params["c"], params["max_run_count"])


class LiveNoteV2(_Stepper):
    """livenote_v2.py:3-236 (insert path)."""

    def __init__(self, ref, params, debug_params=None, chroma_diff=False):
        _Stepper.__init__(self, KIND_LN2, ref, params["search_band_width"], params["max_run_count"], 1 if chroma_diff else 0)


class LiveNote(_Stepper):
    """livenote.py (v1)."""

    def __init__(self, ref, params, debug_params=None):
        _Stepper.__init__(self, KIND_LN1, ref, params["search_band_width"], params["max_run_count"])


# ---------------------------------------------------------------- wtw.py
class WTW(object):
    """wtw.py:19-240.  `ref_recording` may be a WAV path (as in the reference)
    or an ndarray of samples."""

    def __init__(self, ref_recording, params, debug_params=None):
        if isinstance(ref_recording, str):
            ref, sr = _lr.load(ref_recording)
            assert sr == 22050
        else:
            ref = np.asarray(ref_recording)
        self.fft_len = params["fft_len"]
        self.hop_size = params["hop_size"]
        self.W = params["dtw_win_size"] // self.hop_size
        self.h = params["dtw_hop_size"] // self.hop_size
        assert self.fft_len == fft_len and self.hop_size == hop_size
        self.chroma_ref = wav_samples_to_chroma(ref)
        self._init_from_chroma(self.chroma_ref)

    @classmethod
    def from_chroma(cls, chroma_ref, W, h):
        self = cls.__new__(cls)
        self.fft_len, self.hop_size, self.W, self.h = fft_len, hop_size, W, h
        self.chroma_ref = _c64(chroma_ref)
        self._init_from_chroma(self.chroma_ref)
        return self

    def _init_from_chroma(self, cr):
        cr = _c64(cr)
        self._h = lib().orc_wtw_create(_p(cr), cr.shape[0], cr.shape[1], self.W, self.h)
        self.buf = []

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_wtw_destroy(self._h)
            self._h = None

    def insert(self, live_audio_buf):
        self.buf += list(live_audio_buf)
        if lib().orc_wtw_should_stop(self._h):
            return "stop"
        while len(self.buf) >= self.fft_len:
            col = wav_to_chroma_col(self.buf[: self.fft_len])
            self.buf = self.buf[self.hop_size :]
            if self.insert_chroma(col) == "stop":
                return "stop"
        return None

    def insert_chroma(self, col):
        col = _c64(col).reshape(-1)
        r = lib().orc_wtw_push_chroma(self._h, _p(col))
        return "stop" if r == 1 else None

    @property
    def path(self):
        n = lib().orc_wtw_path_len(self._h)
        if not n:
            return []
        arr = np.ctypeslib.as_array(lib().orc_wtw_path(self._h), shape=(n, 2))
        return [(int(x), int(y)) for x, y in arr]

    @property
    def live_ptr(self):
        return lib().orc_wtw_live_ptr(self._h)

    @property
    def ref_ptr(self):
        return lib().orc_wtw_ref_ptr(self._h)
