"""TEST INFRASTRUCTURE — numpy restatement of the three librosa functions whose
arithmetic sits on the reference's hot path.  librosa itself is NOT vendored in
the reference and its version is unpinned (no requirements file; era ≈ 0.5–0.6).
The published algorithm is restated here; call sites in the reference:

* ``librosa.load(path)``                  chroma.py:27,79  wtw.py:23
* ``librosa.filters.chroma(fs, fft_len)`` chroma.py:69     wtw.py:39
* ``librosa.util.normalize(S, norm=2, axis=0)``  chroma.py:74  wtw.py:41,90

Parity status: the filterbank formula is bit-identical to the independent port
in ``transformers.audio_utils.chroma_filter_bank`` (checked in
tests/test_oracle_cpu.py when transformers is importable) and the whole chain is
pinned end-to-end by the reference's own golden file
``Songs/chopin/tests/wtw_test_20b.txt`` (509/509 points).
"""
import wave

import numpy as np


def load(path, sr=22050):
    """librosa.load defaults: mono float32 at 22 050 Hz.  Only the no-resample
    case is supported (both WAVs present in the reference are 22 050 Hz)."""
    with wave.open(path, "rb") as w:
        rate = w.getframerate()
        nch = w.getnchannels()
        width = w.getsampwidth()
        raw = w.readframes(w.getnframes())
    if rate != sr:
        raise ValueError("resampling not restated (native rate %d)" % rate)
    if width != 2:
        raise ValueError("only int16 PCM restated")
    x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / np.float32(32768.0)
    if nch > 1:
        x = x.reshape(-1, nch).mean(axis=1, dtype=np.float32)
    return np.ascontiguousarray(x, dtype=np.float32), sr


def filters_chroma(sr, n_fft, n_chroma=12, A440=440.0, ctroct=5.0, octwidth=2.0, base_c=True):
    """librosa.filters.chroma with its defaults (norm=2) -> (n_chroma, 1 + n_fft//2) float64."""
    freqs = np.linspace(0, sr, n_fft, endpoint=False)[1:]
    frqbins = n_chroma * np.log2(freqs / (float(A440) / 16.0))
    frqbins = np.concatenate(([frqbins[0] - 1.5 * n_chroma], frqbins))
    binwidth = np.concatenate((np.maximum(frqbins[1:] - frqbins[:-1], 1.0), [1.0]))
    D = np.subtract.outer(frqbins, np.arange(0, n_chroma, dtype="d")).T
    half = np.round(float(n_chroma) / 2)
    D = np.remainder(D + half + 10 * n_chroma, n_chroma) - half
    wts = np.exp(-0.5 * (2 * D / np.tile(binwidth, (n_chroma, 1))) ** 2)
    wts = wts / np.sqrt(np.sum(wts ** 2, axis=0, keepdims=True))  # norm=2 per column
    wts *= np.tile(np.exp(-0.5 * (((frqbins / n_chroma - ctroct) / octwidth) ** 2)), (n_chroma, 1))
    if base_c:
        wts = np.roll(wts, -3 * (n_chroma // 12), axis=0)
    return np.ascontiguousarray(wts[:, : int(1 + n_fft / 2)])


def util_normalize(S, norm=2, axis=0):
    """librosa.util.normalize(norm=2): divide by the L2 length; lengths below
    ``finfo.tiny`` are replaced by 1 (silent frames stay zero, never NaN)."""
    assert norm == 2
    S = np.asarray(S, dtype=float)
    mag = np.abs(S)
    length = np.sum(mag ** 2, axis=axis, keepdims=True) ** 0.5
    length[length < np.finfo(S.dtype).tiny] = 1.0
    return S / length
