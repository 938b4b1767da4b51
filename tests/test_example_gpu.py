"""GPU: the end-to-end example (examples/align_pair.py = the flow of the reference's test_simple.py) on two synthetic
PCM16 WAV files, every aligner: the path must start at the origin, be monotone, and end near the end of both recordings;
the DTW path written to the reference's log format must read back identically."""
import os
import subprocess
import sys
import wave

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def melody(notes, dur, sr=22050, seed=0):
    """A sequence of harmonic tones (one per note), `dur[k]` seconds each."""
    rng = np.random.default_rng(seed)
    out = []
    for f, d in zip(notes, dur):
        t = np.arange(int(d * sr)) / sr
        x = sum(a * np.sin(2 * np.pi * f * h * t) for h, a in ((1, 1.0), (2, 0.5), (3, 0.25)))
        out.append(x * np.hanning(len(t)) ** 0.25)
    x = np.concatenate(out)
    x = 0.3 * x / np.abs(x).max() + 0.002 * rng.standard_normal(len(x))
    return (x * 32767).astype("<i2")


def write_wav(path, pcm):
    with wave.open(path, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(22050)
        w.writeframes(pcm.tobytes())


@pytest.fixture(scope="module")
def wavs(tmp_path_factory):
    d = tmp_path_factory.mktemp("wavs")
    rng = np.random.default_rng(5)
    notes = 220.0 * 2 ** (rng.integers(0, 24, size=40) / 12.0)
    ref_d = np.full(40, 0.5)
    live_d = ref_d * (1.0 + 0.3 * np.sin(np.linspace(0, 3, 40)))          # tempo drifts up to +30 %
    ref, live = os.path.join(str(d), "ref.wav"), os.path.join(str(d), "live.wav")
    write_wav(ref, melody(notes, ref_d, seed=1))
    write_wav(live, melody(notes, live_d, seed=2))
    return ref, live, str(d)


@pytest.mark.parametrize("method", ["dtw", "otw", "livenote_v2", "livenote", "wtw"])
def test_example_aligns_two_wavs(wavs, method):
    ref, live, d = wavs
    log = os.path.join(d, "path_%s.txt" % method)
    cmd = [sys.executable, os.path.join(ROOT, "examples", "align_pair.py"), ref, live, "--method", method, "--c", "60", "--log", log]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    ev = g.submodule("evalutil")
    path = np.array(ev.read_path_log(log))
    n_live = (len(np.frombuffer(open(live, "rb").read()[44:], dtype="<i2")) + 2048 - 4096) // 2048 + 1
    n_ref = (len(np.frombuffer(open(ref, "rb").read()[44:], dtype="<i2")) + 2048 - 4096) // 2048 + 1
    assert path[0][0] <= 1 and path[0][1] <= 1          # the online aligners log their first point after the first frame
    dlt = np.diff(path, axis=0)
    if method == "dtw":                       # unit steps (dtw.py:43-52)
        assert ((dlt >= 0) & (dlt <= 1)).all()
    elif method in ("livenote_v2", "wtw"):    # forward-only filter (livenote_v2.py:197-199) / stitched windows (wtw.py:113-127)
        assert (dlt >= 0).all()
    # OTW and LiveNote v1 log the best frontier point of every step: not necessarily monotone (otw_eran.py:153-188)
    # the aligners follow the tempo drift: the last point is close to the end of both recordings
    assert path[-1][1] >= 0.9 * (n_ref - 1), (path[-1], n_ref)
    assert path[-1][0] >= 0.8 * (n_live - 1), (path[-1], n_live)
    # on the diagonal of the true correspondence: the middle note of the live recording maps to the middle note of the ref
    if method == "dtw":
        assert tuple(path[-1]) == (n_live - 1, n_ref - 1)
