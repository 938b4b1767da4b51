"""GPU parity: windowed time warping (kernel K6 + K1) through the drop-in WTW class."""
import os

import numpy as np
import pytest

from conftest import chroma_like, warped_copy

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def mods(entry):
    return {n: entry.submodule(n) for n in ("wtw", "batch", "chroma")}


def test_reference_golden_file_from_reference_chroma(mods):
    """Songs/chopin/tests/wtw_test_20b.txt (the reference's own golden log, 509 points),
    from the chroma columns the reference's WTW object computed."""
    ch = np.load(os.path.join(GOLD, "chopin_chroma.npz"))
    paths = np.load(os.path.join(GOLD, "chopin_paths.npz"))
    gold = np.array([tuple(map(int, l.split())) for l in open(os.path.join(GOLD, "wtw_test_20b.txt")) if l.strip()])
    for W, h, live, want in ((20, 10, ch["wtw_live"], gold), (40, 20, ch["wtw_live_w40"], paths["wtw_path_w40"])):
        b = mods["batch"].WtwBatch([ch["wtw_ref"]], W, h)
        st = b.push(np.ascontiguousarray(live.T).reshape(-1, 1, 12))
        assert np.array_equal(b.paths()[0], want), (W, h)
        # the reference stopped feeding at the first "stop"; every later push reports stop too
        first = int(np.argmax(st[:, 0] == 1)) if (st[:, 0] == 1).any() else None
        if first is not None:
            assert (st[first:, 0] == 1).all()
        b.close()


def test_synthetic_vs_reference_and_oracle(mods, orc):
    syn = np.load(os.path.join(GOLD, "synth_cases.npz"))
    b = mods["batch"].WtwBatch([syn["a_ref"]], 20, 10)
    b.push(np.ascontiguousarray(syn["a_live"].T).reshape(-1, 1, 12))
    assert np.array_equal(b.paths()[0], syn["a_wtw_w20"])
    b.close()
    # several streams, other window shapes, against the oracle column by column
    rng = np.random.default_rng(4)
    refs = [chroma_like(rng, n) for n in (150, 260, 90)]
    lives = [warped_copy(rng, r, 240) for r in refs]
    for W, h in ((8, 3), (32, 16), (50, 49), (12, 20)):
        b = mods["batch"].WtwBatch(refs, W, h)
        orcs = [orc.WTW.from_chroma(r, W, h) for r in refs]
        cols = np.stack([np.stack([lv[:, k] for lv in lives]) for k in range(240)])
        st = b.push(cols)
        got = b.paths()
        pos = b.positions()
        for s in range(3):
            stopped = False
            for k in range(240):
                r = orcs[s].insert_chroma(lives[s][:, k])
                assert (r == "stop") == (st[k, s] == 1), (W, h, s, k)
            assert np.array_equal(got[s], np.asarray(orcs[s].path, dtype=np.int64).reshape(-1, 2)), (W, h, s)
            assert pos[s, 1] == orcs[s].live_ptr and pos[s, 2] == orcs[s].ref_ptr
        b.close()


def test_audio_in_path_out_matches_oracle(mods, orc):
    """Drop-in WTW class on real audio (15 s excerpts of the two chopin recordings): the
    path must equal the oracle's, which is pinned to the reference end-to-end."""
    aud = np.load(os.path.join(GOLD, "audio_15s.npz"))
    mono = {t: (aud[t + "_i16"].astype(np.float32) / np.float32(32768.0)).mean(axis=1, dtype=np.float32) for t in ("ref", "live")}
    params = {"fft_len": 4096, "hop_size": 2048, "dtw_win_size": 4096 * 10, "dtw_hop_size": 2048 * 10}
    w = mods["wtw"].WTW(mono["ref"], params, {"chroma": False})
    o = orc.WTW(mono["ref"], params, {"chroma": False})
    assert np.abs(w.chroma_ref - o.chroma_ref).max() < 1e-9
    for buf in np.array_split(mono["live"], 1800):
        r1 = w.insert(buf.tolist())
        r2 = o.insert(buf.tolist())
        assert r1 == r2
        if r1 == "stop":
            break
    assert w.path == o.path and len(w.path) > 100
    assert (w.live_ptr, w.ref_ptr) == (o.live_ptr, o.ref_ptr)
    # an aligner that has reported "stop", fed several frames per call, stays in step with the reference (the kernel must not
    # consume the frames that follow a stop inside one launch: the reference returns there and re-reads them next time)
    for k in range(4):
        big = mono["live"][k * 3000 : k * 3000 + 9000].tolist()
        assert w.insert(big) == o.insert(big), k
        assert (w.live_ptr, w.ref_ptr) == (o.live_ptr, o.ref_ptr)
    assert w.path == o.path
