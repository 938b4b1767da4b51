"""CPU suite: the oracle (C restatement + numpy chroma) against the golden vectors the
reference itself produced (tests/golden/make_golden.py), and against the live reference
when /root/reference is present (build container only)."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def chroma():
    return np.load(os.path.join(GOLD, "chopin_chroma.npz"))


@pytest.fixture(scope="module")
def paths():
    return np.load(os.path.join(GOLD, "chopin_paths.npz"))


@pytest.fixture(scope="module")
def syn():
    return np.load(os.path.join(GOLD, "synth_cases.npz"))


def run_insert(obj, live):
    for i in range(live.shape[1]):
        if obj.insert(live[:, i]) == "stop":
            break
    return np.asarray(obj.path, dtype=np.int64).reshape(-1, 2)


def test_dtw_chopin(orc, chroma, paths):
    cost, acc, path = orc.DTW(chroma["live"], chroma["ref"])
    assert np.array_equal(path, paths["dtw_path"])
    # numpy's dgemm rounds a few tail columns differently (<= 2 ulp of cost); paths are identical
    assert abs(acc[-1, -1] - float(paths["dtw_acc_end"])) < 1e-10
    assert np.allclose(acc[-1], paths["dtw_acc_row_last"], rtol=0, atol=1e-10)


def test_dtw_synthetic(orc, syn):
    cost, acc, path = orc.DTW(syn["a_live"], syn["a_ref"])
    assert np.array_equal(path, syn["a_dtw_path"])
    assert np.allclose(acc, syn["a_dtw_acc"], rtol=0, atol=1e-10)
    cost, acc, path = orc.DTW(syn["e_a"], syn["e_b"])
    assert np.array_equal(path, syn["e_dtw_path"]) and np.array_equal(acc, syn["e_dtw_acc"])   # exact arithmetic
    cost, acc, path = orc.DTW(np.zeros((12, 4)), np.zeros((12, 5)))
    assert np.array_equal(path, syn["z_dtw_path"]) and np.array_equal(acc, syn["z_dtw_acc"])


@pytest.mark.parametrize("c", [10, 50])
def test_otw_family_chopin(orc, chroma, paths, c):
    ref, live = chroma["ref"], chroma["live"]
    assert np.array_equal(run_insert(orc.OnlineTimeWarping(ref, {"c": c, "max_run_count": 3}), live), paths["otw_c%d" % c])
    assert np.array_equal(run_insert(orc.LiveNote(ref, {"search_band_width": c, "max_run_count": 3}), live), paths["ln1_c%d" % c])
    assert np.array_equal(run_insert(orc.LiveNoteV2(ref, {"search_band_width": c, "max_run_count": 3}), live), paths["ln2_c%d" % c])


def test_livenote_v2_chroma_diff(orc, chroma, paths):
    got = run_insert(orc.LiveNoteV2(chroma["dref"], {"search_band_width": 50, "max_run_count": 3}, chroma_diff=True), chroma["dlive"])
    assert np.array_equal(got, paths["ln2_diff_c50"])


def test_otw_family_synthetic(orc, syn):
    r, l = syn["a_ref"], syn["a_live"]
    for c in (10, 50):
        assert np.array_equal(run_insert(orc.OnlineTimeWarping(r, {"c": c, "max_run_count": 3}), l), syn["a_otw_c%d" % c])
        assert np.array_equal(run_insert(orc.LiveNoteV2(r, {"search_band_width": c, "max_run_count": 3}), l), syn["a_ln2_c%d" % c])
    assert np.array_equal(run_insert(orc.OnlineTimeWarping(r, {"c": 7, "max_run_count": 2}), l), syn["a_otw_c7_r2"])
    assert np.array_equal(run_insert(orc.LiveNoteV2(r, {"search_band_width": 33, "max_run_count": 5}), l), syn["a_ln2_c33_r5"])
    o = orc.OnlineTimeWarping(r, {"c": 20, "max_run_count": 3})
    assert np.array_equal(run_insert(o, syn["a_live_long"]), syn["a_otw_long_c20"])
    assert [o.t, o.j] == syn["a_otw_long_tj"].tolist()
    assert np.array_equal(run_insert(orc.OnlineTimeWarping(syn["e_b"], {"c": 10, "max_run_count": 3}), syn["e_a"]), syn["e_otw_c10"])
    assert np.array_equal(run_insert(orc.LiveNoteV2(syn["e_b"], {"search_band_width": 10, "max_run_count": 3}), syn["e_a"]), syn["e_ln2_c10"])


def test_otw_c500(orc, syn):
    r, l = syn["c500_ref"], syn["c500_live"]
    assert np.array_equal(run_insert(orc.OnlineTimeWarping(r, {"c": 500, "max_run_count": 3}), l), syn["c500_otw"])
    assert np.array_equal(run_insert(orc.LiveNoteV2(r, {"search_band_width": 500, "max_run_count": 3}), l), syn["c500_ln2"])


def test_wtw_golden_file_from_reference_chroma(orc, chroma, paths):
    """The reference's own golden log (Songs/chopin/tests/wtw_test_20b.txt) from the chroma
    columns the reference's WTW object computed."""
    gold = np.array([tuple(map(int, l.split())) for l in open(os.path.join(GOLD, "wtw_test_20b.txt")) if l.strip()])
    assert gold.shape == (509, 2) and np.array_equal(gold, paths["wtw_path"])
    w = orc.WTW.from_chroma(chroma["wtw_ref"], 20, 10)
    for k in range(chroma["wtw_live"].shape[1]):
        if w.insert_chroma(chroma["wtw_live"][:, k]) == "stop":
            break
    assert np.array_equal(np.asarray(w.path), gold)
    w = orc.WTW.from_chroma(chroma["wtw_ref"], 40, 20)
    for k in range(chroma["wtw_live_w40"].shape[1]):
        if w.insert_chroma(chroma["wtw_live_w40"][:, k]) == "stop":
            break
    assert np.array_equal(np.asarray(w.path), paths["wtw_path_w40"])


def test_wtw_synthetic(orc, syn):
    w = orc.WTW.from_chroma(syn["a_ref"], 20, 10)
    for k in range(syn["a_live"].shape[1]):
        if w.insert_chroma(syn["a_live"][:, k]) == "stop":
            break
    assert np.array_equal(np.asarray(w.path), syn["a_wtw_w20"])


def test_chroma_oracle_audio(orc):
    aud = np.load(os.path.join(GOLD, "audio_15s.npz"))
    for tag in ("ref", "live"):
        st = aud[tag + "_i16"]
        mono = (st.astype(np.float32) / np.float32(32768.0)).mean(axis=1, dtype=np.float32)
        got = orc.wav_samples_to_chroma(mono)
        assert got.shape == aud[tag + "_chroma"].shape
        assert np.abs(got - aud[tag + "_chroma"]).max() < 1e-12
        for k, s in enumerate((0, 2048, 50000, 200000)):
            assert np.abs(orc.wav_to_chroma_col(mono[s : s + 4096]) - aud[tag + "_cols"][k]).max() < 1e-12


def test_filterbank_matches_independent_port():
    from oracle import librosa_restated as lr
    try:
        from transformers.audio_utils import chroma_filter_bank
    except Exception:
        pytest.skip("transformers not importable")
    assert np.array_equal(lr.filters_chroma(22050, 4096), np.asarray(chroma_filter_bank(4096, 12, 22050)))


# ---------------------------------------------------------------- live reference (build container only)
def _shim():
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("/root/reference not present (GPU box)")
    return ref_shim


def test_oracle_vs_live_reference_dtw(orc):
    shim = _shim()
    dtw = shim.load("dtw")
    rng = np.random.default_rng(42)
    for M, N in [(64, 64), (120, 96), (200, 184)]:     # N multiple of 8: numpy's dgemm == fma chain everywhere
        a, b = rng.random((12, M)), rng.random((12, N))
        c, acc, p = dtw.DTW(a, b)
        c2, acc2, p2 = orc.DTW(a, b)
        assert np.array_equal(c, c2) and np.array_equal(acc, acc2) and np.array_equal(p, p2)
    a, b = rng.random((12, 77)), rng.random((12, 93))    # tail columns may differ by an ulp of cost
    c, acc, p = dtw.DTW(a, b)
    c2, acc2, p2 = orc.DTW(a, b)
    assert np.array_equal(p, p2) and np.allclose(acc, acc2, rtol=0, atol=1e-11)


def test_oracle_vs_live_reference_streams(orc):
    shim = _shim()
    otw, ln2 = shim.load("otw_eran"), shim.load("livenote_v2")
    rng = np.random.default_rng(43)
    r = rng.random((12, 150))
    r /= np.linalg.norm(r, axis=0)
    l = rng.random((12, 170))
    l /= np.linalg.norm(l, axis=0)
    for c, mr in [(5, 3), (16, 2), (40, 4)]:
        ro = otw.OnlineTimeWarping(r, {"c": c, "max_run_count": mr})
        oo = orc.OnlineTimeWarping(r, {"c": c, "max_run_count": mr})
        assert np.array_equal(run_insert(ro, l), run_insert(oo, l))
        assert all(ro.acc_cost[x, y] == oo.acc(x, y) for x in range(0, 170, 3) for y in range(0, 150, 3))
        rl = ln2.LiveNoteV2(r, {"search_band_width": c, "max_run_count": mr}, {})
        ol = orc.LiveNoteV2(r, {"search_band_width": c, "max_run_count": mr})
        assert np.array_equal(run_insert(rl, l), run_insert(ol, l))


def test_oracle_vs_live_reference_wtw_end_to_end(orc):
    """Audio in, path out: the reference's own golden file through the oracle's chroma + WTW."""
    shim = _shim()
    from oracle import librosa_restated as lr
    params = {"fft_len": 4096, "hop_size": 2048, "dtw_win_size": 4096 * 10, "dtw_hop_size": 2048 * 10}
    w = orc.WTW(shim.song("chopin/chopin_rubinstein_20b.wav"), params, {"chroma": False})
    x, _ = lr.load(shim.song("chopin/chopin_rachmaninoff_20b.wav"))
    for buf in np.array_split(x, 4096):
        if w.insert(buf.tolist()) == "stop":
            break
    gold = [tuple(map(int, l.split())) for l in open(shim.song("chopin/tests/wtw_test_20b.txt")) if l.strip()]
    assert w.path == gold


# ---------------------------------------------------------------- evaluation glue, CPU only
def test_beat_scorer_matches_reference_scorer(entry, paths):
    import json
    ev = entry.submodule("evalutil")
    gold = json.load(open(os.path.join(GOLD, "eval_golden.json")))
    sc = ev.BeatScorer(os.path.join(GOLD, "chopin_rubinstein_20b.csv"), os.path.join(GOLD, "chopin_rachmaninoff_20b.csv"))
    cases = {k: paths[k] for k in ("dtw_path", "otw_c50", "ln2_c50", "wtw_path")}
    shifted = paths["dtw_path"].copy()
    shifted[:, 1] = np.clip(shifted[:, 1] - gold["shift"], 0, None)
    cases["dtw_shifted45"] = shifted
    for key, pth in cases.items():
        got = sc.score([tuple(p) for p in pth.tolist()])
        want = gold[key]["printed"]
        mine = [got["pct_off_%d_beats" % t] for t in (1, 3, 5, 10)] + [got["pct_off_%d_secs" % t] for t in (1, 3, 5, 10)]
        assert np.allclose(mine, want, rtol=0, atol=1e-12), key
        assert abs(sc.get_error([tuple(p) for p in pth.tolist()]) - gold[key]["returned"]) < 1e-12
    assert gold["dtw_shifted45"]["returned"] > 0          # the bad path really scores badly


def test_path_log_round_trip(entry, tmp_path, paths):
    import json
    ev = entry.submodule("evalutil")
    gold = json.load(open(os.path.join(GOLD, "eval_golden.json")))
    log = ev.read_path_log(os.path.join(GOLD, "field_log_sample.txt"))       # a log the reference's live app wrote
    assert len(log) == gold["field_log_points"] and list(log[0]) == gold["field_log_first"]
    out = os.path.join(str(tmp_path), "log.txt")
    pth = [tuple(p) for p in paths["otw_c50"].tolist()]
    ev.write_path_log(out, "Songs/chopin/chopin_rubinstein_20b.wav", 4096, 2048, {"search_band_width": 50, "max_run_count": 3}, pth)
    raw = open(out, "rb").read()
    assert raw.count(b"\r\n") == 5 + len(pth) and raw.startswith(b"Songs/chopin/chopin_rubinstein_20b.wav\r\nfft_len: 4096\r\n")
    assert ev.read_path_log(out) == pth
