"""Generate the committed golden vectors by running the UNMODIFIED reference
(/root/reference, through oracle/ref_shim.py) in the build container.

    python tests/golden/make_golden.py

Outputs (all under tests/golden/):
  chopin_chroma.npz     reference chroma of the two WAVs present in the reference
                        (rubinstein_20b -> ref (12,380), rachmaninoff_20b -> live (12,401)),
                        their chroma-diff features, and the per-frame live chroma + ref
                        chroma the reference's WTW object computed (wtw.py:37-41,82-93)
  chopin_paths.npz      reference outputs on that pair: DTW path + acc_end, OTW / LiveNote
                        / LiveNoteV2 insert-loop paths (c = 10, 50), set_live paths,
                        LiveNoteV2 chroma_diff path, WTW path (== the reference's own
                        golden file Songs/chopin/tests/wtw_test_20b.txt, also stored)
  synth_cases.npz       seeded synthetic chroma pairs + reference DTW / OTW / LiveNoteV2 /
                        WTW outputs, including an exact-arithmetic tie case and c = 500
  audio_15s.npz         first 15 s of both WAVs as int16 stereo (exactly what librosa.load
                        mixes down) + reference wav_to_chroma of that excerpt and
                        wav_to_chroma_col of a few frames
  wtw_test_20b.txt      verbatim copy of the reference's golden path log (data, not code)
"""
import os
import shutil
import sys
import wave

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle import librosa_restated as lr  # noqa: E402

REF_WAV = "chopin/chopin_rubinstein_20b.wav"
LIVE_WAV = "chopin/chopin_rachmaninoff_20b.wav"


def insert_loop(obj, live):
    for i in range(live.shape[1]):
        if obj.insert(live[:, i]) == "stop":
            break
    return np.asarray(obj.path, dtype=np.int64).reshape(-1, 2)


def chroma_like(rng, n, smooth=0.7):
    x = rng.random((12, n))
    for k in range(1, n):
        x[:, k] = smooth * x[:, k - 1] + (1 - smooth) * x[:, k]
    return x / np.linalg.norm(x, axis=0)


def warped_copy(rng, ref, m, noise=0.05):
    n = ref.shape[1]
    u = np.linspace(0, 1, m)
    pos = np.clip((u + 0.08 * np.sin(6 * np.pi * u)) * (n - 1), 0, n - 1)
    y = ref[:, np.round(pos).astype(int)] + noise * rng.random((12, m))
    return y / np.linalg.norm(y, axis=0)


def read_stereo_i16(path, seconds):
    with wave.open(path, "rb") as w:
        assert w.getframerate() == 22050 and w.getsampwidth() == 2
        nch = w.getnchannels()
        raw = w.readframes(int(seconds * 22050))
    return np.frombuffer(raw, dtype="<i2").reshape(-1, nch).copy()


def main():
    assert ref_shim.available(), "reference tree not found"
    ch = ref_shim.load("chroma")
    dtw = ref_shim.load("dtw")
    otw = ref_shim.load("otw_eran")
    ln1 = ref_shim.load("livenote")
    ln2 = ref_shim.load("livenote_v2")
    wtw = ref_shim.load("wtw")
    ref_path, live_path = ref_shim.song(REF_WAV), ref_shim.song(LIVE_WAV)

    # ---------------- chopin pair ----------------
    ref = ch.wav_to_chroma(ref_path)
    live = ch.wav_to_chroma(live_path)
    dref = ch.wav_to_chroma_diff(ref_path)
    dlive = ch.wav_to_chroma_diff(live_path)
    out = {}
    cost, acc, path = dtw.DTW(live, ref)
    out["dtw_path"] = path.astype(np.int64)
    out["dtw_acc_end"] = np.array(acc[-1, -1])
    out["dtw_acc_row_last"] = acc[-1].copy()
    for c in (10, 50):
        out["otw_c%d" % c] = insert_loop(otw.OnlineTimeWarping(ref, {"c": c, "max_run_count": 3}), live)
        out["ln1_c%d" % c] = insert_loop(ln1.LiveNote(ref, {"search_band_width": c, "max_run_count": 3}, {}), live)
        out["ln2_c%d" % c] = insert_loop(ln2.LiveNoteV2(ref, {"search_band_width": c, "max_run_count": 3}, {}), live)
        o = otw.OnlineTimeWarping(ref, {"c": c, "max_run_count": 3})
        o.set_live(live)
        out["otw_setlive_c%d" % c] = np.asarray(o.path, dtype=np.int64)
        o = ln2.LiveNoteV2(ref, {"search_band_width": c, "max_run_count": 3}, {})
        o.set_live(live)
        out["ln2_setlive_c%d" % c] = np.asarray(o.path, dtype=np.int64)
    out["ln2_diff_c50"] = insert_loop(ln2.LiveNoteV2(dref, {"search_band_width": 50, "max_run_count": 3}, {}, chroma_diff=True), dlive)
    # WTW exactly as test_simple.py:168-182
    params = {"fft_len": 4096, "hop_size": 2048, "dtw_win_size": 4096 * 10, "dtw_hop_size": 2048 * 10}
    dbg = {"chroma": False, "song": False, "error": True, "error_detail": False, "alg": False}
    w = wtw.WTW(ref_path, params, dbg)
    x, _ = lr.load(live_path)
    for buf in np.array_split(x, 4096):
        if w.insert(buf.tolist()) == "stop":
            break
    out["wtw_path"] = np.asarray(w.path, dtype=np.int64)
    gold_txt = ref_shim.song("chopin/tests/wtw_test_20b.txt")
    gold = np.array([tuple(map(int, l.split())) for l in open(gold_txt) if l.strip()], dtype=np.int64)
    assert np.array_equal(out["wtw_path"], gold), "shim run does not reproduce the reference's own golden file"
    shutil.copyfile(gold_txt, os.path.join(HERE, "wtw_test_20b.txt"))
    # second WTW configuration (the live app's ratio, scaled: W = 40, h = 20)
    params2 = dict(params, dtw_win_size=4096 * 20, dtw_hop_size=2048 * 20)
    w2 = wtw.WTW(ref_path, params2, dbg)
    for buf in np.array_split(x, 4096):
        if w2.insert(buf.tolist()) == "stop":
            break
    out["wtw_path_w40"] = np.asarray(w2.path, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "chopin_paths.npz"), **out)
    np.savez_compressed(os.path.join(HERE, "chopin_chroma.npz"), ref=ref, live=live, dref=dref, dlive=dlive,
                        wtw_ref=w.chroma_ref, wtw_live=w.chroma_live[:, : w.chroma_ptr].copy(),
                        wtw_live_w40=w2.chroma_live[:, : w2.chroma_ptr].copy())

    # ---------------- synthetic cases ----------------
    syn = {}
    rng = np.random.default_rng(20251018)
    r = chroma_like(rng, 417)
    l = warped_copy(rng, r, 400)
    syn["a_ref"], syn["a_live"] = r, l
    c_, a_, p_ = dtw.DTW(l, r)
    syn["a_dtw_path"], syn["a_dtw_acc"] = p_.astype(np.int64), a_
    for c in (10, 50):
        syn["a_otw_c%d" % c] = insert_loop(otw.OnlineTimeWarping(r, {"c": c, "max_run_count": 3}), l)
        syn["a_ln2_c%d" % c] = insert_loop(ln2.LiveNoteV2(r, {"search_band_width": c, "max_run_count": 3}, {}), l)
    syn["a_otw_c7_r2"] = insert_loop(otw.OnlineTimeWarping(r, {"c": 7, "max_run_count": 2}), l)
    syn["a_ln2_c33_r5"] = insert_loop(ln2.LiveNoteV2(r, {"search_band_width": 33, "max_run_count": 5}, {}), l)
    # live longer than the reference allows: runs into "stop"
    l_long = warped_copy(rng, r, 900)
    o = otw.OnlineTimeWarping(r, {"c": 20, "max_run_count": 3})
    syn["a_live_long"] = l_long
    syn["a_otw_long_c20"] = insert_loop(o, l_long)
    syn["a_otw_long_tj"] = np.array([o.t, o.j])
    # exact-arithmetic ties (entries are multiples of 2^-10)
    ea = rng.integers(0, 8, size=(12, 150)) / 1024.0
    eb = rng.integers(0, 8, size=(12, 131)) / 1024.0
    syn["e_a"], syn["e_b"] = ea, eb
    c_, a_, p_ = dtw.DTW(ea, eb)
    syn["e_dtw_path"], syn["e_dtw_acc"] = p_.astype(np.int64), a_
    syn["e_otw_c10"] = insert_loop(otw.OnlineTimeWarping(eb, {"c": 10, "max_run_count": 3}), ea)
    syn["e_ln2_c10"] = insert_loop(ln2.LiveNoteV2(eb, {"search_band_width": 10, "max_run_count": 3}, {}), ea)
    # silence
    c_, a_, p_ = dtw.DTW(np.zeros((12, 4)), np.zeros((12, 5)))
    syn["z_dtw_path"], syn["z_dtw_acc"] = p_.astype(np.int64), a_
    # c = 500 (BASELINE cfg[3] band) on a 1300-frame reference
    r5 = chroma_like(rng, 1300)
    l5 = warped_copy(rng, r5, 1400)
    syn["c500_ref"], syn["c500_live"] = r5, l5
    syn["c500_otw"] = insert_loop(otw.OnlineTimeWarping(r5, {"c": 500, "max_run_count": 3}), l5)
    syn["c500_ln2"] = insert_loop(ln2.LiveNoteV2(r5, {"search_band_width": 500, "max_run_count": 3}, {}), l5)
    # WTW on synthetic chroma columns (bypasses audio): W = 20, h = 10
    w3 = wtw.WTW.__new__(wtw.WTW)
    w3.fft_len, w3.hop_size, w3.dtw_win_size, w3.dtw_hop_size = 4096, 2048, 4096 * 10, 2048 * 10
    w3.chroma_info = False
    w3.chroma_ref = r
    w3.N, w3.M = 2 * r.shape[1], r.shape[1]
    w3.chroma_live = np.zeros((12, w3.N))
    w3.acc_cost = np.full((w3.N, w3.M), np.inf)
    w3.buf, w3.path, w3.windows = [], [], []
    w3.chroma_ptr = w3.live_ptr = w3.ref_ptr = 0
    W = 20
    for k in range(l.shape[1]):
        # the body of the reference's per-frame loop, wtw.py:91-128, fed a ready chroma column
        if w3.ref_ptr >= w3.M - 1 or w3.live_ptr >= w3.N - 1:
            break
        w3.chroma_live[:, w3.chroma_ptr] = l[:, k]
        w3.chroma_ptr += 1
        if w3.ref_ptr >= (w3.M - 1 - W) or w3.live_ptr >= (w3.N - 1 - W):
            break
        while w3.chroma_ptr - w3.live_ptr >= W:
            cx = w3.chroma_live[:, w3.live_ptr : w3.live_ptr + W]
            cy = w3.chroma_ref[:, w3.ref_ptr : w3.ref_ptr + W]
            D, B = w3.run_dtw(w3.get_cost_matrix(cx, cy))
            sub = w3.find_path(B)
            change, index = False, None
            for i in range(len(sub)):
                if sub[i][0] <= 10:
                    w3.path.append((sub[i][0] + w3.live_ptr, sub[i][1] + w3.ref_ptr))
                else:
                    change, index = True, i - 1
                    break
            if change:
                w3.live_ptr, w3.ref_ptr = sub[index][0] + w3.live_ptr, sub[index][1] + w3.ref_ptr
            else:
                w3.live_ptr, w3.ref_ptr = w3.live_ptr + 10, w3.ref_ptr + 10
    syn["a_wtw_w20"] = np.asarray(w3.path, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "synth_cases.npz"), **syn)

    # ---------------- audio excerpt ----------------
    aud = {}
    for tag, p in (("ref", ref_path), ("live", live_path)):
        st = read_stereo_i16(p, 15.0)
        aud[tag + "_i16"] = st
        mono = (st.astype(np.float32) / np.float32(32768.0)).mean(axis=1, dtype=np.float32)
        full, _ = lr.load(p)
        assert np.array_equal(mono, full[: len(mono)])
        stft = ch.create_stft(mono)
        aud[tag + "_chroma"] = ch.create_chroma(stft)
        aud[tag + "_raw_chroma"] = ch.create_chroma(stft, normalize=False)
        cols = [ch.wav_to_chroma_col(mono[s : s + 4096]) for s in (0, 2048, 50000, 200000)]
        aud[tag + "_cols"] = np.stack(cols)
    np.savez_compressed(os.path.join(HERE, "audio_15s.npz"), **aud)
    for f in sorted(os.listdir(HERE)):
        print("%-24s %8d bytes" % (f, os.path.getsize(os.path.join(HERE, f))))


if __name__ == "__main__" and "--eval-only" not in sys.argv:
    main()


def make_eval_golden():
    """Score the reference's own DTW path of the chopin_20b pair with the reference's own scorer
    (tests.py:29-137, class test_simple), loaded through the shim up to the point where the script
    starts evaluating at import time."""
    import io, re, types, contextlib
    src = open(os.path.join(ref_shim.REFERENCE_ROOT, "tests.py")).read()
    src = src[: src.index("params = {")]
    src = src[src.index("def lines_from_file"):]
    src = re.sub(r"^(\s*)print (.+)$", r"\1print(\2)", src, flags=re.M)
    mod = types.ModuleType("_ref_tests_head")
    mod.__dict__["csv"] = __import__("csv")
    exec(compile(src, "tests.py", "exec"), mod.__dict__)
    paths = np.load(os.path.join(HERE, "chopin_paths.npz"))
    ref_wav, live_wav = ref_shim.song(REF_WAV), ref_shim.song(LIVE_WAV)
    out = {}
    cases = {k: paths[k] for k in ("dtw_path", "otw_c50", "ln2_c50", "wtw_path")}
    shifted = paths["dtw_path"].copy()
    shifted[:, 1] = np.clip(shifted[:, 1] - 45, 0, None)        # a deliberately bad path (ref lags ~4 s)
    cases["dtw_shifted45"] = shifted
    out["shift"] = 45
    for key, pth in cases.items():
        scorer = mod.test_simple(ref_wav, live_wav, [tuple(p) for p in pth.tolist()])
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            ret = scorer.get_error()
        printed = [float(l.split(":")[1].strip().split()[0]) for l in buf.getvalue().splitlines() if l.startswith("Percent incorrect")]
        out[key] = {"returned": ret, "printed": printed}
    # the two ground-truth CSVs are data fixtures (19 rows each)
    for name in ("chopin_rubinstein_20b.csv", "chopin_rachmaninoff_20b.csv"):
        shutil.copyfile(ref_shim.song("chopin/" + name), os.path.join(HERE, name))
    # and one field log as a format fixture (tests.py:20-27 reader)
    log = sorted(f for f in os.listdir(os.path.join(ref_shim.REFERENCE_ROOT, "tests")) if f.startswith("livenote_test_live_"))
    sizes = [(os.path.getsize(os.path.join(ref_shim.REFERENCE_ROOT, "tests", f)), f) for f in log]
    sizes = [x for x in sizes if x[0] > 200]
    pick = min(sizes)[1]
    shutil.copyfile(os.path.join(ref_shim.REFERENCE_ROOT, "tests", pick), os.path.join(HERE, "field_log_sample.txt"))
    out["field_log_points"] = len(mod.data_from_file(os.path.join(HERE, "field_log_sample.txt")))
    out["field_log_first"] = list(mod.data_from_file(os.path.join(HERE, "field_log_sample.txt"))[0])
    import json
    with open(os.path.join(HERE, "eval_golden.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(out)


if __name__ == "__main__":
    make_eval_golden()
