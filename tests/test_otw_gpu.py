"""GPU parity: OnlineTimeWarping / LiveNoteV2 / LiveNote through the C ABI (kernel K5)
vs the reference-generated golden paths and the CPU oracle (bit-exact)."""
import os

import numpy as np
import pytest

from conftest import chroma_like, warped_copy

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def mods(entry):
    return {n: entry.submodule(n) for n in ("otw_eran", "livenote_v2", "livenote", "batch", "_native")}


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


@pytest.mark.parametrize("c", [10, 50])
def test_chopin_insert_loops(mods, chroma, paths, c):
    ref, live = chroma["ref"], chroma["live"]
    o = mods["otw_eran"].OnlineTimeWarping(ref, {"c": c, "max_run_count": 3})
    assert np.array_equal(run_insert(o, live), paths["otw_c%d" % c])
    assert o.path[0] == (1, 0) and isinstance(o.path, list)
    o = mods["livenote_v2"].LiveNoteV2(ref, {"search_band_width": c, "max_run_count": 3}, {})
    assert np.array_equal(run_insert(o, live), paths["ln2_c%d" % c])
    o = mods["livenote"].LiveNote(ref, {"search_band_width": c, "max_run_count": 3}, {})
    assert np.array_equal(run_insert(o, live), paths["ln1_c%d" % c])


def test_chopin_chroma_diff(mods, chroma, paths):
    o = mods["livenote_v2"].LiveNoteV2(chroma["dref"], {"search_band_width": 50, "max_run_count": 3}, {}, chroma_diff=True)
    assert np.array_equal(run_insert(o, chroma["dlive"]), paths["ln2_diff_c50"])


@pytest.mark.parametrize("c", [10, 50])
def test_set_live(mods, chroma, paths, c):
    o = mods["otw_eran"].OnlineTimeWarping(chroma["ref"], {"c": c, "max_run_count": 3})
    o.set_live(chroma["live"])
    assert isinstance(o.path, np.ndarray) and np.array_equal(o.path, paths["otw_setlive_c%d" % c])
    o = mods["livenote_v2"].LiveNoteV2(chroma["ref"], {"search_band_width": c, "max_run_count": 3}, {})
    o.set_live(chroma["live"])
    assert np.array_equal(np.asarray(o.path), paths["ln2_setlive_c%d" % c])


def test_synthetic_goldens(mods, syn):
    r, l = syn["a_ref"], syn["a_live"]
    OTW, LN2 = mods["otw_eran"].OnlineTimeWarping, mods["livenote_v2"].LiveNoteV2
    for c in (10, 50):
        assert np.array_equal(run_insert(OTW(r, {"c": c, "max_run_count": 3}), l), syn["a_otw_c%d" % c])
        assert np.array_equal(run_insert(LN2(r, {"search_band_width": c, "max_run_count": 3}, {}), l), syn["a_ln2_c%d" % c])
    assert np.array_equal(run_insert(OTW(r, {"c": 7, "max_run_count": 2}), l), syn["a_otw_c7_r2"])
    assert np.array_equal(run_insert(LN2(r, {"search_band_width": 33, "max_run_count": 5}, {}), l), syn["a_ln2_c33_r5"])
    assert np.array_equal(run_insert(OTW(syn["e_b"], {"c": 10, "max_run_count": 3}), syn["e_a"]), syn["e_otw_c10"])
    assert np.array_equal(run_insert(LN2(syn["e_b"], {"search_band_width": 10, "max_run_count": 3}, {}), syn["e_a"]), syn["e_ln2_c10"])


def test_stop_and_positions(mods, syn):
    o = mods["otw_eran"].OnlineTimeWarping(syn["a_ref"], {"c": 20, "max_run_count": 3})
    live = syn["a_live_long"]
    stopped_at = None
    for i in range(live.shape[1]):
        if o.insert(live[:, i]) == "stop":
            stopped_at = i
            break
    assert stopped_at is not None
    assert np.array_equal(np.asarray(o.path), syn["a_otw_long_c20"])
    assert [o.t, o.j] == syn["a_otw_long_tj"].tolist()
    assert o.insert(live[:, 0]) == "stop"          # finished objects stay finished


def test_c500_batched_many_frames_per_launch(mods, syn):
    """BASELINE cfg[3] band (c = 500): whole live sequence in one launch, two kinds."""
    r, l = syn["c500_ref"], syn["c500_live"]
    import torch
    for kind, key in (("otw", "c500_otw"), ("livenote_v2", "c500_ln2")):
        b = mods["batch"].OtwBatch([r, r, r], 500, 3, kind=kind)
        frames = torch.from_numpy(np.ascontiguousarray(np.repeat(l.T[:, None, :], 3, axis=1))).cuda()
        st, npts, pts = b.step_device(frames)
        got = b.paths()
        for s in range(3):
            assert np.array_equal(got[s], syn[key]), (kind, s)
        # per-step outputs agree with the device-resident path
        npts = npts.cpu().numpy()
        pts = pts.cpu().numpy()
        flat = np.concatenate([pts[f, 0, : npts[f, 0]] for f in range(frames.shape[0])])
        assert np.array_equal(flat, syn[key])
        b.close()


def test_ragged_batch_vs_oracle(mods, orc):
    """Streams with different reference lengths and an inactive-stream mask."""
    rng = np.random.default_rng(9)
    refs = [chroma_like(rng, n) for n in (120, 333, 64, 250)]
    lives = [warped_copy(rng, r, 300) for r in refs]
    import torch
    b = mods["batch"].OtwBatch(refs, 25, 3, kind="otw")
    oracles = [orc.OnlineTimeWarping(r, {"c": 25, "max_run_count": 3}) for r in refs]
    done = [False] * 4
    active = torch.ones(4, dtype=torch.uint8, device="cuda")
    for k in range(300):
        if k == 100:
            active[1] = 0                      # pause stream 1 for 50 steps
        if k == 150:
            active[1] = 1
        fr = np.stack([lv[:, k] for lv in lives])
        st, _, _ = b.step_device(torch.from_numpy(fr).cuda().reshape(1, 4, 12), active=active)
        st = st.cpu().numpy()[0]
        for s in range(4):
            if s == 1 and 100 <= k < 150:
                continue
            if not done[s]:
                r = oracles[s].insert(lives[s][:, k])
                assert (r == "stop") == (st[s] == 1), (k, s)
                done[s] = r == "stop"
    got = b.paths()
    for s in range(4):
        assert np.array_equal(got[s], oracles[s].path_array()), s
    b.close()


@pytest.mark.parametrize("kind,c,mr", [("otw", 7, 2), ("otw", 40, 3), ("livenote_v2", 7, 3), ("livenote_v2", 33, 5),
                                        ("livenote", 12, 3), ("livenote_v2", 100, 2)])
def test_fuzz_many_ragged_streams(mods, orc, kind, c, mr):
    """48 streams with random reference lengths (some shorter than c, some one frame long), each fed a live sequence of
    different tempo until it stops or the live ends, in one batch: every stream's path equals the oracle's."""
    import torch
    rng = np.random.default_rng(c * 131 + mr)
    n = 48
    ref_lens = [int(x) for x in rng.integers(1, 260, size=n)]
    ref_lens[0], ref_lens[1], ref_lens[2] = 1, 2, c + 1
    refs = [chroma_like(rng, m) for m in ref_lens]
    T = 320
    lives = [warped_copy(rng, r, T) if r.shape[1] > 4 else chroma_like(rng, T) for r in refs]
    b = mods["batch"].OtwBatch(refs, c, mr, kind=kind)
    params = {"c": c, "max_run_count": mr, "search_band_width": c}      # the reference's two spellings of the band width
    ocls = {"otw": orc.OnlineTimeWarping, "livenote_v2": orc.LiveNoteV2, "livenote": orc.LiveNote}[kind]
    oracles = [ocls(r, dict(params)) for r in refs]
    done = [False] * n
    for k in range(T):
        fr = np.stack([lv[:, k] for lv in lives])
        st, _, _ = b.step_device(torch.from_numpy(fr).cuda().reshape(1, n, 12))
        st = st.cpu().numpy()[0]
        for s in range(n):
            if not done[s]:
                r = oracles[s].insert(lives[s][:, k])
                assert (r == "stop") == (st[s] == 1), (k, s, ref_lens[s])
                done[s] = r == "stop"
    got = b.paths()
    for s in range(n):
        assert np.array_equal(got[s], oracles[s].path_array()), (s, ref_lens[s])
    b.close()


def test_checkpoint_resume_of_stream_state(mods, orc):
    """export_state() after 90 frames, import_state() into a fresh batch, continue: identical to an uninterrupted run
    (paths, positions, stop flags) for OTW and WTW batches; a snapshot of a different configuration is refused."""
    import torch
    rng = np.random.default_rng(31)
    refs = [chroma_like(rng, n) for n in (150, 90, 260)]
    lives = [warped_copy(rng, r, 220) for r in refs]
    frames = np.stack([np.stack([lv[:, k] for lv in lives]) for k in range(220)])          # (T, n, 12)
    B = mods["batch"]
    for make in (lambda: B.OtwBatch(refs, 20, 3, kind="livenote_v2"), lambda: B.WtwBatch(refs, 16, 8)):
        full, first = make(), make()
        step = (lambda b, x: b.step_device(x)[0]) if hasattr(full, "step_device") else (lambda b, x: b.push_device(x))
        d = torch.from_numpy(frames).cuda()
        st_full = step(full, d).cpu().numpy()
        st_a = step(first, d[:90].contiguous()).cpu().numpy()
        snap = first.export_state()
        first.close()
        second = make()
        second.import_state(snap)
        st_b = step(second, d[90:].contiguous()).cpu().numpy()
        assert np.array_equal(np.concatenate([st_a, st_b]), st_full)
        assert np.array_equal(second.positions(), full.positions())
        for x, y in zip(second.paths(), full.paths()):
            assert np.array_equal(x, y)
        full.close()
        second.close()
    other = B.OtwBatch(refs, 21, 3, kind="livenote_v2")
    with pytest.raises(mods["_native"].AfsError):
        other.import_state(snap)
    other.close()


@pytest.mark.parametrize("kind", ["otw", "livenote_v2"])
def test_full_config_4096_streams_c500_sampled_parity(mods, orc, kind):
    """BASELINE config[3] at full size — 4096 concurrent streams, c = 500, 3000-frame references, the bench's own data
    generator — stepped one launch per live frame for 700 frames; the complete paths of 10 sampled streams (first, last,
    group boundaries, random) are bit-equal to the oracle's."""
    import sys
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    argv, sys.argv = sys.argv, sys.argv[:1]
    try:
        import bench
    finally:
        sys.argv = argv
    S, N, T, c = 4096, 3000, 700, 500
    ref, frames = bench.synth_streams(torch, S, N, T, 777, "cuda")
    b = mods["batch"].OtwBatch(ref, c, 3, kind=kind)
    for k in range(T):
        b.step_device(frames[k], want_points=False)
    got = b.paths()
    pos = b.positions()
    params = {"c": c, "max_run_count": 3, "search_band_width": c}
    ocls = orc.OnlineTimeWarping if kind == "otw" else orc.LiveNoteV2
    rng = np.random.default_rng(3)
    sample = [0, 6, 7, 13, 4095, 4094] + [int(x) for x in rng.integers(0, S, size=4)]
    h_ref = ref.cpu().numpy()
    h_frames = frames.cpu().numpy()
    for s in sample:
        o = ocls(np.ascontiguousarray(h_ref[s]), dict(params))
        for k in range(T):
            if o.insert(h_frames[k, s]) == "stop":
                break
        assert np.array_equal(got[s], o.path_array()), (kind, s)
        assert (int(pos[s, 0]), int(pos[s, 1])) == (o.t, o.j), (kind, s)
    b.close()


@pytest.mark.parametrize("kind", ["otw", "livenote_v2", "livenote_v2_diff"])
def test_on_demand_state_views_match_the_dense_matrices(mods, orc, kind):
    """The reference's .acc_cost / .cost / .direction / .previous / .run_count (otw_eran.py:23-35, livenote_v2.py:22-38)
    are read back on demand from the device's moving window: every cell of row t over columns j-c..j and of column j over
    rows t-c..t must equal the oracle's dense acc_cost bit for bit, at every step of a run."""
    rng = np.random.default_rng(11)
    ref = chroma_like(rng, 140)
    live = warped_copy(rng, ref, 170)
    c, mr = 12, 3
    diff = kind == "livenote_v2_diff"
    if kind == "otw":
        obj = mods["otw_eran"].OnlineTimeWarping(ref, {"c": c, "max_run_count": mr})
        ora = orc.OnlineTimeWarping(ref, {"c": c, "max_run_count": mr})
    else:
        p = {"search_band_width": c, "max_run_count": mr}
        obj = mods["livenote_v2"].LiveNoteV2(ref, p, chroma_diff=diff)
        ora = orc.LiveNoteV2(ref, p, chroma_diff=diff)
    checked = 0
    for i in range(live.shape[1]):
        r1, r2 = obj.insert(live[:, i]), ora.insert(live[:, i])
        assert r1 == r2
        if r1 == "stop":
            break
        if i % 7 and i > 3:
            continue
        w = obj._batch.window(0)
        assert (w["t"], w["j"]) == (ora.t, ora.j)
        assert w["direction"] in ("Both", "Row", "Column") and w["previous"] in (None, "Row", "Column") and w["run_count"] >= 0
        for (x, y), v in obj.acc_cost_window().items():
            want = ora.acc(x, y)
            assert v == want or (np.isinf(v) and np.isinf(want)), (i, x, y, v, want)
            checked += 1
        cw = obj.cost_window()
        lv = live[:, ora.t]
        for y in w["cols"]:
            want = float(np.sqrt(np.sum((lv - ref[:, y]) ** 2))) if diff else float(1 - np.dot(lv, ref[:, y]))
            assert abs(cw[(w["t"], int(y))] - want) < 1e-12
    assert checked > 300
    assert obj.direction in ("Both", "Row", "Column")
