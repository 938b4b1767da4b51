"""GPU edge cases: extreme parameters, tiny inputs, non-default hops (aligned and unaligned frame
starts -> TMA-staged and fallback load paths), error paths."""
import numpy as np
import pytest

from conftest import chroma_like, warped_copy

pytestmark = pytest.mark.gpu


def run_insert(obj, live):
    for i in range(live.shape[1]):
        if obj.insert(live[:, i]) == "stop":
            break
    return np.asarray(obj.path, dtype=np.int64).reshape(-1, 2)


@pytest.mark.parametrize("c,mr", [(1, 1), (1, 3), (2, 1), (3, 2), (200, 3), (64, 7)])
def test_otw_extreme_parameters_vs_oracle(entry, orc, c, mr):
    batch = entry.submodule("batch")
    rng = np.random.default_rng(c * 10 + mr)
    ref = chroma_like(rng, 90)
    live = warped_copy(rng, ref, 140)
    import torch
    for kind, okind in (("otw", orc.OnlineTimeWarping), ("livenote_v2", orc.LiveNoteV2), ("livenote", orc.LiveNote)):
        key = "c" if kind == "otw" else "search_band_width"
        o = okind(ref, {key: c, "max_run_count": mr})
        want = run_insert(o, live)
        b = batch.OtwBatch([ref], c, mr, kind=kind)
        frames = torch.from_numpy(np.ascontiguousarray(live.T).reshape(-1, 1, 12)).cuda()
        st, npts, _ = b.step_device(frames)
        assert np.array_equal(b.paths()[0], want), (kind, c, mr)
        assert int(npts.max()) <= b.pts
        b.close()


def test_otw_tiny_references(entry, orc):
    batch = entry.submodule("batch")
    rng = np.random.default_rng(5)
    import torch
    for n_ref in (1, 2, 3):
        ref = chroma_like(rng, n_ref)
        live = chroma_like(rng, 12)
        o = orc.OnlineTimeWarping(ref, {"c": 4, "max_run_count": 3})
        want = run_insert(o, live)
        b = batch.OtwBatch([ref], 4, 3, kind="otw")
        st, _, _ = b.step_device(torch.from_numpy(np.ascontiguousarray(live.T).reshape(-1, 1, 12)).cuda())
        assert np.array_equal(b.paths()[0], want), n_ref
        assert (st.cpu().numpy()[:, 0] == 1).any()          # the reference runs out ("stop")
        b.close()


def test_live_buffer_exhaustion(entry, orc):
    """otw_eran.py:53-55: the pre-allocated live buffer holds 2N frames; afterwards insert returns None forever."""
    otw = entry.submodule("otw_eran")
    rng = np.random.default_rng(6)
    ref = chroma_like(rng, 40)
    # a live stream that never advances along the reference: all-zero frames keep every cost at 1
    o = otw.OnlineTimeWarping(ref, {"c": 3, "max_run_count": 1000})
    oo = orc.OnlineTimeWarping(ref, {"c": 3, "max_run_count": 1000})
    rets, orets = [], []
    live = warped_copy(rng, ref, 30)
    slow = np.repeat(live[:, :20], 6, axis=1)               # 120 frames > 2N = 80
    for k in range(slow.shape[1]):
        rets.append(o.insert(slow[:, k]))
        orets.append(oo.insert(slow[:, k]))
        if rets[-1] == "stop":
            break
    assert rets == orets
    assert o.path == oo.path and (o.t, o.j) == (oo.t, oo.j)


@pytest.mark.parametrize("hop", [2048, 1024, 1000, 1002, 4096])
def test_chroma_other_hops(entry, orc, hop):
    chroma = entry.submodule("chroma")
    import torch
    rng = np.random.default_rng(hop)
    x = (0.2 * rng.standard_normal(40001)).astype(np.float32)
    plan = chroma.ChromaPlan(4096, hop)
    want = orc.create_chroma(orc.create_stft(x, 4096, hop))
    for compute, tol in (("tc", 1e-4), ("fp32", 1e-4), ("fp64", 1e-9)):
        xa = np.concatenate((x, np.zeros(3, np.float32)))
        d_out, foffs = plan.run(torch.from_numpy(xa).cuda(), [0, len(x)], out_dtype=torch.float64, compute=compute)
        got = d_out.cpu().numpy().reshape(12, -1)
        assert got.shape == want.shape, (hop, got.shape, want.shape)
        assert np.abs(got - want).max() < tol, (hop, compute)
    plan.close()


def test_chroma_unaligned_track_offsets(entry, orc):
    """Track offsets that are even but not multiples of 4 samples: frames are not 16-byte aligned, the
    kernel must take the guarded-load path and give the same answer."""
    chroma = entry.submodule("chroma")
    import torch
    rng = np.random.default_rng(3)
    a = (0.2 * rng.standard_normal(30002)).astype(np.float32)
    b = (0.2 * rng.standard_normal(25000)).astype(np.float32)
    flat = np.concatenate((a, b))
    plan = chroma.default_plan()
    d_out, foffs = plan.run(torch.from_numpy(flat).cuda(), [0, 30002, 55002], out_dtype=torch.float64)
    got = d_out.cpu().numpy()
    for k, x in enumerate((a, b)):
        want = orc.wav_samples_to_chroma(x)
        blk = got[12 * foffs[k] : 12 * foffs[k + 1]].reshape(12, -1)
        assert blk.shape == want.shape and np.abs(blk - want).max() < 1e-4


def test_error_paths(entry):
    dtw = entry.submodule("dtw")
    nat = entry.submodule("_native")
    batch = entry.submodule("batch")
    chroma = entry.submodule("chroma")
    with pytest.raises(nat.AfsError):
        dtw.DTW(np.zeros((13, 4)), np.zeros((13, 5)))        # only 12 features on the CUDA path
    with pytest.raises(AssertionError):
        dtw.DTW(np.zeros((12, 4)), np.zeros((11, 5)))
    with pytest.raises(nat.AfsError):
        batch.OtwBatch([np.zeros((12, 10))], 0, 3)             # c >= 1
    with pytest.raises(nat.AfsError):
        batch.WtwBatch([np.zeros((12, 10))], 1, 1)             # W >= 2
    with pytest.raises(nat.AfsError):
        batch.OtwBatch([np.zeros((12, 10))], 20000, 3)         # shared-memory limit, reported not crashed
    with pytest.raises(nat.AfsError):
        chroma.ChromaPlan(2048, 1024)                          # only n_fft = 4096 is implemented
    import ctypes as C
    L = nat.lib()
    assert L.afs_dtw_accumulate(None, None, None, None, None, None, None, None) == -1
    assert b"afs_dtw_accumulate" in L.afs_last_error()
