"""GPU parity: column-striped DTW (K4) with all stripes on one GPU vs the oracle (bit-exact)."""
import numpy as np
import pytest

from conftest import chroma_like, warped_copy

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def striped(entry):
    return entry.submodule("striped")


@pytest.mark.parametrize("M,N,G", [(300, 260, 2), (1000, 1500, 3), (129, 64, 4), (700, 33, 8), (5, 9, 3), (2500, 2000, 5)])
def test_striped_matches_oracle(striped, orc, M, N, G):
    rng = np.random.default_rng(M * 7 + N + G)
    ref = chroma_like(rng, N)
    live = warped_copy(rng, ref, M)
    acc_end, path = striped.dtw_striped_local(live, ref, G)
    _, oend, opath = orc.DTW(live, ref, dense=False)
    assert acc_end == oend
    assert np.array_equal(path, opath)


def test_striped_random_and_ties(striped, orc):
    rng = np.random.default_rng(1)
    a, b = rng.random((12, 400)), rng.random((12, 519))
    acc_end, path = striped.dtw_striped_local(a, b, 4)
    _, oend, opath = orc.DTW(a, b, dense=False)
    assert acc_end == oend and np.array_equal(path, opath)
    ea = rng.integers(0, 8, size=(12, 300)) / 1024.0      # exact arithmetic, heavy ties
    eb = rng.integers(0, 8, size=(12, 277)) / 1024.0
    acc_end, path = striped.dtw_striped_local(ea, eb, 3)
    _, oend, opath = orc.DTW(ea, eb, dense=False)
    assert acc_end == oend and np.array_equal(path, opath)


def test_striped_two_gpus_peer_memory_handoff():
    """One stripe per GPU, boundary columns stored straight into the neighbour's memory (CUDA IPC over
    NVLink) while both kernels run; path and acc_end must equal the oracle's.  Needs >= 2 GPUs."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(root, "tools", "striped_dist.py"), "3000", "4100", "check"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "PARITY path True acc_end True" in out.stdout, out.stdout[-2000:]


def test_stripe_wait_is_bounded(striped, monkeypatch):
    """A stripe whose left neighbour never raises its flags (crashed rank) must not hang the GPU: the kernel gives up
    after AFS_STRIPE_WAIT_MS, finishes, and poisons acc_end with NaN (StripedDtwDistributed.check() raises on it)."""
    import time
    import torch
    monkeypatch.setenv("AFS_STRIPE_WAIT_MS", "200")
    rng = np.random.default_rng(5)
    M, W = 700, 96
    a, b = chroma_like(rng, M), chroma_like(rng, W)
    st = striped._Stripe(M, W, 50, False)
    leftb = torch.zeros(M, dtype=torch.float64, device="cuda")
    in_flag = torch.zeros((M + 127) // 128, dtype=torch.int32, device="cuda")      # never raised
    t0 = time.perf_counter()
    st.accumulate(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), leftb, in_flag, None, None)
    torch.cuda.synchronize()
    waited = time.perf_counter() - t0
    assert 0.15 < waited < 5.0
    assert bool(torch.isnan(st.plan.acc_end[0]))
    # the same stripe with the flags raised completes normally
    in_flag.fill_(1)
    st.accumulate(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), leftb, in_flag, None, None)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(st.plan.acc_end[0]))
    st.close()
