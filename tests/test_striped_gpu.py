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
