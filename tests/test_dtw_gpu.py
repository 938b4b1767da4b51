"""GPU parity: dtw.DTW through the C ABI vs the CPU oracle (bit-exact in fp64)."""
import numpy as np
import pytest

from conftest import chroma_like, warped_copy

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dtw(entry):
    return entry.submodule("dtw")


@pytest.mark.parametrize("M,N", [(1, 1), (1, 7), (9, 1), (4, 5), (33, 31), (128, 128), (129, 257), (300, 260), (1000, 777)])
def test_dtw_random_bit_exact(dtw, orc, M, N):
    rng = np.random.default_rng(M * 1000 + N)
    a = rng.random((12, M))
    b = rng.random((12, N))
    cost, acc, path = dtw.DTW(a, b)
    ocost, oacc, opath = orc.DTW(a, b)
    assert np.array_equal(cost, ocost)
    assert np.array_equal(acc, oacc)
    assert path.dtype == np.int64 and np.array_equal(path, opath)


def test_dtw_silent_tie_breaking(dtw, orc):
    # all-zero chroma: every cost is exactly 1 -> massive ties
    cost, acc, path = dtw.DTW(np.zeros((12, 4)), np.zeros((12, 5)))
    assert path.tolist() == [[0, 0], [1, 0], [2, 0], [3, 0], [3, 1], [3, 2], [3, 3], [3, 4]]
    assert acc[-1, -1] == 8.0


def test_dtw_exact_arithmetic_ties(dtw, orc):
    # entries are multiples of 2^-10: all sums exact, order-independent, many ties
    rng = np.random.default_rng(5)
    a = rng.integers(0, 8, size=(12, 500)) / 1024.0
    b = rng.integers(0, 8, size=(12, 431)) / 1024.0
    cost, acc, path = dtw.DTW(a, b)
    ocost, oacc, opath = orc.DTW(a, b)
    assert np.array_equal(acc, oacc) and np.array_equal(path, opath)


def test_dtw_large_no_dense(dtw, orc):
    rng = np.random.default_rng(11)
    ref = chroma_like(rng, 3033)
    live = warped_copy(rng, ref, 3647)
    cost, acc, path = dtw.DTW(live, ref, dense_limit=0)
    assert cost is None and acc is None
    _, oend, opath = orc.DTW(live, ref, dense=False)
    assert np.array_equal(path, opath)
    assert dtw.DTW.last_acc_end == oend


def test_dtw_batch_ragged(dtw, orc):
    rng = np.random.default_rng(3)
    lens = [(200, 300), (513, 129), (1, 40), (700, 650), (128, 1)]
    A = [rng.random((12, m)) for m, n in lens]
    B = [rng.random((12, n)) for m, n in lens]
    paths, ends = dtw.dtw_batch(A, B)
    for k in range(len(lens)):
        _, oend, opath = orc.DTW(A[k], B[k], dense=False)
        assert np.array_equal(paths[k], opath), k
        assert ends[k] == oend


# ---- fp32 mode: north_star tolerance = accumulated costs within 1e-5 relative of the fp64 reference ----
FP32_REL_TOL = 1e-5


@pytest.mark.parametrize("M,N", [(300, 260), (3000, 2500), (8000, 8000)])
def test_dtw_fp32_mode_accumulated_cost(dtw, orc, M, N):
    rng = np.random.default_rng(M + N)
    ref = chroma_like(rng, N)
    live = warped_copy(rng, ref, M)
    _, oend, opath = orc.DTW(live, ref, dense=False)
    cost, acc, path = dtw.DTW(live, ref, dtype="fp32", dense_limit=0)
    got = dtw.DTW.last_acc_end
    assert abs(got - oend) <= FP32_REL_TOL * abs(oend), (got, oend)
    # the fp32 path is a valid monotone warping path with (almost) the optimal cost
    assert path[0].tolist() == [0, 0] and path[-1].tolist() == [M - 1, N - 1]
    d = np.diff(path, axis=0)
    assert ((d >= 0) & (d <= 1)).all() and (d.sum(axis=1) >= 1).all()


def test_dtw_fp32_dense_matrices(dtw, orc):
    rng = np.random.default_rng(77)
    ref = chroma_like(rng, 700)
    live = warped_copy(rng, ref, 650)
    ocost, oacc, _ = orc.DTW(live, ref)
    cost, acc, _ = dtw.DTW(live, ref, dtype="fp32")
    assert np.abs(cost - ocost).max() < 1e-6
    assert np.abs(acc - oacc).max() <= FP32_REL_TOL * np.abs(oacc).max()


def test_dtw_size_independent_properties_at_full_size(dtw):
    """BASELINE config[2] size (20k x 20k) without an oracle run: (i) a sequence aligned with
    itself gives the diagonal with accumulated cost = 2 * sum of (tiny) self-costs, (ii) swapping
    the arguments transposes the optimal cost, (iii) fp32 mode agrees with fp64 mode to 1e-5."""
    rng = np.random.default_rng(123)
    n = 20000
    ref = chroma_like(rng, n)
    live = warped_copy(rng, ref, n)
    _, _, p = dtw.DTW(ref, ref, dense_limit=0)
    assert np.array_equal(p, np.stack([np.arange(n), np.arange(n)], axis=1))
    assert abs(dtw.DTW.last_acc_end) < 1e-9
    _, _, p1 = dtw.DTW(live, ref, dense_limit=0)
    e1 = dtw.DTW.last_acc_end
    _, _, p2 = dtw.DTW(ref, live, dense_limit=0)
    e2 = dtw.DTW.last_acc_end
    assert abs(e1 - e2) <= 1e-9 * abs(e1)
    assert p1[0].tolist() == [0, 0] and p1[-1].tolist() == [n - 1, n - 1]
    d = np.diff(p1, axis=0)
    assert ((d >= 0) & (d <= 1)).all() and (d.sum(axis=1) >= 1).all()
    _, _, p3 = dtw.DTW(live, ref, dtype="fp32", dense_limit=0)
    e3 = dtw.DTW.last_acc_end
    assert abs(e3 - e1) <= FP32_REL_TOL * abs(e1), (e3, e1)
