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
