import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def entry():
    import __graft_entry__ as g
    return g


@pytest.fixture(scope="session")
def orc():
    from oracle import afs_oracle
    afs_oracle.lib()
    return afs_oracle


def chroma_like(rng, n, smooth=0.7):
    """Synthetic unit-norm chroma (12, n) with temporal smoothness."""
    import numpy as np
    x = rng.random((12, n))
    for k in range(1, n):
        x[:, k] = smooth * x[:, k - 1] + (1 - smooth) * x[:, k]
    return x / np.linalg.norm(x, axis=0)


def warped_copy(rng, ref, m, noise=0.05):
    import numpy as np
    n = ref.shape[1]
    u = np.linspace(0, 1, m)
    pos = np.clip((u + 0.08 * np.sin(6 * np.pi * u)) * (n - 1), 0, n - 1)
    idx = np.round(pos).astype(int)
    y = ref[:, idx] + noise * rng.random((12, m))
    return y / np.linalg.norm(y, axis=0)
