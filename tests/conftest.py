import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    class G:
        def __getattr__(self, name):
            d = np.load(os.path.join(GOLD, name + ".npz"))
            setattr(self, name, d)
            return d
    return G()


@pytest.fixture(scope="session")
def c1(golden):
    """The shipped sample operator D (4^4 x 4 x 3 hopping matrix) and k of src/main.cpp:845-847."""
    m = golden.c1_matrix
    row = m["row"].astype(np.int64)
    col = m["col"].astype(np.int64)
    val = m["val"]
    k = 0.05 + 8 * ((0.17865 - 0.05) / 10.)
    return dict(row=row, col=col, val=val, k=k, n=3072, dims=[4, 4, 4, 4, 4, 3])


def relerr(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
