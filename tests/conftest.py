import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    """Golden vectors produced by the unmodified reference (oracle/_ref/ref_golden, see oracle/make_golden.py)."""
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "xerus_ref_v1.npz")))


def golden_tt(g, name):
    """Cores + core position of a TT dumped by drv::Writer::tt."""
    d = int(g[name + ".d"])
    core = int(g[name + ".core"])
    return [g["%s.c%d" % (name, i)] for i in range(d)], (core if core >= 0 else None)
