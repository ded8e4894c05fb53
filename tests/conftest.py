import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")

MT = 173.0
WT = 1.4915000200271606
G = 1.2177157847767195


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"))

    return load


@pytest.fixture(scope="session", autouse=True)
def _build_libraries():
    """The C-ABI libraries are built in-tree once per session (nvcc cross-compiles without a GPU)."""
    from madflow_b200 import build

    build.build_all()


def sm_params(alpha_s=None, g=G):
    """Parameter dict for the oracle: frozen couplings of tests/mockup_debug_me.py:22-26 by default."""
    if alpha_s is not None:
        g = 2.0 * np.sqrt(np.pi * np.asarray(alpha_s))
    return {"mdl_MT": MT, "mdl_WT": WT, "GC_10": -g + 0j, "GC_11": 1j * g, "GC_12": 1j * g * g}
