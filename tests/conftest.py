import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as graft  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ho():
    """The CPU oracle (test infrastructure)."""
    return graft.load_oracle()


@pytest.fixture(scope="session")
def pkg():
    """The product package; the C-ABI library must be built (no fallback)."""
    return graft.load_package()


@pytest.fixture(scope="session")
def gpu_pkg(pkg):
    import ctypes as C

    cnt = C.c_int()
    pkg._lib.load().hh_device_count(C.byref(cnt))
    if cnt.value < 1:
        pytest.fail("GPU test selected but no CUDA device is visible (no CPU fallback exists)")
    return pkg


def rel_err(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300))


GOLDEN = os.path.join(ROOT, "tests", "golden")
