import os
import sys

# The sharded tests run several ranks as threads of this process; a kernel that is lazily
# loaded while a peer thread spins in a barrier kernel would dead-lock the context.  The
# library force-loads its kernels too (plm_setup_t); this makes the whole process eager.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200 box)")


def _gpu_count():
    try:
        from pydca_b200 import _lib
        return _lib.lib().dca_device_count()
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    if _gpu_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name))
    return load


@pytest.fixture(scope="session")
def port():
    import oracle
    return oracle.port()


def relmax(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
