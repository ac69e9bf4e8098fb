import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with -m gpu on the GPU box")


def _has_gpu():
    try:
        import ctypes
        from swiftest_b200 import _lib
        L = _lib.load()
        h = ctypes.c_void_p()
        rc = L.swcu_create(0, ctypes.byref(h))
        if rc == 0:
            L.swcu_destroy(h)
            return True
    except Exception:
        pass
    return False


@pytest.fixture(scope="session")
def oracle():
    from oracle import load
    return load()


@pytest.fixture(scope="session")
def ctx():
    """GPU context.  Fails (does not skip) when the CUDA library or the GPU is missing: the product path has no
    fallback and a gpu-marked test that silently passed on CPU would be meaningless."""
    from swiftest_b200 import Context
    c = Context(0)
    yield c
    c.close()


def rel_err_scaled(a, ref, scale):
    """max |a - ref| / scale with scale > 0 elementwise (per-component sum of |terms|)."""
    scale = np.where(scale > 0, scale, 1.0)
    return float(np.max(np.abs(a - ref) / scale))
