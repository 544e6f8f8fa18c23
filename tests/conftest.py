import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ref():
    """The compiled, unmodified reference (oracle/_ref).  Built here when /root/reference exists."""
    from oracle import ref as R
    if not R.available():
        try:
            R.build()
        except Exception:
            pass
    if not R.available():
        pytest.skip("oracle/_ref/libfdm_ref.so not built (needs /root/reference)")
    return R


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
    return np.load(path)
