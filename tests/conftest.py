import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # B2M_GATHER_MODE=tma runs the GPU tests with the TMA gather4 row gathers instead of the default cp.async ones
    # (the library itself reads no environment variables: this is the test host calling b2m_set_option)
    mode = {"tma": 1, "cpasync2": 2, "cpasync_all": 3}.get(os.environ.get("B2M_GATHER_MODE", ""))
    if mode:
        from box2mask_b200 import _lib
        _lib.set_option(_lib.OPT_GATHER_MODE, mode)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
