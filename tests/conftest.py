import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The C-ABI library is built in-tree by __graft_entry__.build(); build it here if a fresh
    checkout has not done so yet (nvcc cross-compiles without a GPU)."""
    lib = os.path.join(ROOT, "audio-deepfake-detection-fmsl_b200", "lib", "libb200fe.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "audio-deepfake-detection-fmsl_b200", "csrc")])
    yield


@pytest.fixture(scope="session")
def fe():
    import b200_frontend

    return b200_frontend
