import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _gpu_box() -> bool:
    """A machine with an NVIDIA device node.  On such a box nothing is skipped: a missing library or a failing
    egs_create must FAIL the gpu tests loudly (libegs has no CPU fallback)."""
    return any(os.path.exists(p) for p in ("/dev/nvidiactl", "/dev/nvidia0", "/dev/dxg"))


def pytest_collection_modifyitems(config, items):
    if _gpu_box():
        return
    skip = pytest.mark.skip(reason="no NVIDIA device on this machine (gpu tests run on the B200 box; no CPU fallback exists)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def egs():
    """The product library through its C ABI.  No fallback: fails when CUDA is missing."""
    import egs_b200
    return egs_b200.capi
