import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_present() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) where there is no CUDA device; the product itself has no CPU path."""
    if _cuda_present():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built_checkers():
    """Build the test-only checkers: the oracle's C restatement, the reference build (only where /root/reference
    exists) and the host emulation of the device logic."""
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "all"], check=True, capture_output=True)
    from tests import util
    util.build_hostemu()
    yield
