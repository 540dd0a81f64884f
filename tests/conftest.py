import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long CPU test")


def _gpu_ready():
    """`-m gpu` tests are skipped (not errored) on a box without a CUDA device, so that a failure of a gpu test
    always means the device path is wrong.  With a device present nothing is skipped: a missing
    libepoch_b200.so must fail loudly there, not hide behind a skip."""
    try:
        import torch
        if not torch.cuda.is_available():
            return False, "no CUDA device"
    except Exception as e:  # pragma: no cover
        return False, f"torch: {e}"
    return True, ""


def pytest_collection_modifyitems(config, items):
    ok, why = _gpu_ready()
    if ok:
        return
    skip = pytest.mark.skip(reason=f"gpu test: {why}")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
