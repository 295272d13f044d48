import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def exact_mode(monkeypatch):
    """Plans created inside the test use the exact-order kernel (PIRIP_FSK_MODE_EXACT): the tests that compare every float
    with the oracle bit for bit.  The default mode of the batch entry points is the tolerance-mode kernel, which has its own
    tests (tests/test_fsk_fast_gpu.py)."""
    monkeypatch.setenv("PIRIP_FSK_MODE", "exact")
