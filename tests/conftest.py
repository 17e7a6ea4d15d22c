import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run on the GPU box with -m gpu)")


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
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def scan_pair():
    """BASELINE config-1 stand-in: two synthetic 64-beam scans (~125k points each), guess, ground truth."""
    from lv_slam_b200 import synth
    return synth.config1_pair()


@pytest.fixture(scope="session")
def small_pair():
    """16-beam x 600-azimuth pair (~9k points): sizes the oracle finishes in milliseconds."""
    from lv_slam_b200 import synth
    return synth.config1_pair(n_beams=16, n_az=600, seed=5)
