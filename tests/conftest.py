import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    """Vectors computed by the UNMODIFIED reference (tests/golden/make_golden_from_reference.py)."""
    path = os.path.join(ROOT, "tests", "golden", "reference_golden.npz")
    return dict(np.load(path))


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the in-tree libraries exist (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g

    g.build()
