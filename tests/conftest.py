import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


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
def golden_prop():
    return np.load(os.path.join(GOLDEN, "reference_propagation.npz"))


@pytest.fixture(scope="session")
def golden_masks():
    return np.load(os.path.join(GOLDEN, "reference_masks.npz"))


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build products are made by __graft_entry__.build(); make sure they exist for the tests."""
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.build()
