import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")
REFERENCE = "/root/reference/gnn-recommendations"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")
    config.addinivalue_line("markers", "reference: imports the read-only reference (build container only)")


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isdir(REFERENCE)
    for it in items:
        if "reference" in it.keywords and not have_ref:
            it.add_marker(pytest.mark.skip(reason="/root/reference not present on this box"))


@pytest.fixture(scope="session")
def tiny():
    return dict(np.load(os.path.join(GOLDEN, "tiny.npz")))


@pytest.fixture(scope="session")
def c1gold():
    return dict(np.load(os.path.join(GOLDEN, "c1.npz")))


@pytest.fixture(scope="session")
def c1split():
    from gnn_recommendations_b200.synthetic import synth_split

    return synth_split("C1", 42)
