import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pdb_fixtures():
    return np.load(os.path.join(GOLDEN, "pdb_fixtures.npz"))


@pytest.fixture(scope="session")
def synthetic_fixtures():
    return np.load(os.path.join(GOLDEN, "synthetic.npz"))


@pytest.fixture(scope="session")
def totals():
    import json

    with open(os.path.join(GOLDEN, "totals.json")) as f:
        return json.load(f)
