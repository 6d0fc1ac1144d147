import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, GOLDEN_DIR):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _unjson_float(x):
    return float(x) if isinstance(x, str) else x


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def golden_arrays():
    return np.load(os.path.join(GOLDEN_DIR, "golden_arrays.npz"))


@pytest.fixture(scope="session")
def gf():
    """json float decoder: fixtures store -inf/nan as strings."""
    return _unjson_float
