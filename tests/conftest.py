import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 GPU (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    """Fixtures generated from the real reference class by tests/golden/make_golden.py."""
    arrays = np.load(os.path.join(ROOT, "tests", "golden", "vq_golden.npz"))
    with open(os.path.join(ROOT, "tests", "golden", "vq_golden.json")) as f:
        meta = json.load(f)
    return arrays, meta


@pytest.fixture(scope="session")
def lib():
    import b200vq
    b200vq.build_extension()
    return b200vq.load_library()
