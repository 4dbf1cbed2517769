import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def kat():
    """Known-answer values + geometries extracted from the reference's own tests
    (tests/golden/make_golden.py -> tests/golden/reference_kat.json)."""
    with open(os.path.join(ROOT, "tests", "golden", "reference_kat.json")) as f:
        d = json.load(f)
    for v in d.values():
        v["xyz"] = np.array(v["xyz"], dtype=np.float64).reshape(-1, 3)
        v["numbers"] = np.array(v["numbers"], dtype=np.int32)
    return d
