import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = sorted(f[:-5] for f in os.listdir(GOLDEN_DIR) if f.endswith(".json"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    with open(os.path.join(GOLDEN_DIR, f"{name}.json")) as f:
        g = json.load(f)
    z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    return g, z["labels"], z["tracks"], z["vis"]


@pytest.fixture(params=GOLDEN_CASES)
def golden_case(request):
    return (request.param,) + load_golden(request.param)
