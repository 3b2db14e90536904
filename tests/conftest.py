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


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not failed) on a host without a CUDA device. With a device they always run: a missing
    libs2d_b200.so must then fail loudly (there is no CPU fallback to hide behind)."""
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    with open(os.path.join(GOLDEN_DIR, f"{name}.json")) as f:
        g = json.load(f)
    z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    return g, z["labels"], z["tracks"], z["vis"]


@pytest.fixture(params=GOLDEN_CASES)
def golden_case(request):
    return (request.param,) + load_golden(request.param)
