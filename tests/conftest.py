import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["sq48p8", "r40c60p10", "r24c36p0"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden(dict):
    """npz fixture with tensors on demand."""

    def t(self, key):
        v = self[key]
        return torch.from_numpy(np.array(v))


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, f"{name}.npz")) as z:
        return Golden({k: z[k] for k in z.files})


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    return load_golden(request.param)
