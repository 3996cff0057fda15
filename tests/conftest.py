import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    """tests/golden/<name>.npz -> nested dict of torch tensors ('a/b' keys become d['a']['b'])."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = {}
    for key in z.files:
        arr = z[key]
        if arr.dtype.kind in "US":                       # string arrays (parameter names): kept out of the tensor dict
            continue
        t = torch.from_numpy(arr) if arr.dtype != np.bool_ else torch.from_numpy(arr.astype(np.uint8)).bool()
        if "/" in key:
            head, tail = key.split("/", 1)
            out.setdefault(head, {})[tail] = t
        else:
            out[key] = t
    return out


@pytest.fixture
def golden():
    return load_golden
