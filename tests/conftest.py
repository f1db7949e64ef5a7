"""pytest configuration: the `gpu` marker, golden-fixture loaders, repo root on sys.path."""
import json
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


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def manifest():
    with open(os.path.join(GOLDEN, "MANIFEST.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden():
    class _G:
        def __init__(self):
            self._c = {}

        def __getitem__(self, name):
            if name not in self._c:
                with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
                    self._c[name] = {k: z[k] for k in z.files}
            return self._c[name]
    return _G()
