import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def golden_params(fx, dtype=None):
    out = {}
    for k, v in fx.items():
        if k.startswith("p:"):
            out[k[2:]] = v.astype(dtype) if dtype is not None else v
    return out


@pytest.fixture(scope="session")
def golden():
    return load_golden
