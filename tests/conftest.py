import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (runs on the B200 box only)")


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "vq_golden.npz")
    with np.load(path) as f:
        return {k: f[k] for k in f.files}


def gsub(golden, prefix):
    """All arrays under 'prefix/' with the prefix stripped (one level only)."""
    n = len(prefix) + 1
    return {k[n:]: v for k, v in golden.items() if k.startswith(prefix + "/") and "/" not in k[n:]}
