"""Shared pytest configuration.

``-m "not gpu"``: oracle vs golden vectors, host logic, C-ABI symbol checks, gloo tests (CPU only).
``-m gpu``     : parity tests proper -- they call the CUDA kernels through the C-ABI.
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_golden.json")) as f:
        j = json.load(f)
    a = dict(np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz")))
    return j, a


def have_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
