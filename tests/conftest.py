"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"` runs here (no GPU): oracle vs golden vectors, host logic, C-ABI symbol checks.
`-m gpu` runs on a B200: the parity tests proper, all through libdct_cuda's C-ABI.
"""
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
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")
    config.addinivalue_line("markers", "multigpu: needs two or more CUDA devices; skipped (not passed) on a smaller box")


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    return binding.load("oracle")


@pytest.fixture(scope="session")
def golden_blocks():
    with open(os.path.join(GOLDEN, "golden_blocks.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_planes():
    return np.load(os.path.join(GOLDEN, "golden_planes.npz"))


def unhex(lst, shape=None):
    a = np.array([float.fromhex(s) for s in lst], dtype=np.float64)
    return a.reshape(shape) if shape else a
