"""Shared pytest configuration.

`-m "not gpu"` runs everywhere (oracle vs golden vectors, host logic, C-ABI symbol checks,
world_size-2 gloo tests); `-m gpu` needs a B200 and the built CUDA library.
"""

import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# Two tests drive two ranks from two host threads of THIS process; a lazily loaded kernel may synchronise
# the context while the peer rank spins on it, so load every kernel up front (must be set before CUDA starts).
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200) and the built library")


def golden(name):
    data = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    return {k: data[k] for k in data.files}


def golden_names(prefix):
    files = sorted(glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))
    return [os.path.basename(f)[: -len(".npz")] for f in files]


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    denom = max(np.linalg.norm(b.ravel()), 1e-300)
    return np.linalg.norm((a - b).ravel()) / denom


@pytest.fixture(scope="session")
def has_cuda():
    try:
        import experiments_lanczos_adjoints_b200 as bl

        return bl.device_count() > 0
    except Exception:
        return False
