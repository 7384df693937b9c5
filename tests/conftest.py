"""Shared test plumbing.  `-m "not gpu"` = oracle vs golden vectors, host logic, C-ABI symbols (no GPU needed);
`-m gpu` = parity tests proper: the CUDA path, called through the Cython module / C ABI, against the oracle."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle"), GOLDEN):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the product library and the C oracle once per session (no-ops when up to date)."""
    from assimulo_b200 import build as b
    b.build_lib()
    b.build_cython()
    import ens_oracle
    ens_oracle.build()


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def rel_err(a, b):
    """max |a-b| / max(|b|, 1): relative where the state is large, absolute near zero crossings."""
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0)))


def ref_available():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "assimulo", "solvers", "runge_kutta.py"))
