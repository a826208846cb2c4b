"""pytest configuration.  `-m "not gpu"` = oracle / host-logic / ABI-export tests (CPU only);
`-m gpu` = the parity tests proper (CUDA path vs oracle, through the C ABI)."""
import importlib
import os
import sys
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    warnings.filterwarnings("ignore", category=RuntimeWarning)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Builds libcddp_b200.so / liboracle.so if missing (nvcc cross-compiles without a GPU)."""
    lib = os.path.join(ROOT, "cddp-cpp_b200", "libcddp_b200.so")
    orc = os.path.join(ROOT, "oracle", "liboracle.so")
    if not (os.path.exists(lib) and os.path.exists(orc)):
        import __graft_entry__ as g
        g.build()


@pytest.fixture(scope="session")
def cddp():
    return importlib.import_module("cddp-cpp_b200")


@pytest.fixture(scope="session")
def problems():
    return importlib.import_module("cddp-cpp_b200.problems")


@pytest.fixture(scope="session")
def ob():
    import oracle_binding
    return oracle_binding


@pytest.fixture(scope="session")
def npo():
    import np_oracle
    return np_oracle


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def rel_err(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    if a.size == 0 and b.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300))


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
