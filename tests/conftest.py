"""pytest configuration: marker registration, import paths, shared fixtures.

`-m "not gpu"` runs here (no GPU): oracle vs golden vectors, host-side logic, C-ABI symbol checks,
world_size-2 gloo tests.  `-m gpu` runs on a B200: parity of the CUDA path against the oracle,
called through the C ABI.
"""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "nbody-gnn-hpc_b200"
for p in (str(ROOT), str(PKG)):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = ROOT / "tests" / "golden"

# sharded mode: how long a kernel waits for another rank before it reports it lost (read once by the library);
# short, so that the lost-peer test takes seconds
import os  # noqa: E402
os.environ.setdefault("NB_PEER_TIMEOUT_MS", "1500")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_ok() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _cuda_ok():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(GOLDEN / name)
    return load


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def engine():
    from hpc import _cuda
    return _cuda.get_engine()


def rel_rows(a, b):
    """Per-particle vector-norm relative difference (the norm the 1e-10 bar is stated in)."""
    a, b = np.asarray(a), np.asarray(b)
    den = np.linalg.norm(b, axis=-1)
    den = np.where(den > 0, den, 1.0)
    return np.linalg.norm(a - b, axis=-1) / den
