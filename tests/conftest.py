"""pytest configuration: the `gpu` marker and shared fixtures.

CPU tier (`-m "not gpu"`): oracle vs pins/goldens, host logic (the product's device functions compiled
for the host by tests/emul), C-ABI export check, gloo sharding.  GPU tier (`-m gpu`): parity of the CUDA
path, called through the C ABI, against the oracle and the committed goldens.
"""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(Path(__file__).resolve().parent))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _build_native():
    """Build the oracle and the host-emulation object (seconds); the CUDA library is built by __graft_entry__.build()."""
    subprocess.check_call(["make", "-s", "-C", str(ROOT / "oracle")])
    subprocess.check_call(["make", "-s", "-C", str(ROOT / "tests" / "emul")])


@pytest.fixture(scope="session")
def O():
    from oracle import oracle
    return oracle


@pytest.fixture(scope="session")
def nn(O):
    return O.OracleNN()


@pytest.fixture(scope="session")
def ee_home(O):
    return O.fk(O.Q_HOME)[0]


@pytest.fixture(scope="session")
def track_wp(O, ee_home):
    X, Y, Z, R = O.load_track()
    X, Y, Z = O.shift_track(X, Y, Z, ee_home)
    return X, Y, Z, R


@pytest.fixture()
def rng():
    return np.random.default_rng(12345)
