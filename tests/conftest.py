import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def lib_built():
    """libbvb.so is built in-tree on demand (nvcc cross-compiles without a GPU)."""
    from bayesianvars_b200.build import build_lib
    return build_lib()


@pytest.fixture(scope="session")
def handle(lib_built):
    """Opens cuda:0 through the C ABI; fails loudly when there is no device."""
    from bayesianvars_b200 import Handle
    h = Handle(0, seed=20240229)
    yield h
    h.close()
