import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib_built():
    """Build (if stale) the product library and the oracle; both are cross-compiled without a GPU."""
    from sarpro_b200 import build as B
    B.build()
    from oracle import pyoracle
    pyoracle.build()
    return True


@pytest.fixture(scope="session")
def ctx(lib_built):
    import sarpro_b200 as S
    c = S.Context(0)  # raises SarproError(NO_DEVICE) without a GPU: there is no CPU fallback
    yield c
    c.close()
