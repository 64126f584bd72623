import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    """The CPU checker (oracle/): built on demand; test infrastructure only."""
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def rb():
    """The product package; the native library must already be built (it travels with the snapshot)."""
    import recometrics_b200
    from recometrics_b200 import _capi
    _capi.load()
    return recometrics_b200
