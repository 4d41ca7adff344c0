import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def metrpo_lib():
    """Built C-ABI library (builds it if stale; nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as ge
    ge.build()
    from me_trpo_b200 import lib
    return lib
