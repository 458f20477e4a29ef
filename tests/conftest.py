import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure; never imported by quiver_b200)."""
    import oracle as _o
    _o.build()
    return _o


@pytest.fixture(scope="session")
def capi():
    from quiver_b200 import capi as _c
    _c.load()
    return _c
