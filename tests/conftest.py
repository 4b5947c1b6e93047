import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import o3bz
    o3bz.lib()
    return o3bz


@pytest.fixture(scope="session")
def engine():
    """The product library through its C ABI.  No fallback: a missing .so or device is an error."""
    import threebz_b200 as t
    t._ffi.lib()
    return t


@pytest.fixture(scope="session")
def ctx(engine):
    return engine.default_ctx(0)
