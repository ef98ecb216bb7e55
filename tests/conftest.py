import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """Build (or reuse) libdla_b200.so and the host shim."""
    import __graft_entry__ as g

    g.build()
    return g


@pytest.fixture(scope="session")
def gpu(built):
    from gpy_dla_detection_b200 import _lib

    _lib.init(0)
    return _lib
