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
    from oracle import pyoracle

    pyoracle.build()
    assert pyoracle.lib().orc_selftest_unfused() == 1, "oracle was built with FMA contraction"
    return pyoracle


@pytest.fixture(scope="session")
def ctx():
    """One sfh context on cuda:0 for the whole GPU session (C ABI through ctypes)."""
    from stonefish_b200 import HydroContext

    c = HydroContext(0)
    yield c
    c.close()
