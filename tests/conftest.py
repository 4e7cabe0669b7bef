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


@pytest.fixture(params=["direct", "lists"])
def surface_mode(request):
    """The surface path has two ways through k_surface_lists: DIRECT mode (no more surface items than streams: every warp classifies
    its own sub-item — what every small test scene would take) and the LIST mode of full-size steps (k_classify, wet / clip roles,
    segments). SFH_NO_DIRECT=1 sends small scenes through the lists too; tests that use this fixture run both ways."""
    old = os.environ.get("SFH_NO_DIRECT")
    os.environ["SFH_NO_DIRECT"] = "1" if request.param == "lists" else "0"
    yield request.param
    if old is None:
        os.environ.pop("SFH_NO_DIRECT", None)
    else:
        os.environ["SFH_NO_DIRECT"] = old
