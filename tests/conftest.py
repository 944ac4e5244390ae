import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def hostsim():
    from tests import util
    return util.hostsim_engine()


@pytest.fixture(scope="session")
def gpu():
    from tests import util
    return util.gpu_engine()


@pytest.fixture(params=["hostsim", pytest.param("gpu", marks=pytest.mark.gpu)])
def engine(request):
    """Drop-in surface tests run twice: on the host-simulation double here (-m "not gpu") and on the CUDA library on
    the B200 box (-m gpu)."""
    return request.getfixturevalue(request.param)
