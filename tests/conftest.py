import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def engine():
    """One lisreg context on cuda:0.  No CPU fallback: raises if the library or device is missing."""
    from lis_slam_b200 import engine as E
    eng = E.Engine(device=0)
    yield eng
    eng.close()
