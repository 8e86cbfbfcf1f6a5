import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
TESTS = os.path.dirname(os.path.abspath(__file__))
if TESTS not in sys.path:
    sys.path.insert(0, TESTS)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ctx():
    """One library context on cuda:0 for the whole GPU test session."""
    import jues.jl_b200 as jb
    try:
        c = jb.Context(0)
    except jb.JuesError as e:            # library built, but no sm_100 device / driver on this machine
        pytest.skip(f"no usable sm_100 device: {e}")
    yield c
    c.close()
