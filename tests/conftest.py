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
    import oracle

    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def wp():
    """The package under test, on a real GPU; fails (does not skip) if the native path is unavailable."""
    import warp_b200

    assert warp_b200.is_cuda_available(), "GPU tests need a CUDA device and the built native library"
    return warp_b200
