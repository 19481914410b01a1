import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def lib():
    """The C-ABI library; built on demand (nvcc cross-compiles sm_100a without a GPU) if it is not there yet."""
    from gvi_gaussian_process_b200 import _lib, build

    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    return _lib.load()
