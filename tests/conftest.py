import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _cuda_available():
    try:
        from fuzzysearch_b200 import _native
        return _native.device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def cuda_device():
    if not _cuda_available():
        pytest.fail("gpu-marked test ran without a usable CUDA device / libfuzzb200.so")
    return 0
