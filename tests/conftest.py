import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


DATA_DIR = os.path.join(ROOT, "data", "_ref")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def data_dir():
    if not os.path.isdir(DATA_DIR):
        pytest.skip("data/_ref missing (run `make -C oracle` where /root/reference exists)")
    return DATA_DIR
