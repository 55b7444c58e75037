import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def vio():
    """The package (its directory name has hyphens, so it is imported by string)."""
    return importlib.import_module("visual-inertial-odometry_b200")


@pytest.fixture(scope="session")
def refshim():
    from tests import refshim as r
    return r
