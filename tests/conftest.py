import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def vo():
    """The product package (hyphenated directory name -> importlib)."""
    return importlib.import_module("visual-odometry_b200")


@pytest.fixture(scope="session")
def synth():
    return importlib.import_module("visual-odometry_b200.synth")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib

    return oracle_lib
