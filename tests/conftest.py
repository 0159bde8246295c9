import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def cb():
    """the package (its directory name has a hyphen, so no plain import statement)"""
    return importlib.import_module("corona-13_b200")


@pytest.fixture(scope="session")
def built():
    """make sure the oracle (and, where the reference tree exists, oracle/_ref) and the product are built"""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def lib(built):
    return importlib.import_module("corona-13_b200.lib")
