import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG = "3d-object-detection-for-autonomous-navigation_b200"
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def synth():
    return importlib.import_module(PKG + ".synth")


@pytest.fixture(scope="session")
def pp():
    """The product package (loads the CUDA C-ABI library; fails loudly if it is not built)."""
    return importlib.import_module(PKG)


@pytest.fixture(scope="session")
def oracle():
    import oracle as o
    o.build()
    return o


def golden(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN, name))
