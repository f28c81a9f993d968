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


@pytest.fixture(params=["table_path", "anygrid_path"])
def vox_path(request):
    """Both implementations behind pp_voxelize_dev must pass the voxelizer tests: the shared-memory table path (forced
    for every batch size the grid allows) and the any-grid path (forced for every batch); the default threshold between
    them is restored afterwards."""
    _lib = importlib.import_module(PKG + "._lib")
    n = 0 if request.param == "table_path" else 1 << 62
    _lib.check(_lib.lib().pp_voxelize_set_small_path_min_points(n))
    yield request.param
    _lib.check(_lib.lib().pp_voxelize_set_small_path_min_points(-1))


def golden(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN, name))
