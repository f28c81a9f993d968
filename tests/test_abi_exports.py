"""CPU-only: the C-ABI library builds, loads, and exports every symbol include/pp_b200.h declares;
the ctypes table covers the same set; compute entries fail loudly without a GPU; the product
package never touches oracle/."""
import ctypes
import importlib
import os
import re

import pytest

from conftest import PKG, ROOT


def header_symbols():
    src = open(os.path.join(ROOT, "include", "pp_b200.h")).read()
    return sorted(set(re.findall(r"PP_API[^;(]*?\b(pp_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    build = importlib.import_module(PKG + ".build")
    path = build.build()
    lib = ctypes.CDLL(path)
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in pp_b200.h but not exported"
    _lib = importlib.import_module(PKG + "._lib")
    assert sorted(_lib.SIGNATURES) == syms


def test_no_gpu_fails_loudly(pp):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    with pytest.raises(pp.PPError):
        pp.points_to_voxel(np.zeros((4, 3), np.float32), [1, 1, 1], [0, 0, 0, 2, 2, 2], 5, True, 10)


def test_grid_size_matches_numpy(pp, synth):
    import numpy as np
    for cfg in (synth.D435, synth.KITTI):
        assert pp.grid_size(cfg["voxel_size"], cfg["point_cloud_range"]) == synth.grid_size(cfg)
    r = np.array([0, 0, 0, 2.5, 3.5, 10.0]); v = np.array([1.0, 1.0, 4.0])
    assert pp.grid_size(v, r) == np.round((r[3:] - r[:3]) / v).astype(np.int32).tolist()


def test_product_never_imports_oracle():
    pkg_dir = os.path.join(ROOT, PKG)
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "libpp_oracle" not in text and "pp_oracle" not in text, f
