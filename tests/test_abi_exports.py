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


def test_argument_validation_needs_no_gpu(pp):
    """Bad arguments are rejected with an error code + message before any CUDA call."""
    import ctypes as C
    _lib = importlib.import_module(PKG + "._lib")
    L = _lib.lib()
    cfg = _lib.make_cfg([0.16, 0.16, 4.0], [0, -39.68, -3, 69.12, 39.68, 1], 100, 12000, True, False)
    one = C.c_void_p(16)
    args = [C.byref(cfg), one, 0, 4, one, 1, 10, 10, 0, one, None, one, 4, one, 100, one, one, None, None, one, 1 << 30, None]
    bad = list(args); bad[3] = 2          # D < 3
    assert L.pp_voxelize_dev(*bad) == -1 and b"D=2" in L.pp_last_error_string()
    bad = list(args); bad[12] = 5         # coors_cols
    assert L.pp_voxelize_dev(*bad) == -1
    bad = list(args); bad[8] = 1          # float64 output for float32 points
    assert L.pp_voxelize_dev(*bad) == -1
    bad = list(args); bad[20] = 16        # workspace too small
    assert L.pp_voxelize_dev(*bad) == -3 and b"workspace" in L.pp_last_error_string()
    assert L.pp_scatter_dev(one, one, 10, None, 64, 0, 4, 4, 0, one, one, 1 << 20, None) == -1
    assert L.pp_nms_dev(7, one, 5, one, None, 1, 10, -1, -1, 0.5, one, 10, one, one, 1 << 20, None) == -1
    assert L.pp_scatter_cells_dev(one, one, 5, 64, 1, 4, 4, 0, one, None) == -1 and b"slabs" in L.pp_last_error_string()   # nz > 4
    assert L.pp_scatter_cells_dev(one, None, 1, 64, 1, 4, 4, 0, one, None) == -1                                           # no cell map
    assert L.pp_rotate_iou_dev(one, 4, one, 4, 9, one, None) == -1
    # "next" rows: predict glue and sensor ingest
    pc = _lib.PredictCfg(1, 1, 100, 100, 50, 0, 0.5, 0.0, 1)
    pargs = [C.byref(pc), one, one, one, one, None, None, None, 2, 100, 50, one, None, one, one, one, one, one, 1 << 20, None]
    big = _lib.PredictCfg(1, 1, 500, -1, 50, 0, 0.5, 0.0, 1)                                    # 500 boxes: the general NMS path
    assert L.pp_predict_workspace_bytes(C.byref(big), 2, 100000, 50) > L.pp_predict_workspace_bytes(C.byref(pc), 2, 100000, 50) >= 800000
    bad = list(pargs); bad[0] = C.byref(big); bad[9] = 100000; bad[18] = L.pp_predict_workspace_bytes(C.byref(pc), 2, 100000, 50)
    assert L.pp_predict_dev(*bad) == -3
    bad = list(pargs); bad[6] = one                                                             # rect without Trv2c
    assert L.pp_predict_dev(*bad) == -1
    bad = list(pargs); bad[18] = 16
    assert L.pp_predict_dev(*bad) == -3
    rot = (C.c_double * 9)(1, 0, 0, 0, 1, 0, 0, 0, 1)
    iargs = [one, 1, 1000, 12, 0, 4, 8, 1, 4, rot, 1, None, one, 250, one, one, 1 << 20, None]
    bad = list(iargs); bad[3] = 10                                                              # point_step not a multiple of 4
    assert L.pp_ingest_dev(*bad) == -1
    bad = list(iargs); bad[6] = 12                                                              # z field outside the record
    assert L.pp_ingest_dev(*bad) == -1
    bad = list(iargs); bad[8] = 0                                                               # step 0
    assert L.pp_ingest_dev(*bad) == -1
    bad = list(iargs); bad[10] = 5                                                              # too many rotations
    assert L.pp_ingest_dev(*bad) == -1
    bad = list(iargs); bad[16] = 8
    assert L.pp_ingest_dev(*bad) == -3
    with pytest.raises(pp.PPError):
        _lib.check(-1)
    assert L.pp_voxelize_workspace_bytes(C.byref(cfg), 120000 * 64, 64, 120000, 4, 0) > 64 * 214272 * 16


def test_header_is_plain_c_and_exports_match():
    """include/pp_b200.h compiles as C99 (no C++ / CUDA / torch types at the boundary) and the library exports no
    pp_* symbol the header does not declare."""
    import subprocess
    import tempfile
    hdr = os.path.join(ROOT, "include", "pp_b200.h")
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "t.c")
        with open(src, "w") as f:
            f.write('#include "pp_b200.h"\nint main(void) { pp_voxel_cfg c; pp_predict_cfg p; (void)c; (void)p; return PP_OK; }\n')
        r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.dirname(hdr), src],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    build = importlib.import_module(PKG + ".build")
    out = subprocess.run(["nm", "-D", "--defined-only", build.build()], capture_output=True, text=True).stdout
    exported = sorted({ln.split()[-1] for ln in out.splitlines() if ln.split() and ln.split()[-1].startswith("pp_")})
    assert exported == header_symbols()


def test_voxelizer_workspace_covers_smaller_batches(pp, synth):
    """A workspace sized for the largest batch must serve every smaller one: the table path picks its chunk size (4 096 or
    16 384 points) from the batch it is handed, and pp_voxelize_dev falls back to the any-grid path when the bytes it is
    given do not cover the instance it would pick -- silently slower, so the sizing function has to be monotone."""
    import ctypes as C
    _lib = importlib.import_module(PKG + "._lib")
    L = _lib.lib()
    for c, n in ((synth.D435, 407040),     # 10 240 cells: table path eligible
                 (synth.KITTI, 120000)):   # 214 272 cells: any-grid path only
        cfg = _lib.make_cfg(c["voxel_size"], c["point_cloud_range"], c["max_points"], c["max_voxels"], True, False)
        big = L.pp_voxelize_workspace_bytes(C.byref(cfg), 64 * n, 64, n, 3, 0)
        assert big > 0
        prev = 0
        for k in (1, 2, 3, 6, 7, 8, 16, 33, 64):
            ws = L.pp_voxelize_workspace_bytes(C.byref(cfg), k * n, k, n, 3, 0)
            assert 0 < ws <= big, (k, ws, big)
            assert ws >= prev, (k, ws, prev)     # monotone in the batch size
            prev = ws
        # smaller frames inside the same capacity
        assert L.pp_voxelize_workspace_bytes(C.byref(cfg), 64 * 1000, 64, 1000, 3, 0) <= big
