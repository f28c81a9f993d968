"""CPU-only: host-side logic of the drop-in layer that needs no device (record layouts, config parsing, frame
sharding arithmetic, reference rotation constants)."""
import importlib

import numpy as np
import pytest

from conftest import PKG


def test_pointcloud2_layouts():
    ing = importlib.import_module(PKG + ".ingest")
    xyz = np.arange(30, dtype=np.float32).reshape(10, 3)
    buf, n, ps, offs = ing._layout(xyz, None, None)
    assert (n, ps, offs) == (10, 12, (0, 4, 8)) and buf.dtype == np.float32
    xyzi = np.arange(40, dtype=np.float64).reshape(10, 4)           # wider rows, wrong dtype: converted, stride kept
    buf, n, ps, offs = ing._layout(xyzi, None, None)
    assert (n, ps, offs) == (10, 16, (0, 4, 8)) and buf.dtype == np.float32
    rec = np.zeros(7, dtype=np.dtype({"names": ["x", "y", "z", "rgb"], "formats": ["<f4"] * 4, "offsets": [0, 4, 8, 16], "itemsize": 20}))
    buf, n, ps, offs = ing._layout(rec, None, None)
    assert (n, ps, offs) == (7, 20, (0, 4, 8))
    buf, n, ps, offs = ing._layout(rec.tobytes(), 20, (0, 4, 8))
    assert (n, ps, offs) == (7, 20, (0, 4, 8)) and buf.dtype == np.uint8
    with pytest.raises(ValueError):
        ing._layout(rec.tobytes(), None, None)                       # raw bytes need point_step
    with pytest.raises(ValueError):
        ing._layout(np.zeros((5, 2), np.float32), None, None)
    bad = np.zeros(3, dtype=np.dtype([("x", "<f8"), ("y", "<f8"), ("z", "<f8")]))
    with pytest.raises(ValueError):
        ing._layout(bad, None, None)


def test_reference_rotation_constants_match_scipy():
    """R.from_euler('y', -90, degrees=True) / ('x', 90): the matrices of load_data.py:2438-2439 (as_dcm == as_matrix)."""
    from scipy.spatial.transform import Rotation as R
    ing = importlib.import_module(PKG + ".ingest")
    assert np.array_equal(R.from_euler("y", -90, degrees=True).as_matrix(), ing.R_Y_NEG90)
    assert np.array_equal(R.from_euler("x", 90, degrees=True).as_matrix(), ing.R_X_POS90)


def test_predict_config_parsing():
    pr = importlib.import_module(PKG + ".predict")
    _lib = importlib.import_module(PKG + "._lib")
    yaml_like = {"model": {"second": {"num_class": 1, "nms_pre_max_size": 100, "nms_post_max_size": 50, "nms_iou_threshold": 0.5}}}
    assert pr._second(yaml_like)["nms_post_max_size"] == 50
    assert pr._second({"nms_post_max_size": 7})["nms_post_max_size"] == 7 and pr._second(None) == {}
    c = pr.make_cfg(nms_pre_max_size=None, nms_post_max_size=None, rotated=True, anchors_per_frame=False)
    assert (c.nms_pre_max_size, c.nms_post_max_size, c.nms_kind, c.anchors_per_frame) == (-1, -1, _lib.PP_NMS_ROTATED, 0)
    c = pr.make_cfg()
    assert (c.num_class, c.top_k, c.nms_pre_max_size, c.nms_post_max_size, c.nms_kind) == (1, 100, 100, 50, _lib.PP_NMS_STANDUP)
    assert abs(c.nms_iou_threshold - 0.5) < 1e-7 and c.nms_score_threshold == 0.0


def test_shard_frames_partitions_exactly():
    pipeline = importlib.import_module(PKG + ".pipeline")
    for n in (0, 1, 7, 64, 512, 513):
        for ws in (1, 2, 3, 8):
            parts = [pipeline.shard_frames(n, ws, r) for r in range(ws)]
            assert sum(c for _, c in parts) == n
            assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(ws - 1)) and parts[0][0] == 0
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
