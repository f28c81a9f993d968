"""ctypes front-end of the CPU oracle (oracle/pp_oracle.c).

TEST INFRASTRUCTURE ONLY -- the parity checker and the "port" CPU baseline.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import
this package; the product package never does (tests/test_abi_exports.py::test_product_never_imports_oracle enforces it).

Function names and argument meaning mirror the reference (file:line in pp_oracle.c).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpp_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile pp_oracle.c with gcc (seconds)."""
    src = os.path.join(_HERE, "pp_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libpp_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.ppo_rotate_iou_pair.restype = C.c_double
        _lib.ppo_full_path_batch.restype = C.c_int64
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _d3(x):
    return (C.c_double * len(x))(*[float(v) for v in x])


def arith_is_f32(points, voxel_size, coors_range) -> bool:
    """numba's promotion at load_data.py:622: float32 only if every operand is float32.
    Python lists are cast to points.dtype by the wrapper (load_data.py:726-729)."""
    def dt(x):
        return x.dtype if isinstance(x, np.ndarray) else points.dtype
    return points.dtype == np.float32 and dt(voxel_size) == np.float32 and dt(coors_range) == np.float32


def grid_size(voxel_size, coors_range, arith_f32=False):
    g = (C.c_int32 * 3)()
    lib().ppo_grid_size(_d3(voxel_size), _d3(coors_range), int(arith_f32), g)
    return [int(v) for v in g]


def points_to_voxel(points, voxel_size, coors_range, max_points, reverse_index, max_voxels,
                    return_slots=False):
    """load_data.py:695-771."""
    points = np.ascontiguousarray(points)
    assert points.dtype in (np.float32, np.float64) and points.ndim == 2
    f32 = arith_is_f32(points, voxel_size, coors_range)
    if not isinstance(voxel_size, np.ndarray):
        voxel_size = np.array(voxel_size, dtype=points.dtype)
    if not isinstance(coors_range, np.ndarray):
        coors_range = np.array(coors_range, dtype=points.dtype)
    N, D = points.shape
    voxels = np.empty((max_voxels, max_points, D), points.dtype)
    coors = np.empty((max_voxels, 3), np.int32)
    num = np.empty((max_voxels,), np.int32)
    slots = np.empty((N,), np.int32) if return_slots else None
    m = lib().ppo_points_to_voxel(_p(points), int(points.dtype == np.float64), C.c_int64(N), D,
                                  _d3(voxel_size), _d3(coors_range), int(f32), int(max_points),
                                  int(max_voxels), int(bool(reverse_index)), _p(voxels), _p(coors),
                                  _p(num), _p(slots) if return_slots else None)
    assert m >= 0
    out = (voxels[:m], coors[:m], num[:m])
    return out + (slots,) if return_slots else out


def decorate(voxels, num_points, coors, vx, vy, x_offset, y_offset):
    """model/pointpillars.py:143-203."""
    voxels = np.ascontiguousarray(voxels, np.float32)
    num_points = np.ascontiguousarray(num_points, np.int32)
    coors = np.ascontiguousarray(coors, np.int32)
    M, P, D = voxels.shape
    out = np.empty((M, P, D + 5), np.float32)
    with np.errstate(all="ignore"):
        lib().ppo_decorate(_p(voxels), _p(num_points), _p(coors), C.c_int64(M), P, D,
                           C.c_double(vx), C.c_double(vy), C.c_double(x_offset),
                           C.c_double(y_offset), _p(out))
    return out


def scatter(voxel_features, coords, batch_size, ny, nx, layout="NCHW"):
    """model/pointpillars.py:285-341."""
    f = np.ascontiguousarray(voxel_features, np.float32)
    c = np.ascontiguousarray(coords, np.int32)
    M, Cc = f.shape
    nhwc = layout == "NHWC"
    out = np.empty((batch_size, ny, nx, Cc) if nhwc else (batch_size, Cc, ny, nx), np.float32)
    lib().ppo_scatter(_p(f), _p(c), C.c_int64(M), Cc, batch_size, ny, nx, int(nhwc), _p(out))
    return out


def second_box_decode(box_encodings, anchors):
    """libraries/eval_helper_functions.py:388-461 (default flags)."""
    e = np.ascontiguousarray(box_encodings, np.float32).reshape(-1, 7)
    a = np.ascontiguousarray(anchors, np.float32).reshape(-1, 7)
    out = np.empty_like(e)
    lib().ppo_second_box_decode(_p(e), _p(a), C.c_int64(e.shape[0]), _p(out))
    return out.reshape(np.shape(box_encodings))


def rbox_to_standup(boxes):
    """load_data.py:1525-1594 + 1330-1341 as called at model/voxelnet.py:1233-1249."""
    b = np.ascontiguousarray(boxes, np.float32)
    out = np.empty((b.shape[0], 4), np.float32)
    lib().ppo_rbox_to_standup(_p(b), C.c_int64(b.shape[0]), _p(out))
    return out


def argsort_desc(scores):
    s = np.ascontiguousarray(scores, np.float32)
    o = np.empty(s.shape[0], np.int32)
    lib().ppo_argsort_desc(_p(s), C.c_int64(s.shape[0]), _p(o))
    return o


def nms_postprocess(mask, n):
    mask = np.ascontiguousarray(mask, np.uint64)
    keep = np.empty(max(n, 1), np.int32)
    k = lib().ppo_nms_postprocess(_p(mask), C.c_int64(n), _p(keep))
    return keep[:k]


def nms(bboxes, scores, pre_max_size=None, post_max_size=None, iou_threshold=0.5):
    """libraries/eval_helper_functions.py:463-492; returns None when nothing is kept."""
    b = np.ascontiguousarray(bboxes, np.float32)
    s = np.ascontiguousarray(scores, np.float32)
    keep = np.empty(max(b.shape[0], 1), np.int64)
    k = lib().ppo_nms_standup(_p(b), _p(s), C.c_int64(b.shape[0]),
                              -1 if pre_max_size is None else int(pre_max_size),
                              -1 if post_max_size is None else int(post_max_size),
                              C.c_float(iou_threshold), _p(keep))
    return None if k == 0 else keep[:k].copy()


def standup_mask(boxes_sorted, thresh):
    b = np.ascontiguousarray(boxes_sorted, np.float32)
    n = b.shape[0]
    mask = np.empty(n * ((n + 63) // 64), np.uint64)
    lib().ppo_standup_mask(_p(b), C.c_int64(n), C.c_float(thresh), _p(mask))
    return mask


def rotate_mask(dets_sorted, thresh):
    d = np.ascontiguousarray(dets_sorted, np.float32)
    n = d.shape[0]
    mask = np.empty(n * ((n + 63) // 64), np.uint64)
    lib().ppo_rotate_mask(_p(d), C.c_int64(n), C.c_float(thresh), _p(mask))
    return mask


def rotate_nms_gpu(dets, nms_overlap_thresh, pre_max_size=None, post_max_size=None):
    """second/core/non_max_suppression/nms_gpu.py:455-490 (+ nms()'s optional caps)."""
    d = np.ascontiguousarray(dets, np.float32)
    keep = np.empty(max(d.shape[0], 1), np.int64)
    k = lib().ppo_rotate_nms(_p(d), C.c_int64(d.shape[0]), C.c_float(nms_overlap_thresh),
                             -1 if pre_max_size is None else int(pre_max_size),
                             -1 if post_max_size is None else int(post_max_size), _p(keep))
    return [int(v) for v in keep[:k]]


def rotate_iou_gpu_eval(boxes, query_boxes, criterion=-1):
    """nms_gpu.py:618-653 (criterion -1 == rotate_iou_gpu, 526-561)."""
    b = np.ascontiguousarray(boxes, np.float32)
    q = np.ascontiguousarray(query_boxes, np.float32)
    out = np.zeros((b.shape[0], q.shape[0]), np.float32)
    if b.shape[0] and q.shape[0]:
        with np.errstate(all="ignore"):
            lib().ppo_rotate_iou(_p(b), C.c_int64(b.shape[0]), _p(q), C.c_int64(q.shape[0]),
                                 int(criterion), _p(out))
    return out


def set_exact_trig(on: bool):
    """Evaluate the box rotation's sin/cos in float64 and round (what the CUDA path does) instead of the host libm's
    sinf/cosf (1 ulp off in ~1 % of the arguments): isolates that one source of float difference."""
    lib().ppo_set_exact_trig(int(bool(on)))


def rotate_iou_pair(r1, r2, criterion=-1) -> float:
    a = np.ascontiguousarray(r1, np.float32)
    b = np.ascontiguousarray(r2, np.float32)
    return float(lib().ppo_rotate_iou_pair(_p(a), _p(b), int(criterion)))


def rbbox2d_to_near_bbox(rbboxes):
    """load_data.py:534-548."""
    r = np.ascontiguousarray(rbboxes, np.float32)
    out = np.empty((r.shape[0], 4), np.float32)
    lib().ppo_rbbox2d_to_near_bbox(_p(r), C.c_int64(r.shape[0]), _p(out))
    return out


def anchor_cells(anchors, voxel_size, coors_range):
    a = np.ascontiguousarray(anchors, np.float32).reshape(-1, 7)
    g = (C.c_int32 * 3)(*grid_size(voxel_size, coors_range))
    out = np.empty((a.shape[0], 4), np.int32)
    lib().ppo_anchor_cells(_p(a), C.c_int64(a.shape[0]), _d3(voxel_size), _d3(coors_range), g, _p(out))
    return out


def anchors_mask(coors, anchors, voxel_size, coors_range, threshold=1):
    """load_data.py:3043-3072 for one frame: coors [M,3] (z,y,x) -> (anchors_area f32 [A], mask bool [A])."""
    co = np.ascontiguousarray(coors, np.int32)
    cells = anchor_cells(anchors, voxel_size, coors_range)
    nx, ny, _ = grid_size(voxel_size, coors_range)
    A = cells.shape[0]
    area = np.empty((A,), np.float32)
    mask = np.empty((A,), np.uint8)
    lib().ppo_anchors_mask(_p(co), C.c_int64(co.shape[0]), ny, nx, _p(cells), C.c_int64(A), C.c_float(threshold),
                           _p(area), _p(mask))
    return area, mask.astype(bool)


def d3_box_overlap(boxes, qboxes, criterion=-1):
    """second/utils/eval.py:159-163 (float64 camera boxes [N,7], [K,7]) -> [N,K] float32."""
    b = np.ascontiguousarray(boxes, np.float64)
    q = np.ascontiguousarray(qboxes, np.float64)
    out = np.zeros((b.shape[0], q.shape[0]), np.float32)
    if b.shape[0] and q.shape[0]:
        lib().ppo_d3_box_overlap(_p(b), C.c_int64(b.shape[0]), _p(q), C.c_int64(q.shape[0]), int(criterion), _p(out))
    return out


def full_path_batch(points, frame_off, voxel_size, coors_range, max_points, max_voxels, pfn_feats,
                    box_enc, anchors, scores, pre_max, post_max, thresh, rotated=True, nthreads=0):
    """CPU baseline of the whole path over a batch (bench.py only)."""
    points = np.ascontiguousarray(points)
    fo = np.ascontiguousarray(frame_off, np.int64)
    B = fo.shape[0] - 1
    pf = np.ascontiguousarray(pfn_feats, np.float32)
    be = np.ascontiguousarray(box_enc, np.float32)
    an = np.ascontiguousarray(anchors, np.float32)
    sc = np.ascontiguousarray(scores, np.float32)
    A = an.shape[0]
    det = np.zeros((B, post_max, 8), np.float32)
    cnt = np.zeros((B,), np.int32)
    vc = np.zeros((B,), np.int64)
    lib().ppo_full_path_batch(_p(points), int(points.dtype == np.float64), _p(fo), B,
                              points.shape[1], _d3(voxel_size), _d3(coors_range), int(max_points),
                              int(max_voxels), _p(pf), pf.shape[1], _p(be), _p(an), _p(sc),
                              C.c_int64(A), int(pre_max), int(post_max), C.c_float(thresh),
                              int(rotated), int(nthreads), _p(det), _p(cnt), _p(vc))
    return det, cnt, vc


def predict_frame(box_preds, cls_preds, dir_preds, anchors, a_mask, rect, Trv2c, top_k=100, pre_max_size=100,
                  post_max_size=50, iou_threshold=0.5, score_threshold=0.0, rotated=False,
                  use_direction_classifier=True):
    """Per-frame body of VoxelNet.predict, model/voxelnet.py:1105-1326 -> dict with the reference's keys
    (+ 'anchor_index'), or all-None boxes when nothing is kept."""
    bp = np.ascontiguousarray(box_preds, np.float32).reshape(-1, 7)
    A = bp.shape[0]
    cl = np.ascontiguousarray(cls_preds, np.float32).reshape(A, -1)
    dp = None if dir_preds is None else np.ascontiguousarray(dir_preds, np.float32).reshape(A, 2)
    an = np.ascontiguousarray(anchors, np.float32).reshape(A, 7)
    am = None if a_mask is None else np.ascontiguousarray(a_mask, np.uint8).reshape(A)
    rc = None if rect is None else np.ascontiguousarray(rect, np.float32).reshape(16)
    tv = None if Trv2c is None else np.ascontiguousarray(Trv2c, np.float32).reshape(16)
    cap = min(A, top_k) if A else 1
    lid = np.zeros((cap, 7), np.float32)
    cam = np.zeros((cap, 7), np.float64)
    sc = np.zeros((cap,), np.float32)
    lab = np.zeros((cap,), np.int32)
    idx = np.zeros((cap,), np.int32)
    pn = lambda a: None if a is None else _p(a)  # noqa: E731
    with np.errstate(all="ignore"):
        k = lib().ppo_predict_frame(_p(bp), _p(cl), pn(dp), _p(an), pn(am), pn(rc), pn(tv), C.c_int64(A), cl.shape[1],
                                    int(bool(use_direction_classifier)), int(top_k),
                                    -1 if pre_max_size is None else int(pre_max_size),
                                    -1 if post_max_size is None else int(post_max_size), C.c_float(iou_threshold),
                                    C.c_float(score_threshold), int(bool(rotated)), cap, _p(lid), _p(cam), _p(sc),
                                    _p(lab), _p(idx))
    if k == 0:
        return {"box3d_lidar": None, "box3d_camera": None, "scores": None, "label_preds": None, "anchor_index": None}
    return {"box3d_lidar": lid[:k], "box3d_camera": cam[:k] if rc is not None else None, "scores": sc[:k],
            "label_preds": lab[:k].astype(np.int64), "anchor_index": idx[:k]}


def pointcloud2_to_lidar(xyz, rotations, translation, start=1, step=4):
    """load_data.py:2434-2443 restated with numpy (the reference's own expressions; numpy is the arithmetic it
    uses).  `pointcloud2_to_xyz_array` belongs to ros_numpy (third party, not vendored, unpinned in
    configs/pip/requirements_short.txt); its published algorithm -- get_xyz_points(remove_nans=True, dtype=float):
    keep rows whose x, y and z are all finite, widen to float64 -- is restated in the first two lines.
    xyz: [N,3] float32 sensor points."""
    xyz = np.asarray(xyz, np.float32)
    mask = np.isfinite(xyz[:, 0]) & np.isfinite(xyz[:, 1]) & np.isfinite(xyz[:, 2])
    points = xyz[mask].astype(np.float64)[start::step]
    for r in rotations:
        points = np.dot(points, np.asarray(r, np.float64))
    if translation is not None:
        points = points + np.asarray(translation, np.float64)
    return points
