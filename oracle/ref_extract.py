"""Load the reference's OWN source for the hot path, in the build container only.

TEST INFRASTRUCTURE.  `/root/reference` exists only in the build container, never on the GPU
box, so this module is used solely by `tests/golden/make_golden.py` (fixture generation) and by
`tests/test_oracle_vs_reference.py` (skipped when the reference is absent).  Nothing is copied
into the repo: the functions are AST-extracted from the reference files where they lie and
exec'd in a scratch namespace, because importing the modules pulls TensorFlow / rospy / a
CPython-3.6 `nms.so` that do not exist here.

  * voxelizer       load_data.py:558-771   (numba CPU jit, runs unmodified)
  * standup prep    load_data.py:1330-1341, 1525-1594
  * decode, sweep   libraries/eval_helper_functions.py:388-461, 529-550
  * rotated IoU     second/core/non_max_suppression/nms_gpu.py:180-415, 564-576 -- numba.cuda device
                    functions.  There is no GPU here, so the `@cuda.jit(..., device=True)`
                    decorators are textually swapped for `@numba.njit` and `cuda.local.array`
                    for `np.empty`; bodies, operation order and numba's type inference (the
                    f32/f64 promotion map of SURVEY 3.5) are untouched.
  * standup IoU     eval_helper_functions.py:553-564, same mechanical swap.
  * decoration, scatter   model/pointpillars.py:23-49, 128-203, 285-341: the method bodies run unmodified with
                    oracle/tf_shim.py (numpy) bound to `tf` -- TensorFlow itself cannot be installed here
  * predict         model/voxelnet.py:1060-1389 (method body, run with a stand-in `self`; the numpy-1.19
                    `x[[index_array]]` idiom is rewritten to `x[index_array]`, see _numpy119_indexing)
"""
from __future__ import annotations

import ast
import os
import re
import types

REFERENCE_ROOT = os.environ.get("PP_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "load_data.py"))


def _read(rel):
    with open(os.path.join(REFERENCE_ROOT, rel), encoding="utf-8-sig") as f:
        return f.read()


def _extract_defs(src: str, names) -> str:
    """Source text (decorators included) of the named top-level functions, in file order."""
    tree = ast.parse(src)
    lines = src.splitlines()
    out = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            start = min([node.lineno] + [d.lineno for d in node.decorator_list]) - 1
            out.append("\n".join(lines[start:node.end_lineno]))
    missing = set(names) - {n.name for n in tree.body if isinstance(n, ast.FunctionDef)}
    if missing:
        raise RuntimeError(f"reference functions not found: {sorted(missing)}")
    return "\n\n".join(out) + "\n"


_CUDA_DECORATOR = re.compile(r"@cuda\.jit\((?:[^()]|\([^()]*\))*\)", re.S)
_LOCAL_ARRAY = re.compile(r"cuda\.local\.array\(\((\d+),\s*\),\s*dtype=numba\.float32\)")


def _cuda_device_to_njit(src: str) -> str:
    src = _CUDA_DECORATOR.sub('@numba.njit(error_model="numpy")', src)
    return _LOCAL_ARRAY.sub(r"np.empty(\1, np.float32)", src)


_LIST_OF_ARRAY_INDEX = re.compile(r"\[\[(\w+)\]\]")


def _numpy119_indexing(src: str) -> str:
    """`x[[idx]]` with idx an index ARRAY: numpy 1.19 (the reference's pin) read the one-element list as a
    tuple, i.e. x[idx]; numpy >= 1.23 reads it as a new leading axis.  Rewritten textually to `x[idx]`."""
    return _LIST_OF_ARRAY_INDEX.sub(r"[\1]", src)


def _extract_method(src: str, cls: str, name: str) -> str:
    """Source of method `name` of top-level class `cls`, dedented to a plain function."""
    import textwrap
    tree = ast.parse(src)
    lines = src.splitlines()
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for m in node.body:
                if isinstance(m, ast.FunctionDef) and m.name == name:
                    return textwrap.dedent("\n".join(lines[m.lineno - 1:m.end_lineno])) + "\n"
    raise RuntimeError(f"reference method {cls}.{name} not found")


_cache = None


def load() -> types.SimpleNamespace:
    """Namespace with the reference's functions, compiled from the reference's files."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    import math

    import numba
    import numpy as np

    ns = {"np": np, "numba": numba, "math": math}

    exec(_extract_defs(_read("load_data.py"), [
        "_points_to_voxel_reverse_kernel", "_points_to_voxel_kernel", "points_to_voxel",
        "corner_to_standup_nd_jit", "center_to_corner_box2d", "rotation_2d", "corners_nd",
        "create_anchors_3d_stride", "sparse_sum_for_anchors_mask", "fused_get_anchors_area",
        "rbbox2d_to_near_bbox", "limit_period", "center_to_minmax_2d", "center_to_minmax_2d_0_5",
    ]), ns)

    def anchors_mask(coors, anchors, voxel_size, point_cloud_range, threshold):
        """load_data.py:3043-3072 verbatim call sequence for one frame."""
        voxel_size = np.asarray(voxel_size, np.float64)
        pcr = np.asarray(point_cloud_range, np.float64)
        grid_size = np.round((pcr[3:] - pcr[:3]) / voxel_size).astype(np.int64)
        anchors_bv = ns["rbbox2d_to_near_bbox"](anchors[:, [0, 1, 3, 4, 6]])
        dense = ns["sparse_sum_for_anchors_mask"](coors, tuple(grid_size[::-1][1:]))
        dense = dense.cumsum(0)
        dense = dense.cumsum(1)
        area = ns["fused_get_anchors_area"](dense, anchors_bv, voxel_size, pcr, grid_size)
        return area, area > threshold

    def create_anchors(feature_size, sizes, strides, offsets, rotations):
        """create_anchors_3d_stride under numpy>=2 (np.meshgrid returns a tuple there; the reference
        assigns into it, so hand it a list-returning meshgrid for the duration of the call)."""
        real = np.meshgrid
        try:
            np.meshgrid = lambda *a, **k: list(real(*a, **k))
            return ns["create_anchors_3d_stride"](feature_size, sizes, strides, offsets, rotations)
        finally:
            np.meshgrid = real

    ehf = _read("libraries/eval_helper_functions.py")
    exec(_extract_defs(ehf, ["second_box_decode", "nms_postprocess", "div_up"]), ns)
    standup = {"np": np, "numba": numba, "math": math}
    exec(_cuda_device_to_njit(_extract_defs(ehf, ["iou_device"])), standup)
    ns["iou_device"] = standup["iou_device"]

    rot = {"np": np, "numba": numba, "math": math}
    exec(_cuda_device_to_njit(_extract_defs(_read("second/core/non_max_suppression/nms_gpu.py"), [
        "trangle_area", "area", "sort_vertex_in_convex_polygon", "line_segment_intersection",
        "point_in_quadrilateral", "quadrilateral_intersection", "rbbox_to_corners", "inter",
        "devRotateIoU", "devRotateIoUEval",
    ])), rot)

    dev_eval = rot["devRotateIoUEval"]
    dev_iou = rot["devRotateIoU"]
    iou_dev = ns["iou_device"]

    # Drivers: the kernels' indexing (nms_gpu.py:445-449, 520-523, 611-615;
    # eval_helper_functions.py:589-596) restated as plain loops around the reference's own
    # device functions.
    @numba.njit(error_model="numpy")
    def rotate_iou_matrix(boxes, qboxes, criterion):
        N, K = boxes.shape[0], qboxes.shape[0]
        out = np.zeros((N, K), np.float32)
        for n in range(N):
            for k in range(K):
                out[n, k] = dev_eval(qboxes[k], boxes[n], criterion)
        return out

    @numba.njit(error_model="numpy")
    def rotate_iou_matrix_f64(boxes, qboxes):
        N, K = boxes.shape[0], qboxes.shape[0]
        out = np.zeros((N, K), np.float64)
        for n in range(N):
            for k in range(K):
                out[n, k] = dev_iou(qboxes[k], boxes[n])
        return out

    @numba.njit(error_model="numpy")
    def rotate_mask(dets, thresh):
        n = dets.shape[0]
        cb = (n + 63) // 64
        mask = np.zeros(n * cb, np.uint64)
        iou_all = np.zeros((n, n), np.float64)
        for i in range(n):
            for j in range(i + 1, n):
                iou = dev_iou(dets[i, :5], dets[j, :5])
                iou_all[i, j] = iou
                if iou > thresh:
                    mask[i * cb + j // 64] |= np.uint64(1) << np.uint64(j % 64)
        return mask, iou_all

    @numba.njit(error_model="numpy")
    def standup_mask(boxes, thresh):
        n = boxes.shape[0]
        cb = (n + 63) // 64
        mask = np.zeros(n * cb, np.uint64)
        iou_all = np.zeros((n, n), np.float64)
        for i in range(n):
            for j in range(i + 1, n):
                iou = iou_dev(boxes[i, :4], boxes[j, :4])
                iou_all[i, j] = iou
                if iou > thresh:
                    mask[i * cb + j // 64] |= np.uint64(1) << np.uint64(j % 64)
        return mask, iou_all

    def rotate_nms(dets, thresh):
        """rotate_nms_gpu (nms_gpu.py:455-490) with the kernel replaced by the loop above.
        Uses a stable argsort so tie order is defined (the reference's is not)."""
        dets = dets.astype(np.float32)
        n = dets.shape[0]
        order = dets[:, 5].argsort(kind="stable")[::-1].astype(np.int32)
        mask, iou_all = rotate_mask(np.ascontiguousarray(dets[order]), np.float32(thresh))
        keep = np.zeros(n, np.int32)
        nk = ns["nms_postprocess"](keep, mask, n)
        return [int(v) for v in order[keep[:nk]]], iou_all

    def standup_nms(dets, thresh):
        """nms_gpu (eval_helper_functions.py:494-527), same substitution."""
        dets = dets.astype(np.float32)
        n = dets.shape[0]
        order = dets[:, 4].argsort(kind="stable")[::-1].astype(np.int32)
        mask, iou_all = standup_mask(np.ascontiguousarray(dets[order]), np.float32(thresh))
        keep = np.zeros(n, np.int32)
        nk = ns["nms_postprocess"](keep, mask, n)
        return [int(v) for v in order[keep[:nk]]], iou_all

    ev = {"np": np, "numba": numba, "math": math}
    exec(_extract_defs(_read("second/utils/eval.py"), ["d3_box_overlap_kernel"]), ev)

    def d3_box_overlap(boxes, qboxes, criterion=-1):
        """second/utils/eval.py:159-163 with rotate_iou_gpu_eval replaced by the CPU run of its device code."""
        rinc = rotate_iou_matrix(np.ascontiguousarray(boxes[:, [0, 2, 3, 5, 6]].astype(np.float32)),
                                 np.ascontiguousarray(qboxes[:, [0, 2, 3, 5, 6]].astype(np.float32)), 2)
        ev["d3_box_overlap_kernel"](boxes, qboxes, rinc, criterion)
        return rinc

    # ---- "next" row N2: VoxelNet.predict (model/voxelnet.py:1060-1389) run as the reference wrote it, with
    #      self.nms_func = nms (voxelnet.py:786, eval_helper_functions.py:463-492) whose numba.cuda nms_gpu is the
    #      CPU run of the same kernel body (standup_nms above).
    vn = _read("model/voxelnet.py")
    pred = {"np": np}
    exec(_extract_defs(_read("load_data.py"), ["lidar_to_camera", "box_lidar_to_camera"]), pred)
    pred.update(second_box_decode=ns["second_box_decode"], center_to_corner_box2d=ns["center_to_corner_box2d"],
                corner_to_standup_nd_jit=ns["corner_to_standup_nd_jit"])
    exec(_extract_defs(vn, ["sigmoid_array"]), pred)
    exec(_numpy119_indexing(_extract_method(vn, "VoxelNet", "predict")), pred)
    nmsns = {"np": np, "nms_gpu": lambda dets, thr: standup_nms(dets, thr)[0]}
    exec(_numpy119_indexing(_extract_defs(ehf, ["nms"])), nmsns)

    class _T:  # stands in for an eager tensor: only .numpy() is used (voxelnet.py:1067-1084)
        def __init__(self, a):
            self.a = a

        def numpy(self):
            return self.a

    def predict(example, preds_dict, config):
        """example: tuple of numpy arrays laid out as the reference's (indices 3,4,5,6,7,8 are read);
        preds_dict: numpy arrays under box_preds / cls_preds / dir_cls_preds."""
        sec = config["model"]["second"]
        me = types.SimpleNamespace(
            num_class=sec["num_class"], encode_background_as_zeros=sec["encode_background_as_zeros"], config=config,
            box_code_size=7, use_direction_classifier=sec["use_direction_classifier"],
            use_multi_class_nms=sec["use_multi_class_nms"], nms_score_threshold=sec["nms_score_threshold"],
            nms_func=nmsns["nms"], nms_pre_max_size=sec["nms_pre_max_size"],
            nms_post_max_size=sec["nms_post_max_size"], nms_iou_threshold=sec["nms_iou_threshold"],
            measure_time_extended=False)
        ex = [None if a is None else _T(a) for a in example]
        pd = {k: _T(v) for k, v in preds_dict.items()}
        return pred["predict"](me, ex, pd)

    # ---- a4 / a5: PillarFeatureNet.call lines 143-203 and PointPillarsScatter.call (model/pointpillars.py) executed
    #      as written, with oracle/tf_shim.py (numpy) standing in for the dozen TensorFlow ops they call
    from . import tf_shim
    ppsrc = _read("model/pointpillars.py")
    pfn = {"np": np, "tf": tf_shim}
    exec(_extract_defs(ppsrc, ["get_paddings_indicator"]), pfn)
    exec(_extract_method(ppsrc, "PillarFeatureNet", "call"), pfn)
    pfn["pfn_call"] = pfn.pop("call")
    exec(_extract_method(ppsrc, "PointPillarsScatter", "call"), pfn)
    pfn["scatter_call"] = pfn.pop("call")

    def pillar_decorate(voxels, num_points, coors, voxel_size, point_cloud_range):
        """The tensor PillarFeatureNet.call hands to self.pfn_layer (model/pointpillars.py:211), i.e. the decorated and
        masked [M,P,D+5] features; constants as __init__ computes them (121-124)."""
        captured = {}

        def pfn_layer(x, training=False):
            captured["x"] = np.array(x)
            return x
        vx, vy = voxel_size[0], voxel_size[1]
        me = types.SimpleNamespace(vx=vx, vy=vy, x_offset=vx / 2 + point_cloud_range[0], y_offset=vy / 2 + point_cloud_range[1],
                                   with_distance=False, pfn_layer=pfn_layer, training=False)
        pfn["pfn_call"](me, tf_shim.convert(voxels), tf_shim.convert(num_points), tf_shim.convert(coors))
        return captured["x"]

    def pointpillars_scatter(voxel_features, coords, batch_size, nchannels, ny, nx):
        me = types.SimpleNamespace(batch_size=batch_size, nchannels=nchannels, ny=ny, nx=nx)
        return np.array(pfn["scatter_call"](me, tf_shim.convert(voxel_features), tf_shim.convert(coords)))

    _cache = types.SimpleNamespace(
        nms=nmsns["nms"],
        pillar_decorate=pillar_decorate,
        pointpillars_scatter=pointpillars_scatter,
        predict=predict,
        points_to_voxel=ns["points_to_voxel"],
        second_box_decode=ns["second_box_decode"],
        nms_postprocess=ns["nms_postprocess"],
        center_to_corner_box2d=ns["center_to_corner_box2d"],
        corner_to_standup_nd_jit=ns["corner_to_standup_nd_jit"],
        create_anchors_3d_stride=create_anchors,
        anchors_mask=anchors_mask,
        d3_box_overlap=d3_box_overlap,
        rbbox2d_to_near_bbox=ns["rbbox2d_to_near_bbox"],
        rotate_iou_matrix=rotate_iou_matrix,
        rotate_iou_matrix_f64=rotate_iou_matrix_f64,
        rotate_nms=rotate_nms,
        standup_nms=standup_nms,
    )
    return _cache
