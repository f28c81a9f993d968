"""Drop-ins for the reference's NMS / rotated IoU entry points.

  nms               libraries/eval_helper_functions.py:463-492 (the live predict() path)
  nms_gpu           libraries/eval_helper_functions.py:494-527 / nms_gpu.py:131-165
  rotate_nms_gpu    second/core/non_max_suppression/nms_gpu.py:455-490
  rotate_iou_gpu    nms_gpu.py:526-561
  rotate_iou_gpu_eval  nms_gpu.py:618-653
Score ties (undefined in the reference: unstable argsort) are ordered by descending index.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _nms(kind, boxes, scores, pre_max_size, post_max_size, thresh, device):
    b = np.ascontiguousarray(boxes, np.float32)
    s = np.ascontiguousarray(scores, np.float32)
    N = b.shape[0]
    keep = np.empty((max(N, 1),), np.int64)
    k = C.c_int32(0)
    c = _lib.ctx(device)
    _lib.check(_lib.lib().pp_nms_host(c.handle, kind, _lib.ptr(b), _lib.ptr(s), N,
                                      -1 if pre_max_size is None else int(pre_max_size),
                                      -1 if post_max_size is None else int(post_max_size),
                                      float(thresh), _lib.ptr(keep), C.byref(k)))
    return keep[:k.value]


def nms(bboxes, scores, pre_max_size=None, post_max_size=None, iou_threshold=0.5, device=None):
    """bboxes [N,4] (xmin,ymin,xmax,ymax), scores [N] -> int64 indices or None when empty."""
    bboxes = np.asarray(bboxes)
    if bboxes.ndim != 2 or bboxes.shape[1] != 4:
        raise ValueError("bboxes must be [N,4]")
    keep = _nms(_lib.PP_NMS_STANDUP, bboxes, scores, pre_max_size, post_max_size, iou_threshold, device)
    return None if keep.shape[0] == 0 else keep


def nms_gpu(dets, nms_overlap_thresh, device_id=0):
    """dets [N,5] (xmin,ymin,xmax,ymax,score) -> list of kept indices."""
    dets = np.asarray(dets, np.float32)
    return [int(v) for v in _nms(_lib.PP_NMS_STANDUP, dets[:, :4], dets[:, 4], None, None, nms_overlap_thresh, device_id)]


def rotate_nms_gpu(dets, nms_overlap_thresh, device_id=0, pre_max_size=None, post_max_size=None):
    """dets [N,6] (x,y,w,l,angle,score) -> list of kept indices in keep order."""
    dets = np.asarray(dets).astype(np.float32)
    if dets.ndim != 2 or dets.shape[1] != 6:
        raise ValueError("dets must be [N,6]")
    return [int(v) for v in _nms(_lib.PP_NMS_ROTATED, dets[:, :5], dets[:, 5], pre_max_size, post_max_size,
                                 nms_overlap_thresh, device_id)]


def rotate_iou_gpu_eval(boxes, query_boxes, criterion=-1, device_id=0):
    """boxes [N,5], query_boxes [K,5] -> [N,K] float32 (criterion as nms_gpu.py:564-576)."""
    boxes = np.asarray(boxes)
    box_dtype = boxes.dtype
    b = np.ascontiguousarray(boxes.astype(np.float32))
    q = np.ascontiguousarray(np.asarray(query_boxes).astype(np.float32))
    N, K = b.shape[0], q.shape[0]
    iou = np.zeros((N, K), np.float32)
    if N == 0 or K == 0:
        return iou
    c = _lib.ctx(device_id)
    _lib.check(_lib.lib().pp_rotate_iou_host(c.handle, _lib.ptr(b), N, _lib.ptr(q), K, int(criterion), _lib.ptr(iou)))
    del box_dtype  # the reference casts back to boxes.dtype AFTER converting boxes to float32 (lines 561, 653)
    return iou


def rotate_iou_gpu(boxes, query_boxes, device_id=0):
    return rotate_iou_gpu_eval(boxes, query_boxes, -1, device_id)


def bev_box_overlap(boxes, qboxes, criterion=-1):
    """second/utils/eval.py:126-128."""
    return rotate_iou_gpu_eval(boxes, qboxes, criterion)


def d3_box_overlap(boxes, qboxes, criterion=-1, device_id=0):
    """second/utils/eval.py:159-163: camera boxes [N,7], [K,7] (x,y,z,l,h,w,ry) -> [N,K] float32 3-D overlap.
    The height/ratio arithmetic is float64 (what numba does for float64 annotations; float32 inputs are
    widened exactly)."""
    b = np.ascontiguousarray(np.asarray(boxes), np.float64)
    q = np.ascontiguousarray(np.asarray(qboxes), np.float64)
    out = np.zeros((b.shape[0], q.shape[0]), np.float32)
    if b.shape[0] == 0 or q.shape[0] == 0:
        return out
    c = _lib.ctx(device_id)
    _lib.check(_lib.lib().pp_d3_box_overlap_host(c.handle, _lib.ptr(b), b.shape[0], _lib.ptr(q), q.shape[0], int(criterion),
                                                 _lib.ptr(out)))
    return out
