"""Drop-in for the data loader's anchor mask, load_data.py:3043-3072."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def anchors_mask(coordinates, anchors, voxel_size, point_cloud_range, anchor_area_threshold=1, device=None):
    """coordinates [M,3] int32 (z,y,x) as points_to_voxel returns them, anchors [A,7] float32 ->
    (anchors_area float32 [A], anchors_mask bool [A]); the reference's sequence
    rbbox2d_to_near_bbox -> sparse_sum_for_anchors_mask -> cumsum(0).cumsum(1) -> fused_get_anchors_area ->
    `anchors_area > anchor_area_threshold`."""
    co = np.ascontiguousarray(coordinates, np.int32)
    an = np.ascontiguousarray(anchors, np.float32).reshape(-1, 7)
    if co.ndim != 2 or co.shape[1] != 3:
        raise ValueError("coordinates must be [M,3]")
    A = an.shape[0]
    area = np.empty((A,), np.float32)
    mask = np.empty((A,), np.uint8)
    c = _lib.ctx(device)
    _lib.check(_lib.lib().pp_anchors_mask_host(
        c.handle, _lib.ptr(co), co.shape[0], _lib.ptr(an), A, (C.c_double * 3)(*map(float, voxel_size)),
        (C.c_double * 6)(*map(float, point_cloud_range)), float(anchor_area_threshold), _lib.ptr(area), _lib.ptr(mask)))
    return area, mask.astype(bool)
