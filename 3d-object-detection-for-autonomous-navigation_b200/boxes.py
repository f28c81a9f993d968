"""Drop-ins for the box stages of libraries/eval_helper_functions.py and load_data.py."""
from __future__ import annotations

import numpy as np

from . import _lib


def second_box_decode(box_encodings, anchors, encode_angle_to_vector=False, smooth_dim=False, device=None):
    """libraries/eval_helper_functions.py:388-461.  Only the default flags are on the hot path
    (model/voxelnet.py:1227); the other two raise."""
    if encode_angle_to_vector or smooth_dim:
        raise NotImplementedError("only the default flags (as called at model/voxelnet.py:1227) are accelerated")
    e = np.ascontiguousarray(box_encodings, np.float32)
    a = np.ascontiguousarray(anchors, np.float32)
    if e.shape[-1] != 7 or a.shape != e.shape:
        raise ValueError("box_encodings and anchors must both be [..., 7]")
    out = np.empty_like(e)
    c = _lib.ctx(device)
    _lib.check(_lib.lib().pp_box_decode_host(c.handle, _lib.ptr(e), _lib.ptr(a), e.size // 7, _lib.ptr(out)))
    return out


def rbox_to_standup(boxes, device=None):
    """corner_to_standup_nd_jit(center_to_corner_box2d(xy, wl, r)) as at model/voxelnet.py:1233-1249
    (load_data.py:1525-1594, 1330-1341).  boxes [N,5] (x,y,w,l,r) -> [N,4]."""
    b = np.ascontiguousarray(boxes, np.float32)
    if b.ndim != 2 or b.shape[1] != 5:
        raise ValueError("boxes must be [N,5]")
    out = np.empty((b.shape[0], 4), np.float32)
    c = _lib.ctx(device)
    _lib.check(_lib.lib().pp_rbox_to_standup_host(c.handle, _lib.ptr(b), b.shape[0], _lib.ptr(out)))
    return out
