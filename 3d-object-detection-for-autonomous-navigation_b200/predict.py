"""Drop-in for the post-network half of `VoxelNet.predict(example, preds_dict)`,
model/voxelnet.py:1060-1389 ("next" row N2): anchor-mask gather, sigmoid, top-100, decode, standup
boxes, NMS, direction flip and lidar->camera boxes for the whole batch in two kernel launches
(pp_predict_dev) instead of nine .numpy() syncs, a python loop over frames and a numba.cuda NMS call
per frame.

    predict(example, preds_dict, config)  ->  list of per-frame dicts with the reference's keys
        "bbox", "box3d_camera", "box3d_lidar", "scores", "label_preds", "batch_idx"

`example` is the reference's tuple (indices as at voxelnet.py:1061): [3] rect, [4] Trv2c, [6] anchors
[B,A,7], [7] anchors_mask [B,A], [8] image_idx; entries may be numpy arrays or anything with .numpy().
`config` is the reference's YAML dict (configs/train.yaml) or a flat dict with the nms_* keys.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _np(x):
    return x.numpy() if hasattr(x, "numpy") else np.asarray(x)


def make_cfg(num_class=1, use_direction_classifier=True, top_k=100, nms_pre_max_size=100, nms_post_max_size=50,
             nms_iou_threshold=0.5, nms_score_threshold=0.0, rotated=False, anchors_per_frame=True) -> _lib.PredictCfg:
    c = _lib.PredictCfg()
    c.num_class = int(num_class)
    c.use_direction_classifier = int(bool(use_direction_classifier))
    c.top_k = int(top_k)
    c.nms_pre_max_size = -1 if nms_pre_max_size is None else int(nms_pre_max_size)
    c.nms_post_max_size = -1 if nms_post_max_size is None else int(nms_post_max_size)
    c.nms_kind = _lib.PP_NMS_ROTATED if rotated else _lib.PP_NMS_STANDUP
    c.nms_iou_threshold = float(nms_iou_threshold)
    c.nms_score_threshold = float(nms_score_threshold)
    c.anchors_per_frame = int(bool(anchors_per_frame))
    return c


def _second(config):
    """The `model.second` block of configs/train.yaml (lines 121-179), or the dict itself when flat."""
    if config is None:
        return {}
    try:
        return config["model"]["second"]
    except (KeyError, TypeError):
        return config


def predict_arrays(box_preds, cls_preds, dir_preds, anchors, anchors_mask=None, rect=None, Trv2c=None, *,
                   num_class=1, use_direction_classifier=True, top_k=100, nms_pre_max_size=100,
                   nms_post_max_size=50, nms_iou_threshold=0.5, nms_score_threshold=0.0, rotated=False, device=None):
    """Batched arrays in, padded arrays out: (box3d_lidar [B,K,7] f32, box3d_camera [B,K,7] f64 or None,
    scores [B,K], label_preds [B,K] int32, anchor_index [B,K] int32, count [B])."""
    an = np.ascontiguousarray(_np(anchors), np.float32)
    per_frame = an.ndim == 3
    A = an.shape[-2]
    bp = np.ascontiguousarray(_np(box_preds), np.float32)
    B = bp.shape[0] if bp.ndim >= 3 or per_frame else 1
    bp = bp.reshape(B, -1, 7)
    if bp.shape[1] != A:
        raise ValueError(f"box_preds has {bp.shape[1]} rows per frame, anchors {A}")
    cl = np.ascontiguousarray(_np(cls_preds), np.float32).reshape(B, A, int(num_class))
    dp = None
    if use_direction_classifier:
        dp = np.ascontiguousarray(_np(dir_preds), np.float32).reshape(B, A, 2)
    am = None if anchors_mask is None else np.ascontiguousarray(_np(anchors_mask), np.uint8).reshape(B, A)
    rc = None if rect is None else np.ascontiguousarray(_np(rect), np.float32).reshape(B, 16)
    tv = None if Trv2c is None else np.ascontiguousarray(_np(Trv2c), np.float32).reshape(B, 16)
    cfg = make_cfg(num_class, use_direction_classifier, top_k, nms_pre_max_size, nms_post_max_size, nms_iou_threshold,
                   nms_score_threshold, rotated, per_frame)
    n_max = cfg.top_k if cfg.nms_pre_max_size <= 0 else min(cfg.top_k, cfg.nms_pre_max_size)
    K = max(1, min(A, n_max if cfg.nms_post_max_size <= 0 else min(n_max, cfg.nms_post_max_size)))
    lid = np.empty((B, K, 7), np.float32)
    cam = np.empty((B, K, 7), np.float64) if rc is not None else None
    sc = np.empty((B, K), np.float32)
    lab = np.empty((B, K), np.int32)
    idx = np.empty((B, K), np.int32)
    cnt = np.empty((B,), np.int32)
    pn = lambda a: None if a is None else _lib.ptr(a)  # noqa: E731
    c = _lib.ctx(device)
    _lib.check(_lib.lib().pp_predict_host(c.handle, C.byref(cfg), _lib.ptr(bp), _lib.ptr(cl), pn(dp), _lib.ptr(an), pn(am),
                                          pn(rc), pn(tv), B, A, K, _lib.ptr(lid), pn(cam), _lib.ptr(sc), _lib.ptr(lab),
                                          _lib.ptr(idx), _lib.ptr(cnt)))
    return lid, cam, sc, lab, idx, cnt


def predict(example, preds_dict, config=None, rotated=False, device=None):
    """VoxelNet.predict(example, preds_dict), model/voxelnet.py:1060-1389."""
    sec = _second(config)
    out = predict_arrays(
        preds_dict["box_preds"], preds_dict["cls_preds"], preds_dict.get("dir_cls_preds"), example[6],
        example[7] if len(example) > 7 else None, example[3], example[4],
        num_class=sec.get("num_class", 1), use_direction_classifier=sec.get("use_direction_classifier", True),
        nms_pre_max_size=sec.get("nms_pre_max_size", 100), nms_post_max_size=sec.get("nms_post_max_size", 50),
        nms_iou_threshold=sec.get("nms_iou_threshold", 0.5), nms_score_threshold=sec.get("nms_score_threshold", 0.0),
        rotated=rotated, device=device)
    lid, cam, sc, lab, _idx, cnt = out
    img_idx = _np(example[8]) if len(example) > 8 else np.arange(lid.shape[0])
    res = []
    for b in range(lid.shape[0]):
        k = int(cnt[b])
        if k == 0:
            res.append({"bbox": None, "box3d_camera": None, "box3d_lidar": None, "scores": None, "label_preds": None,
                        "batch_idx": img_idx[b]})
            continue
        res.append({
            "bbox": np.tile(np.array([400., 200., 500., 400.]), (k, 1)),  # the reference's placeholder, 1357-1360
            "box3d_camera": None if cam is None else cam[b, :k].copy(),
            "box3d_lidar": lid[b, :k].copy(),
            "scores": sc[b, :k].copy(),
            "label_preds": lab[b, :k].astype(np.int64),
            "batch_idx": img_idx[b],
        })
    return res
