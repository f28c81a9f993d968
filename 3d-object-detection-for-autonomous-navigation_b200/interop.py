"""Zero-copy device entry points: DLPack / __cuda_array_interface__ in, DLPack-exportable out.

The numpy drop-ins (voxelizer.py, pillars.py, boxes.py, nms.py) round-trip through the host like the
reference.  A TensorFlow graph keeps its tensors on the GPU, so the same stages are exposed here on
device tensors: anything that implements `__dlpack__` (tf.experimental.dlpack.to_dlpack capsules,
torch, cupy, jax) or `__cuda_array_interface__` (numba) is accepted without a copy, the work is
queued on the current torch stream through the `*_dev` C ABI, and the results are torch tensors
(`torch.utils.dlpack.to_dlpack(t)` / `tf.experimental.dlpack.from_dlpack(...)` hands them on).
torch is only the device-memory and stream provider.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def as_device_tensor(x) -> torch.Tensor:
    """DLPack capsule / __dlpack__ / __cuda_array_interface__ object / torch tensor -> torch CUDA tensor (no copy)."""
    if isinstance(x, torch.Tensor):
        t = x
    elif hasattr(x, "__dlpack__") or type(x).__name__ == "PyCapsule":
        t = torch.utils.dlpack.from_dlpack(x)
    elif hasattr(x, "__cuda_array_interface__"):
        t = torch.as_tensor(x, device="cuda")
    else:
        raise TypeError(f"cannot take a device tensor from {type(x).__name__}")
    if not t.is_cuda:
        raise ValueError("expected a CUDA tensor; host arrays go through the numpy drop-ins")
    return t.contiguous()


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _workspace(device, nbytes):
    """Workspace of one call, from torch's caching allocator: the block is stream-ordered (it is handed out again only
    to work queued behind this call on the same stream), so calls on different streams or threads never share one."""
    return torch.empty((max(int(nbytes), 256),), dtype=torch.uint8, device=device)


def points_to_voxel(points, voxel_size, coors_range, max_points, reverse_index, max_voxels, frame_offsets=None,
                    decorate=False, out_dtype=torch.float32, max_frame_points=None):
    """Device version of load_data.py:695-771 for one frame or a batch (frame_offsets int64 [B+1]).

    Returns a dict of torch tensors: voxels [cap,P,D], coors [cap,4] (frame,c0,c1,c2), num_points [cap],
    voxel_num [B], voxel_base [B+1] (rows of frame b are voxel_base[b]:voxel_base[b+1]) and, when
    `decorate`, decorated [cap,P,D+5] (model/pointpillars.py:143-203 fused).  Nothing synchronises: the number
    of rows stays on the device (voxel_base[-1])."""
    pts = as_device_tensor(points)
    if pts.dtype not in (torch.float32, torch.float64) or pts.dim() != 2:
        raise TypeError("points must be a float32/float64 [N,D] tensor")
    dev = pts.device
    N, D = pts.shape
    if frame_offsets is None:
        off = torch.tensor([0, N], dtype=torch.int64, device=dev)
        B, max_frame = 1, N
    else:
        off = as_device_tensor(frame_offsets).to(torch.int64)
        B = off.numel() - 1
        # the host does not read the offsets back: the caller knows the largest frame, else N bounds it
        max_frame = N if max_frame_points is None else int(max_frame_points)
    cfg = _lib.make_cfg(voxel_size, coors_range, max_points, max_voxels, reverse_index, False)
    L = _lib.lib()
    nx, ny, nz = _lib.grid_size(voxel_size, coors_range)
    cap = B * min(int(max_voxels), nx * ny * nz)
    f64_out = out_dtype == torch.float64
    out = dict(
        voxels=torch.empty((cap, max_points, D), dtype=out_dtype, device=dev),
        coors=torch.empty((cap, 4), dtype=torch.int32, device=dev),
        num_points=torch.empty((cap,), dtype=torch.int32, device=dev),
        voxel_num=torch.empty((B,), dtype=torch.int32, device=dev),
        voxel_base=torch.empty((B + 1,), dtype=torch.int32, device=dev),
    )
    if decorate:
        out["decorated"] = torch.empty((cap, max_points, D + 5), dtype=torch.float32, device=dev)
    ws_bytes = int(L.pp_voxelize_workspace_bytes(C.byref(cfg), N, B, max_frame, D, _lib.PP_F64 if f64_out else _lib.PP_F32))
    ws = _workspace(dev, ws_bytes)
    with torch.cuda.device(dev):
        _lib.check(L.pp_voxelize_dev(
            C.byref(cfg), _p(pts), _lib.PP_F64 if pts.dtype == torch.float64 else _lib.PP_F32, D, _p(off), B, N, max_frame,
            _lib.PP_F64 if f64_out else _lib.PP_F32, _p(out["voxels"]), _p(out.get("decorated")), _p(out["coors"]), 4,
            _p(out["num_points"]), cap, _p(out["voxel_num"]), _p(out["voxel_base"]), None, None, _p(ws), ws_bytes,
            _stream(pts)))
    return out


def pillar_decorate(voxels, num_points, coors, vx, vy, x_offset, y_offset):
    """model/pointpillars.py:143-203 on device tensors: [M,P,D] f32, [M] i32, [M,4] i32 -> [M,P,D+5] f32."""
    v = as_device_tensor(voxels)
    n = as_device_tensor(num_points)
    c = as_device_tensor(coors)
    M, P, D = v.shape
    out = torch.empty((M, P, D + 5), dtype=torch.float32, device=v.device)
    with torch.cuda.device(v.device):
        _lib.check(_lib.lib().pp_decorate_dev(_p(v), _p(n), _p(c), M, P, D, float(vx), float(vy), float(x_offset),
                                              float(y_offset), _p(out), _stream(v)))
    return out


def scatter(voxel_features, coords, batch_size, ny, nx, layout="NCHW", num_rows=None):
    """model/pointpillars.py:285-341 on device tensors.  `num_rows`: optional int32 device scalar (e.g.
    voxel_base[-1:]) when only a prefix of the rows is valid."""
    f = as_device_tensor(voxel_features)
    c = as_device_tensor(coords)
    M, Cc = f.shape
    nhwc = layout == "NHWC"
    out = torch.empty((batch_size, ny, nx, Cc) if nhwc else (batch_size, Cc, ny, nx), dtype=torch.float32, device=f.device)
    L = _lib.lib()
    ws_bytes = int(L.pp_scatter_workspace_bytes(batch_size, ny, nx, M))
    ws = _workspace(f.device, ws_bytes)
    with torch.cuda.device(f.device):
        _lib.check(L.pp_scatter_dev(_p(f), _p(c), M, _p(as_device_tensor(num_rows)) if num_rows is not None else None, Cc,
                                    batch_size, ny, nx, _lib.PP_LAYOUT_NHWC if nhwc else _lib.PP_LAYOUT_NCHW, _p(out), _p(ws),
                                    ws_bytes, _stream(f)))
    return out


def second_box_decode(box_encodings, anchors):
    """libraries/eval_helper_functions.py:388-461 on device tensors; anchors may be one set [A,7] for a batch [B,A,7]."""
    e = as_device_tensor(box_encodings)
    a = as_device_tensor(anchors)
    n = e.numel() // 7
    period = a.numel() // 7 if a.numel() != e.numel() else 0
    out = torch.empty_like(e)
    with torch.cuda.device(e.device):
        _lib.check(_lib.lib().pp_box_decode_dev(_p(e), _p(a), n, period, _p(out), _stream(e)))
    return out


def nms(boxes, scores, pre_max_size=None, post_max_size=None, iou_threshold=0.5, rotated=False):
    """Batched device NMS.  boxes [B,N,4] standup / [B,N,5] rotated / [B,N,7] decoded (rotated only),
    scores [B,N] -> (keep [B,K] int32, keep_count [B] int32), K = min(N, pre, post)."""
    b = as_device_tensor(boxes)
    s = as_device_tensor(scores)
    if b.dim() == 2:
        b, s = b[None], s[None]
    B, N, stride = b.shape
    kind = _lib.PP_NMS_ROTATED if rotated else _lib.PP_NMS_STANDUP
    pre = -1 if pre_max_size is None else int(pre_max_size)
    post = -1 if post_max_size is None else int(post_max_size)
    K = max(1, min(x for x in (N, pre if pre > 0 else N, post if post > 0 else N)))
    keep = torch.empty((B, K), dtype=torch.int32, device=b.device)
    cnt = torch.empty((B,), dtype=torch.int32, device=b.device)
    L = _lib.lib()
    ws_bytes = int(L.pp_nms_workspace_bytes(kind, B, N, pre))
    ws = _workspace(b.device, ws_bytes)
    with torch.cuda.device(b.device):
        _lib.check(L.pp_nms_dev(kind, _p(b), stride, _p(s), None, B, N, pre, post, float(iou_threshold), _p(keep), K, _p(cnt),
                                _p(ws), ws_bytes, _stream(b)))
    return keep, cnt


def rotate_iou(boxes, query_boxes, criterion=-1):
    """nms_gpu.py:526-561 / 618-653 on device tensors: [N,5], [K,5] -> [N,K] float32."""
    b = as_device_tensor(boxes).to(torch.float32)
    q = as_device_tensor(query_boxes).to(torch.float32)
    out = torch.zeros((b.shape[0], q.shape[0]), dtype=torch.float32, device=b.device)
    if b.shape[0] and q.shape[0]:
        with torch.cuda.device(b.device):
            _lib.check(_lib.lib().pp_rotate_iou_dev(_p(b), b.shape[0], _p(q), q.shape[0], int(criterion), _p(out), _stream(b)))
    return out


def to_numpy(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
