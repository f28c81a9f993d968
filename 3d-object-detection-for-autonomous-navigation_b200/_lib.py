"""ctypes binding of libpp_b200.so (include/pp_b200.h).  No CPU fallback: if the library is not
built, or no CUDA device is usable, every compute call raises."""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

from . import build as _build

PP_F32, PP_F64 = 0, 1
PP_LAYOUT_NCHW, PP_LAYOUT_NHWC = 0, 1
PP_NMS_STANDUP, PP_NMS_ROTATED = 0, 1


class PPError(RuntimeError):
    pass


class VoxelCfg(C.Structure):
    _fields_ = [("voxel_size", C.c_double * 3), ("coors_range", C.c_double * 6),
                ("max_points", C.c_int32), ("max_voxels", C.c_int32),
                ("reverse_index", C.c_int32), ("arith_f32", C.c_int32)]


class PredictCfg(C.Structure):
    _fields_ = [("num_class", C.c_int32), ("use_direction_classifier", C.c_int32), ("top_k", C.c_int32),
                ("nms_pre_max_size", C.c_int32), ("nms_post_max_size", C.c_int32), ("nms_kind", C.c_int32),
                ("nms_iou_threshold", C.c_float), ("nms_score_threshold", C.c_float),
                ("anchors_per_frame", C.c_int32)]


class StreamCfg(C.Structure):
    _fields_ = [("vox", VoxelCfg), ("D", C.c_int32), ("point_dtype", C.c_int32), ("C", C.c_int32), ("layout", C.c_int32),
                ("nms_kind", C.c_int32), ("pre_max", C.c_int32), ("post_max", C.c_int32), ("iou_threshold", C.c_float),
                ("max_frames", C.c_int32), ("keep_voxels", C.c_int32), ("max_frame_points", C.c_int64)]


class StreamTensors(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("voxels", "decorated", "coors", "num_points", "voxel_num", "voxel_base", "canvas",
                                          "dets", "keep_count", "compute_stream")]


_vp, _i32, _i64, _f32, _f64, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double, C.c_size_t
_cfgp = C.POINTER(VoxelCfg)
_pcfgp = C.POINTER(PredictCfg)

# name -> (restype, argtypes); must list every symbol include/pp_b200.h declares
SIGNATURES = {
    "pp_last_error_string": (C.c_char_p, []),
    "pp_version": (C.c_int, []),
    "pp_launch_count": (_i64, [C.c_int]),
    "pp_profile_start": (C.c_int, []),
    "pp_profile_stop": (C.c_int, [C.c_char_p, _sz, C.POINTER(_f32), C.c_int]),
    "pp_grid_size": (C.c_int, [C.POINTER(_f64), C.POINTER(_f64), C.c_int, C.POINTER(_i32)]),
    "pp_voxelize_workspace_bytes": (_sz, [_cfgp, _i64, C.c_int, _i64, C.c_int, C.c_int]),
    "pp_voxelize_dev": (C.c_int, [_cfgp, _vp, C.c_int, C.c_int, _vp, C.c_int, _i64, _i64, C.c_int, _vp, _vp,
                                  _vp, C.c_int, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "pp_voxelize_set_small_path_min_points": (C.c_int, [_i64]),
    "pp_decorate_dev": (C.c_int, [_vp, _vp, _vp, _i64, C.c_int, C.c_int, _f64, _f64, _f64, _f64, _vp, _vp]),
    "pp_scatter_workspace_bytes": (_sz, [C.c_int, C.c_int, C.c_int, _i64]),
    "pp_scatter_dev": (C.c_int, [_vp, _vp, _i64, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _sz, _vp]),
    "pp_scatter_cells_dev": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "pp_box_decode_dev": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _vp]),
    "pp_rbox_to_standup_dev": (C.c_int, [_vp, C.c_int, _i64, _vp, _vp]),
    "pp_nms_workspace_bytes": (_sz, [C.c_int, C.c_int, _i64, C.c_int]),
    "pp_nms_dev": (C.c_int, [C.c_int, _vp, C.c_int, _vp, _vp, C.c_int, _i64, C.c_int, C.c_int, _f32, _vp, _i64,
                             _vp, _vp, _sz, _vp]),
    "pp_decode_nms_dev": (C.c_int, [C.c_int, _vp, _vp, _i64, _vp, _vp, C.c_int, _i64, C.c_int, C.c_int, _f32, _vp, _i64, _vp,
                                    _vp, C.c_int, _vp, _sz, _vp]),
    "pp_gather_dets_dev": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _i64, _vp, _i64, _vp, C.c_int, _vp, _vp]),
    "pp_rotate_iou_dev": (C.c_int, [_vp, _i64, _vp, _i64, C.c_int, _vp, _vp]),
    "pp_d3_box_overlap_dev": (C.c_int, [_vp, _i64, _vp, _i64, C.c_int, _vp, _vp]),
    "pp_anchor_cells_dev": (C.c_int, [_vp, _i64, C.POINTER(_f64), C.POINTER(_f64), _vp, _vp]),
    "pp_anchor_mask_workspace_bytes": (_sz, [C.c_int, C.c_int, C.c_int]),
    "pp_anchor_mask_dev": (C.c_int, [_vp, C.c_int, _i64, _vp, C.c_int, C.c_int, C.c_int, _vp, _i64, _f32, _vp, _vp, _vp,
                                     _vp, _vp, _sz, _vp]),
    "pp_predict_workspace_bytes": (_sz, [_pcfgp, C.c_int, _i64, C.c_int]),
    "pp_predict_dev": (C.c_int, [_pcfgp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, _i64, C.c_int, _vp, _vp, _vp, _vp,
                                 _vp, _vp, _vp, _sz, _vp]),
    "pp_predict_host": (C.c_int, [_vp, _pcfgp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, _i64, C.c_int, _vp, _vp, _vp,
                                  _vp, _vp, _vp]),
    "pp_ingest_workspace_bytes": (_sz, [C.c_int, _i64]),
    "pp_ingest_dev": (C.c_int, [_vp, C.c_int, _i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_int, _vp,
                                _vp, _i64, _vp, _vp, _sz, _vp]),
    "pp_ingest_host": (C.c_int, [_vp, _vp, _i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_int, _vp,
                                 _vp, _i64, C.POINTER(_i32)]),
    "pp_ctx_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "pp_ctx_destroy": (None, [_vp]),
    "pp_ctx_stream": (_vp, [_vp]),
    "pp_ctx_device": (C.c_int, [_vp]),
    "pp_ctx_sync": (C.c_int, [_vp]),
    "pp_host_alloc": (C.c_int, [_sz, C.POINTER(_vp)]),
    "pp_host_free": (None, [_vp]),
    "pp_points_to_voxel_host": (C.c_int, [_vp, _cfgp, _vp, C.c_int, _i64, C.c_int, _vp, _vp, _vp, C.POINTER(_i32), _vp]),
    "pp_decorate_host": (C.c_int, [_vp, _vp, _vp, _vp, _i64, C.c_int, C.c_int, _f64, _f64, _f64, _f64, _vp]),
    "pp_scatter_host": (C.c_int, [_vp, _vp, _vp, _i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "pp_box_decode_host": (C.c_int, [_vp, _vp, _vp, _i64, _vp]),
    "pp_rbox_to_standup_host": (C.c_int, [_vp, _vp, _i64, _vp]),
    "pp_nms_host": (C.c_int, [_vp, C.c_int, _vp, _vp, _i64, C.c_int, C.c_int, _f32, _vp, C.POINTER(_i32)]),
    "pp_anchors_mask_host": (C.c_int, [_vp, _vp, _i64, _vp, _i64, C.POINTER(_f64), C.POINTER(_f64), _f32, _vp, _vp]),
    "pp_d3_box_overlap_host": (C.c_int, [_vp, _vp, _i64, _vp, _i64, C.c_int, _vp]),
    "pp_rotate_iou_host": (C.c_int, [_vp, _vp, _i64, _vp, _i64, C.c_int, _vp]),
    "pp_stream_create": (C.c_int, [C.c_int, C.POINTER(StreamCfg), _vp, _i64, C.POINTER(_vp)]),
    "pp_stream_destroy": (None, [_vp]),
    "pp_stream_cap_rows": (_i64, [_vp]),
    "pp_stream_anchor_count": (_i64, [_vp]),
    "pp_stream_bind": (C.c_int, [_vp, _vp, _vp, _vp]),
    "pp_stream_submit": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp, _vp, C.POINTER(_i64)]),
    "pp_stream_wait": (C.c_int, [_vp, _i64]),
    "pp_stream_view": (C.c_int, [_vp, C.POINTER(StreamTensors)]),
}

_lib = None
_lock = threading.Lock()


def lib():
    """The loaded shared library (built on first use when sources are newer)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                path = _build.LIB
                if _build.needs_build():
                    try:
                        _build.build()
                    except Exception as e:  # noqa: BLE001
                        if not os.path.exists(path):
                            raise PPError(f"libpp_b200.so is not built and nvcc failed: {e}") from e
                l = C.CDLL(path)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(l, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = l
    return _lib


def check(rc: int):
    if rc != 0:
        msg = lib().pp_last_error_string().decode(errors="replace")
        raise PPError(f"libpp_b200 error {rc}: {msg}")


def launch_count(reset=False) -> int:
    return int(lib().pp_launch_count(int(reset)))


def profile_start():
    check(lib().pp_profile_start())


def profile_stop(max_records=65536):
    """-> list of (kernel name, device ms) in launch order since profile_start()."""
    names = C.create_string_buffer(max_records * 24)
    ms = (_f32 * max_records)()
    n = lib().pp_profile_stop(names, len(names), ms, max_records)
    labels = names.value.decode().split("\n")[:n]
    return list(zip(labels, [float(ms[i]) for i in range(n)]))


class Ctx:
    """pp_ctx wrapper: stream + device arena for the host-buffer entry points."""

    def __init__(self, device: int = 0):
        h = _vp()
        check(lib().pp_ctx_create(int(device), C.byref(h)))
        self.handle = h
        self.device = int(device)

    @property
    def stream(self) -> int:
        return int(lib().pp_ctx_stream(self.handle) or 0)

    def sync(self):
        check(lib().pp_ctx_sync(self.handle))

    def close(self):
        if self.handle:
            lib().pp_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


_tls = threading.local()
_default_device = 0


def set_device(device: int):
    """Device used by the numpy drop-in functions on this thread (reference: device_id=0)."""
    _tls.device = int(device)


def ctx(device=None) -> Ctx:
    """One context per (thread, device): the reference calls the voxelizer on the tf.data thread
    and NMS on the main thread concurrently."""
    if device is None:
        device = getattr(_tls, "device", _default_device)
    cache = getattr(_tls, "ctxs", None)
    if cache is None:
        cache = _tls.ctxs = {}
    c = cache.get(device)
    if c is None:
        c = cache[device] = Ctx(device)
    return c


class _PinnedBlock:
    """Owner of one pp_host_alloc block (freed when the last numpy view dies)."""

    def __init__(self, nbytes):
        h = _vp()
        check(lib().pp_host_alloc(int(nbytes), C.byref(h)))
        self.ptr, self.nbytes = h.value, int(nbytes)

    def __del__(self):
        try:
            if self.ptr:
                lib().pp_host_free(_vp(self.ptr))
                self.ptr = None
        except Exception:  # noqa: BLE001
            pass


def pinned_empty(shape, dtype=np.float32) -> np.ndarray:
    """numpy array in page-locked host memory: the *_host entry points DMA straight from / into it."""
    dt = np.dtype(dtype)
    shape = tuple(int(x) for x in (shape if isinstance(shape, (tuple, list)) else (shape,)))
    n = int(np.prod(shape)) * dt.itemsize
    blk = _PinnedBlock(max(n, 1))
    blk.__array_interface__ = {"shape": shape, "typestr": dt.str, "data": (blk.ptr, False), "version": 3}
    return np.asarray(blk)   # the array's base keeps the block alive


def make_cfg(voxel_size, coors_range, max_points, max_voxels, reverse_index, arith_f32) -> VoxelCfg:
    cfg = VoxelCfg()
    for i in range(3):
        cfg.voxel_size[i] = float(voxel_size[i])
    for i in range(6):
        cfg.coors_range[i] = float(coors_range[i])
    cfg.max_points = int(max_points)
    cfg.max_voxels = int(max_voxels)
    cfg.reverse_index = int(bool(reverse_index))
    cfg.arith_f32 = int(bool(arith_f32))
    return cfg


def grid_size(voxel_size, coors_range, arith_f32=False):
    g = (_i32 * 3)()
    check(lib().pp_grid_size((_f64 * 3)(*map(float, voxel_size)), (_f64 * 6)(*map(float, coors_range)),
                             int(arith_f32), g))
    return [int(v) for v in g]


def ptr(a: np.ndarray):
    return a.ctypes.data_as(_vp)
