"""Device-resident, batched frame pipeline: voxelize(+decorate) -> scatter -> decode -> NMS.

The numpy drop-ins (voxelizer.py, pillars.py, boxes.py, nms.py) move every tensor through the
host, like the reference does.  This module chains the same `*_dev` C-ABI entry points on one
stream with all intermediates left in HBM -- what a DLPack consumer (TensorFlow) would drive --
and copies back only the final detections.  torch is used for device memory,
streams and pinned host buffers only.

Stand-ins: the PFN Dense/BN/ReLU/max (model/pointpillars.py:211-225) and the RPN
(model/voxelnet.py:517-717) stay in the host framework; their outputs enter here as tensors
(`pfn_feats`, `box_enc`, `scores`).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from . import synth as _synth


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class FramePipeline:
    """One GPU, up to `max_frames` frames of at most `max_total_points` points per run."""

    def __init__(self, cfg, device=0, max_frames=64, max_total_points=None, rotated_nms=True,
                 layout="NCHW", fused_decorate=True, keep_voxels=True, anchors=None, overlap_post=True,
                 anchor_area_threshold=None, production=False, sensor_points=None, fused_post=True,
                 max_frame_points=None, scatter_from_cells=True):
        self.cfg = cfg
        self.fused_post = fused_post
        self.dev = torch.device("cuda", device)
        self.B = int(max_frames)
        self.D = cfg["num_point_features"]
        self.P = cfg["max_points"]
        self.MV = cfg["max_voxels"]
        self.C = cfg["num_filters"]
        self.f64 = cfg["point_dtype"] == "float64"
        self.nx, self.ny, self.nz = _synth.grid_size(cfg)
        self.rotated = rotated_nms
        self.layout = layout
        self.fused = fused_decorate
        self.keep_voxels = keep_voxels
        self.vcfg = _lib.make_cfg(cfg["voxel_size"], cfg["point_cloud_range"], self.P, self.MV, True, False)
        self.vx, self.vy = cfg["voxel_size"][:2]
        self.xo = self.vx / 2 + cfg["point_cloud_range"][0]
        self.yo = self.vy / 2 + cfg["point_cloud_range"][1]
        self.pre = cfg["nms_pre_max_size"]
        self.post = cfg["nms_post_max_size"]
        self.thr = cfg["nms_iou_threshold"]
        if max_total_points is None:
            per_frame = 410_000 if not production else ((sensor_points or 848 * 480) + 3) // 4
            max_total_points = self.B * per_frame
            if max_frame_points is None:
                max_frame_points = per_frame
        self.max_pts = int(max_total_points)
        # largest single frame the workspaces are sized for.  Default: the batch capacity (always enough); callers
        # that know their frames are of similar size pass the real bound, which keeps the small-grid workspace small
        self.max_frame_pts = int(max_frame_points) if max_frame_points is not None else self.max_pts
        L = _lib.lib()
        with torch.cuda.device(self.dev):
            an = _synth.anchors_stride(cfg) if anchors is None else np.asarray(anchors, np.float32)
            self.anchors = torch.from_numpy(np.ascontiguousarray(an)).to(self.dev)
            self.A = self.anchors.shape[0]
            B, MV, P, D, Cc = self.B, self.MV, self.P, self.D, self.C
            # The pillar cap: per frame at most min(max_voxels, cells) rows
            self.cap_rows = B * min(MV, self.nx * self.ny * self.nz)
            e = dict(device=self.dev)
            self.voxels = torch.empty((self.cap_rows, P, D), dtype=torch.float32, **e) if keep_voxels else None
            self.decorated = torch.empty((self.cap_rows, P, D + 5), dtype=torch.float32, **e)
            self.coors = torch.empty((self.cap_rows, 4), dtype=torch.int32, **e)
            self.num_points = torch.empty((self.cap_rows,), dtype=torch.int32, **e)
            self.voxel_num = torch.zeros((B,), dtype=torch.int32, **e)
            self.voxel_base = torch.zeros((B + 1,), dtype=torch.int32, **e)
            # the voxelizer's cell -> row map feeds the scatter directly (pp_scatter_cells_dev: no link pass); grids with
            # more than 4 z slabs go through coors (pp_scatter_dev), as do callers that ask for it
            self.cell_voxel = (torch.empty((B * self.nx * self.ny * self.nz,), dtype=torch.int32, **e)
                               if scatter_from_cells and self.nz <= 4 else None)
            self.canvas = torch.empty((B, Cc, self.ny, self.nx) if layout == "NCHW" else (B, self.ny, self.nx, Cc),
                                      dtype=torch.float32, **e)
            # all-anchor decoded / standup tensors exist only on the unfused path (decode -> standup -> NMS -> gather)
            self.boxes = None if fused_post else torch.empty((B, self.A, 7), dtype=torch.float32, **e)
            self.standup = None if (fused_post or rotated_nms) else torch.empty((B, self.A, 4), dtype=torch.float32, **e)
            self.keep = torch.empty((B, self.post), dtype=torch.int32, **e)
            self.keep_count = torch.zeros((B,), dtype=torch.int32, **e)
            self.dets = torch.empty((B, self.post, 8), dtype=torch.float32, **e)
            self.ws_vox_bytes = int(L.pp_voxelize_workspace_bytes(C.byref(self.vcfg), self.max_pts, B, self.max_frame_pts,
                                                                  self.D, _lib.PP_F32))
            self.ws_sc_bytes = int(L.pp_scatter_workspace_bytes(B, self.ny, self.nx, self.cap_rows))
            kind = _lib.PP_NMS_ROTATED if rotated_nms else _lib.PP_NMS_STANDUP
            self.nms_kind = kind
            self.ws_nms_bytes = int(L.pp_nms_workspace_bytes(kind, B, self.A, self.pre))
            self.ws_vox = torch.empty((self.ws_vox_bytes,), dtype=torch.uint8, **e)
            self.ws_sc = torch.empty((self.ws_sc_bytes,), dtype=torch.uint8, **e)
            self.ws_nms = torch.empty((self.ws_nms_bytes,), dtype=torch.uint8, **e)
            # "next" row N1: anchor mask from the voxelizer's coors (load_data.py:3043-3072); the masked
            # scores then feed NMS, which is the device form of `box_preds[a_mask]` (model/voxelnet.py:1119)
            self.area_thr = anchor_area_threshold
            if anchor_area_threshold is not None:
                self.anchor_cells = torch.empty((self.A, 4), dtype=torch.int32, **e)
                self.anchor_mask = torch.empty((B, self.A), dtype=torch.uint8, **e)
                self.masked_scores = torch.empty((B, self.A), dtype=torch.float32, **e)
                self.ws_am_bytes = int(L.pp_anchor_mask_workspace_bytes(B, self.ny, self.nx))
                self.ws_am = torch.empty((self.ws_am_bytes,), dtype=torch.uint8, **e)
                vs3 = (C.c_double * 3)(*map(float, cfg["voxel_size"]))
                pcr6 = (C.c_double * 6)(*map(float, cfg["point_cloud_range"]))
                _lib.check(L.pp_anchor_cells_dev(_p(self.anchors), self.A, vs3, pcr6, _p(self.anchor_cells),
                                                 C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)))
                overlap_post = False  # the post stage now depends on the voxelizer's coors
            # "next" rows N2 / N3: the live production chain (sensor cloud in, camera boxes out)
            self.production = production
            if production:
                from . import ingest as _ingest
                from .predict import make_cfg as _make_predict_cfg
                self.n_sensor = int(sensor_points if sensor_points is not None else 848 * 480)
                self.in_start, self.in_step = 1, 4  # load_data.py:2434
                self.in_cap = max(0, (self.n_sensor - self.in_start + self.in_step - 1) // self.in_step)
                self.in_points = torch.empty((B, self.in_cap, 3), dtype=torch.float64, **e)
                self.in_count = torch.zeros((B,), dtype=torch.int32, **e)
                self.in_off = (torch.arange(B + 1, dtype=torch.int64) * self.in_cap).to(self.dev)
                self.in_rot = np.ascontiguousarray(np.stack([_ingest.R_Y_NEG90, _ingest.R_X_POS90]))
                self.in_lift = np.ascontiguousarray(_ingest.LIFT)
                self.ws_in_bytes = int(L.pp_ingest_workspace_bytes(B, self.n_sensor))
                self.ws_in = torch.empty((self.ws_in_bytes,), dtype=torch.uint8, **e)
                self.pcfg = _make_predict_cfg(1, True, 100, self.pre, self.post, self.thr, 0.0, rotated_nms, False)
                self.ws_pr_bytes = int(L.pp_predict_workspace_bytes(C.byref(self.pcfg), B, self.A, self.post))
                self.ws_pr = torch.empty((self.ws_pr_bytes,), dtype=torch.uint8, **e)
                self.box3d_lidar = torch.empty((B, self.post, 7), dtype=torch.float32, **e)
                self.box3d_camera = torch.empty((B, self.post, 7), dtype=torch.float64, **e)
                self.det_scores = torch.empty((B, self.post), dtype=torch.float32, **e)
                self.det_labels = torch.empty((B, self.post), dtype=torch.int32, **e)
                self.det_index = torch.empty((B, self.post), dtype=torch.int32, **e)
                self.cam_host = torch.empty((B, self.post, 7), dtype=torch.float64).pin_memory()
                self.lidar_host = torch.empty((B, self.post, 7), dtype=torch.float32).pin_memory()
                self.score_host = torch.empty((B, self.post), dtype=torch.float32).pin_memory()
            self.post_stream = torch.cuda.Stream(device=self.dev) if overlap_post else None
            self._ev_fork, self._ev_join = torch.cuda.Event(), torch.cuda.Event()
            self.dets_host = torch.empty((B, self.post, 8), dtype=torch.float32).pin_memory()
            self.keep_count_host = torch.empty((B,), dtype=torch.int32).pin_memory()

    # ---- stages (async on the current torch stream) ------------------------------------------
    def voxelize(self, points, frame_off, n_frames, total_points, max_frame_points, stream):
        L = _lib.lib()
        _lib.check(L.pp_voxelize_dev(
            C.byref(self.vcfg), _p(points), _lib.PP_F64 if points.dtype == torch.float64 else _lib.PP_F32, self.D,
            _p(frame_off), n_frames, total_points, max_frame_points, _lib.PP_F32,
            _p(self.voxels) if self.keep_voxels else None, _p(self.decorated) if self.fused else None,
            _p(self.coors), 4, _p(self.num_points), self.cap_rows, _p(self.voxel_num), _p(self.voxel_base),
            None, _p(self.cell_voxel) if self.cell_voxel is not None else None, _p(self.ws_vox), self.ws_vox_bytes, stream))
        if not self.fused:
            # M is only known on the device: decorate the capacity (rows past M are scratch)
            raise NotImplementedError("unfused decoration needs the host to know M; use fused_decorate=True")

    def scatter(self, pfn_feats, n_frames, stream):
        L = _lib.lib()
        if self.cell_voxel is not None:
            _lib.check(L.pp_scatter_cells_dev(_p(pfn_feats), _p(self.cell_voxel), self.nz, self.C, n_frames, self.ny, self.nx,
                                              _lib.PP_LAYOUT_NCHW if self.layout == "NCHW" else _lib.PP_LAYOUT_NHWC,
                                              _p(self.canvas), stream))
            return
        _lib.check(L.pp_scatter_dev(_p(pfn_feats), _p(self.coors), min(pfn_feats.shape[0], self.cap_rows),
                                    C.c_void_p(self.voxel_base.data_ptr() + 4 * n_frames), self.C, n_frames, self.ny,
                                    self.nx, _lib.PP_LAYOUT_NCHW if self.layout == "NCHW" else _lib.PP_LAYOUT_NHWC,
                                    _p(self.canvas), _p(self.ws_sc), self.ws_sc_bytes, stream))

    def postprocess(self, box_enc, scores, n_frames, stream):
        L = _lib.lib()
        A = self.A
        if self.fused_post:
            # top-k on the scores first, decode only what reaches NMS (the reference's own order, voxelnet.py:1207-1265)
            _lib.check(L.pp_decode_nms_dev(self.nms_kind, _p(box_enc), _p(self.anchors), A, _p(scores), None, n_frames, A,
                                           self.pre, self.post, self.thr, _p(self.keep), self.post, _p(self.keep_count),
                                           _p(self.dets), self.post, _p(self.ws_nms), self.ws_nms_bytes, stream))
            return
        _lib.check(L.pp_box_decode_dev(_p(box_enc), _p(self.anchors), n_frames * A, A, _p(self.boxes), stream))
        if self.rotated:
            _lib.check(L.pp_nms_dev(self.nms_kind, _p(self.boxes), 7, _p(scores), None, n_frames, A, self.pre,
                                    self.post, self.thr, _p(self.keep), self.post, _p(self.keep_count),
                                    _p(self.ws_nms), self.ws_nms_bytes, stream))
        else:
            _lib.check(L.pp_rbox_to_standup_dev(_p(self.boxes), 7, n_frames * A, _p(self.standup), stream))
            _lib.check(L.pp_nms_dev(self.nms_kind, _p(self.standup), 4, _p(scores), None, n_frames, A, self.pre,
                                    self.post, self.thr, _p(self.keep), self.post, _p(self.keep_count),
                                    _p(self.ws_nms), self.ws_nms_bytes, stream))
        _lib.check(L.pp_gather_dets_dev(_p(self.boxes), 7, _p(scores), n_frames, A, _p(self.keep), self.post,
                                        _p(self.keep_count), self.post, _p(self.dets), stream))

    # ---- "next" rows: sensor ingest (N3) and the predict glue (N2) -------------------------------
    def ingest(self, cloud, n_frames, point_step, offsets, stream):
        """cloud: device uint8/float32 tensor holding [n_frames, n_sensor] PointCloud2 records
        (load_data.py:2434-2443).  -> (points [n_frames*cap,3] f64, frame_off, total, cap)"""
        ox, oy, oz = offsets
        _lib.check(_lib.lib().pp_ingest_dev(
            _p(cloud), n_frames, self.n_sensor, point_step, ox, oy, oz, self.in_start, self.in_step,
            _lib.ptr(self.in_rot), 2, _lib.ptr(self.in_lift), _p(self.in_points), self.in_cap, _p(self.in_count),
            _p(self.ws_in), self.ws_in_bytes, stream))
        return self.in_points, self.in_off, n_frames * self.in_cap, self.in_cap

    def predict(self, box_preds, cls_preds, dir_preds, rect, trv2c, anchors_mask, n_frames, stream):
        """VoxelNet.predict's post-network half (model/voxelnet.py:1105-1326) on device tensors."""
        _lib.check(_lib.lib().pp_predict_dev(
            C.byref(self.pcfg), _p(box_preds), _p(cls_preds), _p(dir_preds), _p(self.anchors), _p(anchors_mask),
            _p(rect), _p(trv2c), n_frames, self.A, self.post, _p(self.box3d_lidar), _p(self.box3d_camera),
            _p(self.det_scores), _p(self.det_labels), _p(self.det_index), _p(self.keep_count), _p(self.ws_pr),
            self.ws_pr_bytes, stream))

    def run_production(self, cloud, n_frames, point_step, offsets, pfn_feats, box_preds, cls_preds, dir_preds, rect, trv2c):
        """The reference's live chain for a batch of sensor frames, all on the current stream:
        ingest -> voxelize(+decorate) -> scatter -> anchor mask -> predict (the TF layers in between are
        stand-in tensors)."""
        st = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        pts, off, total, cap = self.ingest(cloud, n_frames, point_step, offsets, st)
        self.voxelize(pts, off, n_frames, total, cap, st)
        self.scatter(pfn_feats, n_frames, st)
        mask = None
        if self.area_thr is not None:
            _lib.check(_lib.lib().pp_anchor_mask_dev(
                _p(self.coors), 4, self.cap_rows, C.c_void_p(self.voxel_base.data_ptr() + 4 * n_frames), n_frames,
                self.ny, self.nx, _p(self.anchor_cells), self.A, float(self.area_thr), None, None,
                _p(self.anchor_mask), None, _p(self.ws_am), self.ws_am_bytes, st))
            mask = self.anchor_mask
        self.predict(box_preds, cls_preds, dir_preds, rect, trv2c, mask, n_frames, st)

    def fetch_production(self, n_frames):
        self.cam_host[:n_frames].copy_(self.box3d_camera[:n_frames], non_blocking=True)
        self.lidar_host[:n_frames].copy_(self.box3d_lidar[:n_frames], non_blocking=True)
        self.score_host[:n_frames].copy_(self.det_scores[:n_frames], non_blocking=True)
        self.keep_count_host[:n_frames].copy_(self.keep_count[:n_frames], non_blocking=True)
        return self.lidar_host, self.cam_host, self.score_host, self.keep_count_host

    # ---- whole path ---------------------------------------------------------------------------
    def run(self, points, frame_off, n_frames, total_points, max_frame_points, pfn_feats, box_enc, scores):
        """All arguments are device tensors (frame_off int64 [n_frames+1]); async on the current stream.

        Pre-processing (voxelize, scatter) and post-processing (decode, NMS) touch disjoint tensors --
        in serving, the post stage of one batch runs while the next batch is being voxelized -- so the
        post stage is issued on a side stream that forks from and joins back into the current stream."""
        main = torch.cuda.current_stream(self.dev)
        st = C.c_void_p(main.cuda_stream)
        if self.post_stream is None:
            self.voxelize(points, frame_off, n_frames, total_points, max_frame_points, st)
            self.scatter(pfn_feats, n_frames, st)
            if self.area_thr is not None:
                _lib.check(_lib.lib().pp_anchor_mask_dev(
                    _p(self.coors), 4, self.cap_rows, C.c_void_p(self.voxel_base.data_ptr() + 4 * n_frames), n_frames,
                    self.ny, self.nx, _p(self.anchor_cells), self.A, float(self.area_thr), _p(scores), None,
                    _p(self.anchor_mask), _p(self.masked_scores), _p(self.ws_am), self.ws_am_bytes, st))
                scores = self.masked_scores
            self.postprocess(box_enc, scores, n_frames, st)
            return
        self._ev_fork.record(main)
        self.post_stream.wait_event(self._ev_fork)
        self.postprocess(box_enc, scores, n_frames, C.c_void_p(self.post_stream.cuda_stream))
        self._ev_join.record(self.post_stream)
        self.voxelize(points, frame_off, n_frames, total_points, max_frame_points, st)
        self.scatter(pfn_feats, n_frames, st)
        main.wait_event(self._ev_join)

    def fetch(self, n_frames):
        """Detections to pinned host memory (async on the current stream)."""
        self.dets_host[:n_frames].copy_(self.dets[:n_frames], non_blocking=True)
        self.keep_count_host[:n_frames].copy_(self.keep_count[:n_frames], non_blocking=True)
        return self.dets_host, self.keep_count_host


class FrameStream:
    """pp_stream (include/pp_b200.h): batches of frames from HOST clouds to HOST detections through one C-ABI call per
    batch; the copy of the next batch overlaps the processing of the current one inside the library.

    The reference's call sites on this boundary: points_to_voxel on the tf.data thread (load_data.py:2966) and the
    numpy detections VoxelNet.predict returns (model/voxelnet.py:1259-1326)."""

    def __init__(self, cfg, device=0, max_frames=64, max_frame_points=410_000, rotated_nms=True, layout="NCHW",
                 keep_voxels=True, anchors=None):
        L = _lib.lib()
        sc = _lib.StreamCfg()
        sc.vox = _lib.make_cfg(cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_points"], cfg["max_voxels"], True, False)
        sc.D = cfg["num_point_features"]
        sc.point_dtype = _lib.PP_F64 if cfg["point_dtype"] == "float64" else _lib.PP_F32
        sc.C = cfg["num_filters"]
        sc.layout = _lib.PP_LAYOUT_NCHW if layout == "NCHW" else _lib.PP_LAYOUT_NHWC
        sc.nms_kind = _lib.PP_NMS_ROTATED if rotated_nms else _lib.PP_NMS_STANDUP
        sc.pre_max, sc.post_max = cfg["nms_pre_max_size"], cfg["nms_post_max_size"]
        sc.iou_threshold = cfg["nms_iou_threshold"]
        sc.max_frames, sc.keep_voxels, sc.max_frame_points = int(max_frames), int(keep_voxels), int(max_frame_points)
        an = np.ascontiguousarray(_synth.anchors_stride(cfg) if anchors is None else anchors, np.float32)
        h = C.c_void_p()
        _lib.check(L.pp_stream_create(int(device), C.byref(sc), _lib.ptr(an), an.shape[0], C.byref(h)))
        self.handle, self.cfg, self.sc = h, cfg, sc
        self.A, self.post = an.shape[0], sc.post_max
        self.cap_rows = int(L.pp_stream_cap_rows(h))
        self.D = sc.D
        self.dtype = np.float64 if sc.point_dtype == _lib.PP_F64 else np.float32

    def bind(self, pfn_feats, box_enc, scores):
        """Device tensors of the host framework (torch tensors or raw pointers): PFN output [cap_rows, C], RPN box
        encodings [max_frames, A, 7], scores [max_frames, A]."""
        self._bound = (pfn_feats, box_enc, scores)  # keep them alive
        _lib.check(_lib.lib().pp_stream_bind(self.handle, _p(pfn_feats), _p(box_enc), _p(scores)))

    def submit(self, points, frame_offsets, dets_out, counts_out):
        """points: numpy [N, D] (pinned: direct DMA); frame_offsets: numpy int64 [n+1]; dets_out [n, post, 8] float32,
        counts_out [n] int32 numpy arrays that are filled when wait() returns.  -> ticket"""
        n = frame_offsets.shape[0] - 1
        assert points.dtype == self.dtype and points.flags.c_contiguous and frame_offsets.dtype == np.int64
        assert dets_out.dtype == np.float32 and dets_out.shape[0] >= n and counts_out.dtype == np.int32
        t = C.c_int64()
        _lib.check(_lib.lib().pp_stream_submit(self.handle, _lib.ptr(points), _lib.ptr(frame_offsets), n, _lib.ptr(dets_out),
                                               _lib.ptr(counts_out), C.byref(t)))
        return int(t.value)

    def wait(self, ticket=-1):
        _lib.check(_lib.lib().pp_stream_wait(self.handle, int(ticket)))

    def view(self):
        v = _lib.StreamTensors()
        _lib.check(_lib.lib().pp_stream_view(self.handle, C.byref(v)))
        return v

    def close(self):
        if self.handle:
            _lib.lib().pp_stream_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def capture_graph(step, device=None, warmup=3):
    """Capture `step()` (a closure that only issues pipeline stages on the current stream: kernel launches,
    memsets, event fork/join -- no host synchronisation) into a CUDA graph and return it; `graph.replay()`
    then re-issues the whole chain with one driver call.  A single frame is 11-17 small launches, so the
    per-launch host cost dominates its latency; the replay removes it.  Inputs and outputs are the static
    device tensors the closure captured (copy new data into them before replay)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):          # warm-up on a side stream, as torch's capture rules require
        for _ in range(warmup):
            step()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize(dev)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    return graph


def shard_frames(n_frames, world_size, rank):
    """Contiguous block sharding of independent frames across ranks: no collective
    on the data path.  Returns (first, count)."""
    base, rem = divmod(n_frames, world_size)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def gather_detections(local_dets, local_counts, world_size):
    """Final detections are gathered on the host (rank 0) -- the only cross-rank step.
    Uses torch.distributed (gloo/NCCL object gather) when initialised, else returns local."""
    import torch.distributed as dist
    if world_size == 1 or not dist.is_initialized():
        return [local_dets], [local_counts]
    dets_list = [None] * world_size if dist.get_rank() == 0 else None
    cnts_list = [None] * world_size if dist.get_rank() == 0 else None
    dist.gather_object(local_dets, dets_list, dst=0)
    dist.gather_object(local_counts, cnts_list, dst=0)
    return dets_list, cnts_list
