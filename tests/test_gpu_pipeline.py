"""GPU: the batched, device-resident pipeline (packed pillars, fused decoration, scatter from device M,
decode, NMS, detection gather) against the per-frame CPU oracle."""
import importlib

import numpy as np
import pytest

from conftest import PKG

pytestmark = pytest.mark.gpu


def _run(synth, oracle, cfg, frames, rotated, layout, scatter_from_cells=True, max_voxels=None):
    import torch
    pipeline = importlib.import_module(PKG + ".pipeline")
    B = len(frames)
    D = cfg["num_point_features"]
    tdtype = torch.float64 if cfg["point_dtype"] == "float64" else torch.float32
    pts = np.concatenate(frames)
    off = np.cumsum([0] + [f.shape[0] for f in frames]).astype(np.int64)
    if max_voxels is not None:
        cfg = dict(cfg, max_voxels=max_voxels)
    pipe = pipeline.FramePipeline(cfg, device=0, max_frames=B, max_total_points=pts.shape[0], rotated_nms=rotated,
                                  layout=layout, scatter_from_cells=scatter_from_cells)
    assert (pipe.cell_voxel is not None) == scatter_from_cells
    A = pipe.A
    box = np.stack([synth.rpn_standin(A, 50 + i)[0] for i in range(B)])
    sco = np.stack([synth.rpn_standin(A, 50 + i)[1] for i in range(B)])
    feats = synth.pfn_standin(pipe.cap_rows, cfg["num_filters"], 3)
    dev = torch.device("cuda", 0)
    d_pts = torch.from_numpy(pts).to(dev)
    assert d_pts.dtype == tdtype
    pipe.run(d_pts, torch.from_numpy(off).to(dev), B, pts.shape[0], max(f.shape[0] for f in frames),
             torch.from_numpy(feats).to(dev), torch.from_numpy(box).to(dev), torch.from_numpy(sco).to(dev))
    dets_h, cnt_h = pipe.fetch(B)
    torch.cuda.synchronize()
    vbase = pipe.voxel_base[:B + 1].cpu().numpy()
    vnum = pipe.voxel_num[:B].cpu().numpy()
    M = int(vbase[B])
    vox = pipe.voxels[:M].cpu().numpy(); dec = pipe.decorated[:M].cpu().numpy()
    coors = pipe.coors[:M].cpu().numpy(); num = pipe.num_points[:M].cpu().numpy()
    canvas = pipe.canvas[:B].cpu().numpy()
    vs, pcr = np.array(cfg["voxel_size"]), np.array(cfg["point_cloud_range"])
    vx, vy = cfg["voxel_size"][:2]
    xo, yo = vx / 2 + pcr[0], vy / 2 + pcr[1]
    nx, ny, _ = synth.grid_size(cfg)
    all_c4 = []
    for b, f in enumerate(frames):
        ov, oc, on = oracle.points_to_voxel(f, vs, pcr, cfg["max_points"], True, cfg["max_voxels"])
        lo, hi = int(vbase[b]), int(vbase[b + 1])
        assert hi - lo == ov.shape[0] == int(vnum[b])
        assert np.array_equal(coors[lo:hi, 0], np.full(hi - lo, b)) and np.array_equal(coors[lo:hi, 1:], oc)
        assert np.array_equal(num[lo:hi], on)
        assert np.array_equal(vox[lo:hi], ov.astype(np.float32))
        c4 = np.concatenate([np.full((oc.shape[0], 1), b, np.int32), oc], axis=1)
        all_c4.append(c4)
        want = oracle.decorate(ov.astype(np.float32), on, c4, vx, vy, xo, yo)
        np.testing.assert_allclose(dec[lo:hi], want, rtol=1e-5, atol=1e-5 * (float(np.abs(ov).max()) + 1 if ov.size else 1.0))
        # decode + NMS
        boxes = oracle.second_box_decode(box[b], synth.anchors_stride(cfg))
        if rotated:
            dets = np.concatenate([boxes[:, [0, 1, 3, 4, 6]], sco[b][:, None]], axis=1)
            keep = oracle.rotate_nms_gpu(dets, cfg["nms_iou_threshold"], cfg["nms_pre_max_size"], cfg["nms_post_max_size"])
        else:
            sb = oracle.rbox_to_standup(boxes[:, [0, 1, 3, 4, 6]])
            k = oracle.nms(sb, sco[b], cfg["nms_pre_max_size"], cfg["nms_post_max_size"], cfg["nms_iou_threshold"])
            keep = [] if k is None else k.tolist()
        assert int(cnt_h[b]) == len(keep)
        got = dets_h[b, :len(keep)].numpy()
        np.testing.assert_allclose(got[:, :7], boxes[keep], rtol=1e-5, atol=1e-6)
        assert np.array_equal(got[:, 7], sco[b][keep])
        assert not dets_h[b, len(keep):].numpy().any()
    want_canvas = oracle.scatter(feats[:M], np.concatenate(all_c4), B, ny, nx, layout)
    assert np.array_equal(canvas, want_canvas)


def test_pipeline_d435_batched_ragged(synth, oracle):
    cfg = synth.D435
    frames = [synth.d435_cloud(20), synth.d435_cloud(21, subsample=True), synth.d435_cloud(22)[:1000],
              np.zeros((0, 3), np.float64), synth.d435_cloud(23)[::3]]
    _run(synth, oracle, cfg, frames, rotated=True, layout="NCHW")
    _run(synth, oracle, cfg, frames[:2], rotated=False, layout="NHWC")


def test_pipeline_scatter_paths(synth, oracle, vox_path):
    """The canvas from the voxelizer's cell -> row map (pp_scatter_cells_dev) and from coors (pp_scatter_dev), on both
    voxelizer implementations; with the max_voxels cap binding, cells past the cap must stay empty on the canvas."""
    frames = [synth.d435_cloud(24), synth.d435_cloud(25)[::5], np.zeros((0, 3), np.float64), synth.d435_cloud(26)[:300000]]
    for cells in (True, False):
        _run(synth, oracle, synth.D435, frames, rotated=True, layout="NCHW", scatter_from_cells=cells)
    _run(synth, oracle, synth.D435, frames, rotated=True, layout="NHWC", scatter_from_cells=True, max_voxels=700)
    _run(synth, oracle, synth.D435, frames[:1], rotated=True, layout="NCHW", scatter_from_cells=True)
    kitti = [synth.kitti_cloud(33), synth.uniform_cloud(30000, synth.KITTI, 34)]
    _run(synth, oracle, dict(synth.KITTI), kitti, rotated=False, layout="NCHW", scatter_from_cells=False)


def test_pipeline_kitti_batched(synth, oracle):
    cfg = dict(synth.KITTI)
    frames = [synth.kitti_cloud(30), synth.kitti_cloud(31, shuffled=True), synth.uniform_cloud(50000, cfg, 32)]
    _run(synth, oracle, cfg, frames, rotated=True, layout="NCHW")


def test_pipeline_with_anchor_mask(synth, oracle):
    """N1 wired into the batch pipeline: masked-out anchors never reach NMS (model/voxelnet.py:1119-1137)."""
    import torch
    pipeline = importlib.import_module(PKG + ".pipeline")
    cfg = synth.D435
    frames = [synth.d435_cloud(40)[:60000], synth.d435_cloud(41, subsample=True), synth.d435_cloud(42)[200000:230000]]
    B = len(frames)
    pts = np.concatenate(frames)
    off = np.cumsum([0] + [f.shape[0] for f in frames]).astype(np.int64)
    pipe = pipeline.FramePipeline(cfg, device=0, max_frames=B, max_total_points=pts.shape[0], anchor_area_threshold=1)
    A = pipe.A
    an = synth.anchors_stride(cfg)
    box = np.stack([synth.rpn_standin(A, 70 + i)[0] for i in range(B)])
    sco = np.stack([synth.rpn_standin(A, 70 + i)[1] for i in range(B)])
    feats = synth.pfn_standin(pipe.cap_rows, cfg["num_filters"], 3)
    dev = torch.device("cuda", 0)
    pipe.run(torch.from_numpy(pts).to(dev), torch.from_numpy(off).to(dev), B, pts.shape[0], max(f.shape[0] for f in frames),
             torch.from_numpy(feats).to(dev), torch.from_numpy(box).to(dev), torch.from_numpy(sco).to(dev))
    dets_h, cnt_h = pipe.fetch(B)
    torch.cuda.synchronize()
    vs, pcr = np.array(cfg["voxel_size"]), np.array(cfg["point_cloud_range"])
    gm = pipe.anchor_mask.cpu().numpy().astype(bool)
    for b, f in enumerate(frames):
        _, oc, _ = oracle.points_to_voxel(f, vs, pcr, cfg["max_points"], True, cfg["max_voxels"])
        _, mask = oracle.anchors_mask(oc, an, vs, pcr, 1)
        assert np.array_equal(gm[b], mask) and 0 < mask.sum() < A
        idx = np.nonzero(mask)[0]
        boxes = oracle.second_box_decode(box[b], an)
        dets = np.concatenate([boxes[:, [0, 1, 3, 4, 6]], sco[b][:, None]], axis=1)[idx]
        keep = [int(idx[k]) for k in oracle.rotate_nms_gpu(dets, cfg["nms_iou_threshold"], cfg["nms_pre_max_size"],
                                                           cfg["nms_post_max_size"])]
        assert int(cnt_h[b]) == len(keep)
        got = dets_h[b, :len(keep)].numpy()
        np.testing.assert_allclose(got[:, :7], boxes[keep], rtol=1e-5, atol=1e-6)
        assert np.array_equal(got[:, 7], sco[b][keep])


def test_profiler_and_launch_count(pp, synth):
    _lib = importlib.import_module(PKG + "._lib")
    pts = synth.d435_cloud(1, subsample=True)
    vs, pcr = np.array(synth.D435["voxel_size"]), np.array(synth.D435["point_cloud_range"])
    # the d435i grid (10 240 cells) has the shared-memory table path: four launches, no memset.  Batches of fewer than
    # 150 000 points (this 101 760-point frame) take the any-grid path unless the threshold is lowered.
    small = ["vox_scan", "vox_prefix", "vox_place", "vox_finish"]
    anygrid = ["vox_mark", "vox_cell", "vox_rank", "vox_rowmap", "vox_bucket", "vox_gather"]
    for thr, kernels, memset in ((0, small, []), (-1, anygrid, ["vox_memset"])):
        _lib.check(_lib.lib().pp_voxelize_set_small_path_min_points(thr))
        pp.launch_count(reset=True)
        _lib.profile_start()
        pp.points_to_voxel(pts, vs, pcr, 50, True, 12000)
        rec = _lib.profile_stop()
        assert [n for n, _ in rec] == memset + kernels
        assert all(t >= 0 for _, t in rec)
        assert pp.launch_count() == len(kernels)
    # a KITTI-sized grid (214 272 cells) takes the any-grid path
    kp = synth.kitti_cloud(0)
    pp.launch_count(reset=True)
    _lib.profile_start()
    pp.points_to_voxel(kp, np.array(synth.KITTI["voxel_size"]), np.array(synth.KITTI["point_cloud_range"]), 100, True, 12000)
    names = [n for n, _ in _lib.profile_stop()]
    kernels = ["vox_mark", "vox_cell", "vox_rank", "vox_rowmap", "vox_bucket", "vox_gather"]
    assert names == ["vox_memset"] + kernels
    assert pp.launch_count() == len(kernels)


@pytest.mark.parametrize("rotated", [False, True])
def test_production_chain(synth, oracle, rotated):
    """N3 + a1-a5 + N1 + N2 chained on the device for a batch of sensor frames (the reference's live loop:
    load_data.py:2434-2443 -> 2966 -> 3043-3072 -> model/voxelnet.py:1060-1389) against the per-frame oracle."""
    import torch
    pipeline = importlib.import_module(PKG + ".pipeline")
    ing = importlib.import_module(PKG + ".ingest")
    cfg = synth.D435
    n_sensor = 848 * 480
    clouds = [synth.d435_sensor_cloud(60), synth.d435_sensor_cloud(61, invalid=0.5), synth.d435_sensor_cloud(62, invalid=0.0)]
    clouds[1][:200000] = np.nan          # a frame whose valid pixels are all in the lower half of the image
    B = len(clouds)
    pipe = pipeline.FramePipeline(cfg, device=0, max_frames=B, rotated_nms=rotated, anchor_area_threshold=1,
                                  production=True, sensor_points=n_sensor)
    A = pipe.A
    an = synth.anchors_stride(cfg)
    rng = np.random.default_rng(9)
    bp = rng.normal(0, 0.1, (B, A, 7)).astype(np.float32)
    cl = rng.normal(-2, 1, (B, A, 1)).astype(np.float32)
    dr = rng.normal(0, 1, (B, A, 2)).astype(np.float32)
    rect = np.tile(np.eye(4, dtype=np.float32), (B, 1, 1))
    trv = np.tile(np.array([[0, -1, 0, 0.01], [0, 0, -1, -0.07], [1, 0, 0, -0.27], [0, 0, 0, 1]], np.float32), (B, 1, 1))
    feats = synth.pfn_standin(pipe.cap_rows, cfg["num_filters"], 3)
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    pipe.run_production(t(np.stack(clouds)), B, 12, (0, 4, 8), t(feats), t(bp), t(cl), t(dr), t(rect), t(trv))
    lid_h, cam_h, sc_h, cnt_h = pipe.fetch_production(B)
    torch.cuda.synchronize()
    vbase = pipe.voxel_base[:B + 1].cpu().numpy()
    coors = pipe.coors[:int(vbase[B])].cpu().numpy()
    num = pipe.num_points[:int(vbase[B])].cpu().numpy()
    n_in = pipe.in_count[:B].cpu().numpy()
    idx = pipe.det_index[:B].cpu().numpy()
    gm = pipe.anchor_mask[:B].cpu().numpy()
    vs, pcr = np.array(cfg["voxel_size"]), np.array(cfg["point_cloud_range"])
    for b in range(B):
        pts = oracle.pointcloud2_to_lidar(clouds[b], (ing.R_Y_NEG90, ing.R_X_POS90), ing.LIFT, 1, 4)
        assert n_in[b] == pts.shape[0]
        got_pts = pipe.in_points[b].cpu().numpy()
        assert np.array_equal(got_pts[:n_in[b]], pts) and np.isnan(got_pts[n_in[b]:]).all()
        _, oc, on = oracle.points_to_voxel(pts, vs, pcr, cfg["max_points"], True, cfg["max_voxels"])
        lo, hi = int(vbase[b]), int(vbase[b + 1])
        assert np.array_equal(coors[lo:hi, 1:], oc) and np.array_equal(num[lo:hi], on)
        _, mask = oracle.anchors_mask(oc, an, vs, pcr, 1)
        assert np.array_equal(gm[b].astype(bool), mask)
        want = oracle.predict_frame(bp[b], cl[b], dr[b], an, mask.astype(np.uint8), rect[b], trv[b], rotated=rotated)
        k = int(cnt_h[b])
        assert want["box3d_lidar"] is not None and k == want["box3d_lidar"].shape[0]
        assert np.array_equal(idx[b, :k], want["anchor_index"])
        np.testing.assert_allclose(lid_h[b, :k].numpy(), want["box3d_lidar"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(cam_h[b, :k].numpy(), want["box3d_camera"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(sc_h[b, :k].numpy(), want["scores"], rtol=1e-6, atol=0)


def test_cuda_graph_replay_matches_direct_launches(synth):
    """The whole chain is stream-ordered kernel launches and memsets: it must be capturable into a CUDA graph and the
    replay must reproduce the direct run bit for bit, also after the inputs changed in place."""
    import torch
    pipeline = importlib.import_module(PKG + ".pipeline")
    cfg = synth.D435
    dev = torch.device("cuda", 0)
    f0, f1 = synth.d435_cloud(80, subsample=True), synth.d435_cloud(81, subsample=True)
    n = f0.shape[0]
    pipe = pipeline.FramePipeline(cfg, device=0, max_frames=1, max_total_points=n)
    A = pipe.A
    pts = torch.from_numpy(f0).to(dev)
    off = torch.tensor([0, n], dtype=torch.int64, device=dev)
    box = torch.from_numpy(synth.rpn_standin(A, 5)[0][None]).to(dev)
    sco = torch.from_numpy(synth.rpn_standin(A, 5)[1][None]).to(dev)
    feats = torch.from_numpy(synth.pfn_standin(pipe.cap_rows, cfg["num_filters"], 3)).to(dev)
    step = lambda: pipe.run(pts, off, 1, n, n, feats, box, sco)  # noqa: E731

    def snapshot():
        torch.cuda.synchronize()
        m = int(pipe.voxel_base[1].item())
        return [pipe.coors[:m].clone(), pipe.num_points[:m].clone(), pipe.decorated[:m].clone(), pipe.canvas.clone(),
                pipe.dets.clone(), pipe.keep_count.clone()]
    graph = pipeline.capture_graph(step, dev)
    for frame in (f0, f1, f0):
        pts.copy_(torch.from_numpy(frame).to(dev))
        step()
        want = snapshot()
        for t in (pipe.coors, pipe.num_points, pipe.decorated, pipe.canvas, pipe.dets, pipe.keep_count):
            t.zero_()
        graph.replay()
        got = snapshot()
        assert want[0].shape[0] > 1000 and int(want[5][0]) > 0
        for a, b in zip(want, got):
            assert torch.equal(a, b)


@pytest.mark.parametrize("name,rotated", [("d435", True), ("d435", False), ("kitti", True), ("kitti", False)])
def test_decode_nms_fused_equals_decode_then_nms(synth, name, rotated):
    """pp_decode_nms_dev (top-k, decode of the selected boxes only, NMS, decoded detections) must return bit for bit
    what box_decode on every anchor -> (standup) -> pp_nms_dev -> gather returns."""
    import torch
    pipeline = importlib.import_module(PKG + ".pipeline")
    cfg = synth.D435 if name == "d435" else synth.KITTI
    B = 3
    dev = torch.device("cuda", 0)
    outs = []
    for fused in (True, False):
        pipe = pipeline.FramePipeline(cfg, device=0, max_frames=B, max_total_points=1000, rotated_nms=rotated, fused_post=fused)
        A = pipe.A
        box = torch.from_numpy(np.stack([synth.rpn_standin(A, 90 + i)[0] for i in range(B)])).to(dev)
        sco = np.stack([synth.rpn_standin(A, 90 + i)[1] for i in range(B)])
        sco[2, ::3] = -np.inf     # absent anchors
        sco = torch.from_numpy(sco).to(dev)
        pipe.postprocess(box, sco, B, C_void(torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        outs.append((pipe.keep_count[:B].cpu().numpy().copy(), pipe.keep[:B].cpu().numpy().copy(), pipe.dets[:B].cpu().numpy().copy()))
    (c0, k0, d0), (c1, k1, d1) = outs
    assert np.array_equal(c0, c1) and c0.min() > 0
    for b in range(B):
        assert np.array_equal(k0[b, :c0[b]], k1[b, :c1[b]])
    assert np.array_equal(d0, d1)


@pytest.mark.parametrize("rotated", [True, False])
def test_small_nms_every_cluster_size(synth, rotated):
    """nms_small runs a frame on a thread-block cluster of 8 / 4 / 2 / 1 CTAs depending on the batch size (B <= 9 / 37 / 74 /
    above on 148 SMs): the same frames must give the same keep lists and detections at every size, frame by frame
    (batch 1 is the configuration the per-frame oracle tests pin)."""
    import torch
    pipeline = importlib.import_module(PKG + ".pipeline")
    cfg = synth.D435
    dev = torch.device("cuda", 0)
    Bmax = 80
    an = synth.anchors_stride(cfg)
    A = an.shape[0]
    box_h = np.stack([synth.rpn_standin(A, 300 + (i % 7))[0] for i in range(Bmax)])
    sco_h = np.stack([synth.rpn_standin(A, 300 + (i % 7))[1] for i in range(Bmax)])
    sco_h[5, ::2] = -np.inf      # a frame with absent anchors
    sco_h[6, :] = -np.inf        # a frame with none at all
    sco_h[7, 100:] = -np.inf     # fewer present boxes than pre_max_size
    box, sco = torch.from_numpy(box_h).to(dev), torch.from_numpy(sco_h).to(dev)
    out = {}
    for B in (1, 9, 10, 37, 38, 74, 80):
        pipe = pipeline.FramePipeline(cfg, device=0, max_frames=B, max_total_points=1000, rotated_nms=rotated)
        pipe.postprocess(box[:B].contiguous(), sco[:B].contiguous(), B, C_void(torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        out[B] = (pipe.keep_count[:B].cpu().numpy().copy(), pipe.keep[:B].cpu().numpy().copy(), pipe.dets[:B].cpu().numpy().copy())
    cnt80, keep80, dets80 = out[80]   # one CTA per frame
    assert cnt80[6] == 0 and not dets80[6].any() and 0 < cnt80[7] <= 50 and cnt80[0] > 0
    for B, (cnt, keep, dets) in out.items():
        assert np.array_equal(cnt, cnt80[:B]) and np.array_equal(dets, dets80[:B]), B
        for f in range(B):
            assert np.array_equal(keep[f, :cnt[f]], keep80[f, :cnt80[f]]), (B, f)
    # frames repeat with period 7 (apart from the three edited ones): equal inputs, equal outputs
    for f in range(8, 80):
        a = f % 7
        if a in (5, 6):
            continue
        assert cnt80[f] == cnt80[a] and np.array_equal(keep80[f, :cnt80[f]], keep80[a, :cnt80[a]]) and np.array_equal(dets80[f], dets80[a])

def C_void(x):
    import ctypes
    return ctypes.c_void_p(x)


def test_production_chain_empty_and_tiny_frames(synth, oracle):
    """Edge frames through the whole device chain: a frame whose every pixel is invalid (no points, no pillars, no
    anchors in the mask, no detections -- the reference's None branch) next to a normal one and one with a handful of points."""
    import torch
    pipeline = importlib.import_module(PKG + ".pipeline")
    ing = importlib.import_module(PKG + ".ingest")
    cfg = synth.D435
    n_sensor = 848 * 480
    dead = np.full((n_sensor, 3), np.nan, np.float32)
    tiny = dead.copy()
    tiny[1000:1040] = synth.d435_sensor_cloud(70, invalid=0.0)[200000:200040]
    clouds = [dead, synth.d435_sensor_cloud(71), tiny]
    B = len(clouds)
    pipe = pipeline.FramePipeline(cfg, device=0, max_frames=B, rotated_nms=False, anchor_area_threshold=1,
                                  production=True, sensor_points=n_sensor)
    A = pipe.A
    an = synth.anchors_stride(cfg)
    rng = np.random.default_rng(19)
    bp = rng.normal(0, 0.1, (B, A, 7)).astype(np.float32)
    cl = rng.normal(-2, 1, (B, A, 1)).astype(np.float32)
    dr = rng.normal(0, 1, (B, A, 2)).astype(np.float32)
    eye = np.tile(np.eye(4, dtype=np.float32), (B, 1, 1))
    feats = synth.pfn_standin(pipe.cap_rows, cfg["num_filters"], 3)
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    pipe.run_production(t(np.stack(clouds)), B, 12, (0, 4, 8), t(feats), t(bp), t(cl), t(dr), t(eye), t(eye))
    lid_h, cam_h, sc_h, cnt_h = pipe.fetch_production(B)
    torch.cuda.synchronize()
    vnum = pipe.voxel_num[:B].cpu().numpy()
    n_in = pipe.in_count[:B].cpu().numpy()
    assert n_in[0] == 0 and vnum[0] == 0 and int(cnt_h[0]) == 0 and not pipe.anchor_mask[0].any().item()
    vs, pcr = np.array(cfg["voxel_size"]), np.array(cfg["point_cloud_range"])
    for b in (1, 2):
        pts = oracle.pointcloud2_to_lidar(clouds[b], (ing.R_Y_NEG90, ing.R_X_POS90), ing.LIFT, 1, 4)
        assert n_in[b] == pts.shape[0]
        _, oc, _ = oracle.points_to_voxel(pts, vs, pcr, cfg["max_points"], True, cfg["max_voxels"])
        assert vnum[b] == oc.shape[0]
        _, mask = oracle.anchors_mask(oc, an, vs, pcr, 1)
        want = oracle.predict_frame(bp[b], cl[b], dr[b], an, mask.astype(np.uint8), eye[b], eye[b])
        k = int(cnt_h[b])
        if want["box3d_lidar"] is None:
            assert k == 0
        else:
            assert k == want["box3d_lidar"].shape[0]
            assert np.array_equal(pipe.det_index[b, :k].cpu().numpy(), want["anchor_index"])
    assert n_in[2] == 10 and 0 < vnum[2] <= 10


def test_host_stream_matches_oracle(pp, synth, oracle):
    """pp_stream (the host-buffer batch entry point): host clouds in, host detections out through one C-ABI call per
    batch, pinned and pageable memory, three batches so that both staging slots are reused.  Reference boundary:
    load_data.py:2966 (clouds in) / model/voxelnet.py:1259-1326 (detections out)."""
    import torch
    pipeline = importlib.import_module(PKG + ".pipeline")
    _lib = importlib.import_module(PKG + "._lib")
    cfg = synth.D435
    B = 3
    fs = pipeline.FrameStream(cfg, device=0, max_frames=B, max_frame_points=110_000)
    A, post = fs.A, fs.post
    dev = torch.device("cuda", 0)
    box = np.stack([synth.rpn_standin(A, 80 + i)[0] for i in range(B)])
    sco = np.stack([synth.rpn_standin(A, 80 + i)[1] for i in range(B)])
    feats = synth.pfn_standin(fs.cap_rows, cfg["num_filters"], 3)
    t_feats, t_box, t_sco = (torch.from_numpy(a).to(dev) for a in (feats, box, sco))
    fs.bind(t_feats, t_box, t_sco)
    batches = [[synth.d435_cloud(60 + 3 * j + i, subsample=True)[: 101_760 - 7000 * i] for i in range(B)] for j in range(3)]
    batches[1][1] = np.zeros((0, 3), np.float64)  # an empty frame
    batches[2] = batches[2][:2]                   # a short batch
    outs = []
    for j, frames in enumerate(batches):
        n = len(frames)
        pts = np.concatenate(frames)
        off = np.cumsum([0] + [f.shape[0] for f in frames]).astype(np.int64)
        if j != 1:  # pinned clouds and results: direct DMA
            hp = _lib.pinned_empty(pts.shape, np.float64); hp[...] = pts
            dets, cnt = _lib.pinned_empty((n, post, 8), np.float32), _lib.pinned_empty((n,), np.int32)
        else:       # pageable numpy arrays: staged inside the library
            hp, dets, cnt = pts.copy(), np.empty((n, post, 8), np.float32), np.empty((n,), np.int32)
        outs.append((fs.submit(hp, off, dets, cnt), hp, dets, cnt, frames))
    fs.wait()
    vs, pcr = np.array(cfg["voxel_size"]), np.array(cfg["point_cloud_range"])
    an = synth.anchors_stride(cfg)
    for _, _, dets, cnt, frames in outs:
        for b in range(len(frames)):
            boxes = oracle.second_box_decode(box[b], an)
            d = np.concatenate([boxes[:, [0, 1, 3, 4, 6]], sco[b][:, None]], axis=1)
            keep = oracle.rotate_nms_gpu(d, cfg["nms_iou_threshold"], cfg["nms_pre_max_size"], cfg["nms_post_max_size"])
            assert int(cnt[b]) == len(keep)
            np.testing.assert_allclose(dets[b, :len(keep), :7], boxes[keep], rtol=1e-5, atol=1e-6)
            assert np.array_equal(dets[b, :len(keep), 7], sco[b][keep]) and not dets[b, len(keep):].any()
    # the device tensors of the last batch are the voxelizer's / scatter's results for that batch
    v = fs.view()
    frames = batches[2]
    n = len(frames)
    torch.cuda.synchronize()

    class _Raw:  # a raw device pointer as a __cuda_array_interface__ object
        def __init__(self, ptr, shape, typestr):
            self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (int(ptr), False), "version": 3}
    vbase = torch.as_tensor(_Raw(v.voxel_base, (n + 1,), "<i4"), device="cuda").cpu().numpy()
    M = int(vbase[n])
    coors = torch.as_tensor(_Raw(v.coors, (M, 4), "<i4"), device="cuda").cpu().numpy()
    num = torch.as_tensor(_Raw(v.num_points, (M,), "<i4"), device="cuda").cpu().numpy()
    for b, f in enumerate(frames):
        _, oc, on = oracle.points_to_voxel(f, vs, pcr, cfg["max_points"], True, cfg["max_voxels"])
        lo, hi = int(vbase[b]), int(vbase[b + 1])
        assert hi - lo == oc.shape[0] and np.array_equal(coors[lo:hi, 1:], oc) and np.array_equal(num[lo:hi], on)
    with pytest.raises(_lib.PPError):
        fs.submit(np.zeros((200_000, 3)), np.array([0, 200_000], np.int64), np.empty((1, post, 8), np.float32), np.empty(1, np.int32))
    fs.close()


def test_two_threads_voxelizer_and_nms(pp, synth, oracle):
    """SURVEY 8(b) concurrency contract: the reference runs points_to_voxel on the tf.data generator thread
    (load_data.py:2339, 2389-2392, 2966) while the main thread is inside nms (model/voxelnet.py:1259).  Two python
    threads, each with its own per-thread context, 200 calls each, must reproduce the serial results bit for bit."""
    import threading
    cfg = synth.D435
    vs, pcr = np.array(cfg["voxel_size"]), np.array(cfg["point_cloud_range"])
    clouds = [synth.d435_cloud(90 + i, subsample=True)[: 30_000 + 5000 * i] for i in range(4)]
    an = synth.anchors_stride(cfg)
    dets = []
    for i in range(4):
        be, sc = synth.rpn_standin(an.shape[0], 90 + i)
        boxes = pp.second_box_decode(be, an)
        dets.append((pp.rbox_to_standup(boxes[:, [0, 1, 3, 4, 6]]), sc))
    want_v = [pp.points_to_voxel(c, vs, pcr, cfg["max_points"], True, cfg["max_voxels"]) for c in clouds]
    want_k = [pp.nms(b, s, cfg["nms_pre_max_size"], cfg["nms_post_max_size"], cfg["nms_iou_threshold"]) for b, s in dets]
    for (v, c, n), cl in zip(want_v, clouds):  # the serial results are the oracle's
        ov, oc, on = oracle.points_to_voxel(cl, vs, pcr, cfg["max_points"], True, cfg["max_voxels"])
        assert np.array_equal(v, ov) and np.array_equal(c, oc) and np.array_equal(n, on)
    errors = []

    def voxel_loop():
        try:
            for it in range(200):
                i = it % 4
                v, c, n = pp.points_to_voxel(clouds[i], vs, pcr, cfg["max_points"], True, cfg["max_voxels"])
                assert np.array_equal(v, want_v[i][0]) and np.array_equal(c, want_v[i][1]) and np.array_equal(n, want_v[i][2])
        except BaseException as e:  # noqa: BLE001
            errors.append(("voxelizer thread", e))

    def nms_loop():
        try:
            for it in range(200):
                i = it % 4
                k = pp.nms(dets[i][0], dets[i][1], cfg["nms_pre_max_size"], cfg["nms_post_max_size"], cfg["nms_iou_threshold"])
                assert (k is None and want_k[i] is None) or np.array_equal(k, want_k[i])
        except BaseException as e:  # noqa: BLE001
            errors.append(("nms thread", e))

    ts = [threading.Thread(target=voxel_loop), threading.Thread(target=nms_loop)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors
