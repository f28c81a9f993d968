"""GPU: BASELINE.json's full-size configurations through size-independent properties (the CPU
oracle cannot finish them in seconds), plus oracle checks on sampled pieces.

  configs[2]  KITTI-shaped 120k-point clouds, 432x496 BEV, 12k-pillar cap, batch 64
  configs[3]  rotated BEV NMS, 100k boxes per frame, IoU 0.5
  configs[4]  streaming batch of D435 frames (batch-vs-single-frame equivalence, determinism)
"""
import ctypes as C
import importlib

import numpy as np
import pytest

from conftest import PKG

pytestmark = pytest.mark.gpu


def _voxelize_batch(cfg, frames, decorated=False):
    import torch
    pipeline = importlib.import_module(PKG + ".pipeline")
    _lib = importlib.import_module(PKG + "._lib")
    B = len(frames)
    pts = np.concatenate(frames)
    off = np.cumsum([0] + [f.shape[0] for f in frames]).astype(np.int64)
    pipe = pipeline.FramePipeline(cfg, device=0, max_frames=B, max_total_points=pts.shape[0])
    dev = torch.device("cuda", 0)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    pipe.voxelize(torch.from_numpy(pts).to(dev), torch.from_numpy(off).to(dev), B, pts.shape[0],
                  max(f.shape[0] for f in frames), st)
    torch.cuda.synchronize()
    vb = pipe.voxel_base[:B + 1].cpu().numpy()
    M = int(vb[B])
    return (pipe.voxels[:M].cpu().numpy(), pipe.coors[:M].cpu().numpy(), pipe.num_points[:M].cpu().numpy(), vb,
            pipe.decorated[:M].cpu().numpy() if decorated else None, pipe)


def test_kitti_batch64_properties(pp, oracle, synth):
    cfg = synth.KITTI
    vs, pcr = np.array(cfg["voxel_size"]), np.array(cfg["point_cloud_range"])
    frames = [synth.kitti_cloud(100 + i, shuffled=bool(i & 1)) for i in range(64)]
    vox, coors, num, vb, _, pipe = _voxelize_batch(cfg, frames)
    nx, ny, nz = synth.grid_size(cfg)
    assert np.all(np.diff(vb) <= cfg["max_voxels"]) and np.all(np.diff(vb) > 0)
    # coordinates inside the grid, unique per frame, pillars never empty, padding is zero
    assert coors[:, 1].max() < nz and coors[:, 2].max() < ny and coors[:, 3].max() < nx and coors.min() >= 0
    key = ((coors[:, 0].astype(np.int64) * nz + coors[:, 1]) * ny + coors[:, 2]) * nx + coors[:, 3]
    assert np.unique(key).size == key.size
    assert num.min() >= 1 and num.max() <= cfg["max_points"]
    pad = np.arange(cfg["max_points"])[None, :] >= num[:, None]
    assert not vox[pad].any()
    # every stored point lies in the cell its pillar claims (float64 arithmetic of the reference)
    rows = np.repeat(np.arange(vox.shape[0]), cfg["max_points"]).reshape(vox.shape[0], -1)[~pad]
    p3 = vox[~pad][:, :3].astype(np.float64)
    cell = np.floor((p3 - pcr[:3]) / vs).astype(np.int64)
    assert np.array_equal(cell[:, 0], coors[rows, 3]) and np.array_equal(cell[:, 1], coors[rows, 2])
    # sampled frames against the oracle, bit-exact (ring-major and shuffled, cap binding)
    for b in (0, 1, 31, 62, 63):
        ov, oc, on = oracle.points_to_voxel(frames[b], vs, pcr, cfg["max_points"], True, cfg["max_voxels"])
        lo, hi = vb[b], vb[b + 1]
        assert np.array_equal(coors[lo:hi, 1:], oc) and np.array_equal(num[lo:hi], on) and np.array_equal(vox[lo:hi], ov)
    # scatter of the whole batch: each occupied cell holds its pillar's feature row, the rest is zero
    import torch
    feats = synth.pfn_standin(vox.shape[0], cfg["num_filters"], 7)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    pipe.scatter(torch.from_numpy(feats).cuda(), 64, st)
    torch.cuda.synchronize()
    canvas = pipe.canvas.cpu().numpy()  # [64, C, ny, nx]
    assert np.array_equal(canvas[coors[:, 0], :, coors[:, 2], coors[:, 3]], feats)
    assert np.count_nonzero(canvas.any(axis=1)) == vox.shape[0]
    np.testing.assert_allclose(canvas.sum(dtype=np.float64), feats.sum(dtype=np.float64), rtol=1e-9)


def test_d435_stream_batch_equals_single_and_is_deterministic(pp, synth):
    cfg = synth.D435
    frames = [synth.d435_cloud(200 + i) for i in range(12)]
    a = _voxelize_batch(cfg, frames, decorated=True)
    b = _voxelize_batch(cfg, frames, decorated=True)
    for x, y in zip(a[:5], b[:5]):
        assert np.array_equal(x, y)  # bit-identical run to run (atomics only order bucket internals)
    vb = a[3]
    for i in (0, 5, 11):
        s = _voxelize_batch(cfg, [frames[i]], decorated=True)
        lo, hi = vb[i], vb[i + 1]
        assert np.array_equal(a[0][lo:hi], s[0]) and np.array_equal(a[1][lo:hi, 1:], s[1][:, 1:])
        assert np.array_equal(a[2][lo:hi], s[2]) and np.array_equal(a[4][lo:hi], s[4])
    # point order matters (first come): reversing a frame changes pillar order but not the pillar set
    r = _voxelize_batch(cfg, [frames[0][::-1].copy()])
    k0 = {tuple(c) for c in a[1][vb[0]:vb[1], 1:]}
    assert {tuple(c) for c in r[1][:, 1:]} == k0


@pytest.mark.parametrize("clustered", [False, True])
def test_rotated_nms_100k(pp, oracle, synth, clustered):
    n = 100_000
    d = synth.rotated_boxes(n, 77, clustered=clustered)
    keep = np.asarray(pp.rotate_nms_gpu(d, 0.5))
    assert keep.size > 0 and np.unique(keep).size == keep.size
    # keep order is descending score; the best box is always kept
    assert np.all(np.diff(d[keep, 5]) < 0) and keep[0] == int(np.argmax(d[:, 5]))
    kb = d[keep]
    rng = np.random.default_rng(0)
    # (1) kept boxes do not suppress each other: sampled rows of the kept x kept IoU matrix
    rows = rng.choice(keep.size, size=min(300, keep.size), replace=False)
    iou = pp.rotate_iou_gpu(kb[:, :5], kb[rows, :5])  # [n_keep, rows]
    iou[rows, np.arange(rows.size)] = 0
    assert iou.max() <= 0.5 + 1e-5
    # (2) every suppressed box (sampled) overlaps a kept box with a higher score
    kept_mask = np.zeros(n, bool); kept_mask[keep] = True
    sup = rng.choice(np.nonzero(~kept_mask)[0], size=300, replace=False) if (~kept_mask).any() else np.array([], int)
    if sup.size:
        iou = pp.rotate_iou_gpu(kb[:, :5], d[sup, :5])  # [n_keep, 300]
        higher = kb[:, 5][:, None] > d[sup, 5][None, :]
        assert np.all((np.where(higher, iou, 0).max(axis=0) > 0.5 - 1e-5))
    # (3) idempotence: NMS of the kept set keeps everything
    again = pp.rotate_nms_gpu(kb, 0.5)
    assert again == list(range(keep.size))
    # (4) the first 3000 boxes by score form a prefix-closed problem: oracle on them == our prefix
    top = oracle.argsort_desc(d[:, 5])[:3000]
    want = [int(top[i]) for i in oracle.rotate_nms_gpu(d[top], 0.5)]
    got = [int(k) for k in keep if d[k, 5] >= d[top[-1], 5]]
    assert got == want


def test_stripe_nms_equals_all_pairs_nms(pp, oracle, synth):
    """Above 16 384 boxes pp_nms_dev switches to the stripe-sequential algorithm; it must return exactly
    what the all-pairs bitmask algorithm returns (here: 20 000 boxes vs the same boxes split so that the
    all-pairs path handles them, and vs the CPU oracle on a prefix-closed subset)."""
    n = 20_000
    for clustered in (False, True):
        d = synth.rotated_boxes(n, 91, clustered=clustered)
        got = pp.rotate_nms_gpu(d, 0.5)                                  # stripes
        top = oracle.argsort_desc(d[:, 5])
        want_all = pp.rotate_nms_gpu(d[top[:16000]], 0.5)                # all-pairs path on the 16 000 best
        prefix = [int(top[i]) for i in want_all]
        assert got[:len(prefix)] == prefix
        want = [int(top[i]) for i in oracle.rotate_nms_gpu(d[top[:4000]], 0.5)]
        assert got[:len(want)] == want
        # caps: pre_max keeps the stripe path (17 000 > 16 384), post_max stops early
        g2 = pp.rotate_nms_gpu(d, 0.5, pre_max_size=17000, post_max_size=123)
        assert g2 == got[:123]
        # standup kind through the same path
        sb = oracle.rbox_to_standup(d[:, :5]) * np.float32(10)
        k_all = pp.nms(sb[top[:16000]], d[top[:16000], 5], None, None, 0.5)
        k_str = pp.nms(sb, d[:, 5], None, None, 0.5)
        assert k_str[:len(k_all)].tolist() == [int(top[i]) for i in k_all]


@pytest.mark.parametrize("thr", [0.05, 0.3, 0.7])
def test_stripe_nms_thresholds(pp, oracle, synth, thr):
    """The large-N path skips polygon clips whose intersection-area upper bound cannot reach thr/(1+thr)*(a1+a2);
    the keep list must not depend on that (all-pairs path on the 16 000 best boxes, CPU oracle on the 3 000 best),
    for loose and tight thresholds and for boxes of very different sizes."""
    n = 24_000
    rng = np.random.default_rng(int(thr * 100))
    d = synth.rotated_boxes(n, 300 + int(thr * 100), clustered=True)
    d[:, 2] *= rng.uniform(0.3, 3.0, n).astype(np.float32)     # widths and lengths over an order of magnitude
    d[:, 3] *= rng.uniform(0.3, 3.0, n).astype(np.float32)
    got = pp.rotate_nms_gpu(d, thr)
    top = oracle.argsort_desc(d[:, 5])
    prefix = [int(top[i]) for i in pp.rotate_nms_gpu(d[top[:16000]], thr)]
    assert got[:len(prefix)] == prefix
    want = [int(top[i]) for i in oracle.rotate_nms_gpu(d[top[:3000]], thr)]
    assert got[:len(want)] == want


@pytest.mark.parametrize("rotated", [True, False])
def test_nms_stress_batched(pp, synth, rotated):
    """BASELINE configs[3] is a BATCH of 100k-box frames (nms_gpu.py:455-490 once per frame in the reference): eight
    distinct frames in ONE pp_nms_dev call -- per-frame cursors, kept counts and bins of the large-N path -- must give
    exactly the eight single-frame results, with and without post_max, for the rotated and the standup kind."""
    import torch
    _lib = importlib.import_module(PKG + "._lib")
    oracle = importlib.import_module("oracle")
    B, N = 8, 100_000
    frames = [synth.rotated_boxes(N, 700 + i, clustered=bool(i & 1)) for i in range(B)]
    if rotated:
        boxes = np.stack([f[:, :5] for f in frames])
    else:
        boxes = np.stack([oracle.rbox_to_standup(f[:, :5]) * np.float32(10) for f in frames])
    scores = np.stack([f[:, 5] for f in frames])
    kind = _lib.PP_NMS_ROTATED if rotated else _lib.PP_NMS_STANDUP
    dev = torch.device("cuda", 0)
    L = _lib.lib()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def run(bx, sc, post):
        b, n = sc.shape
        K = n if post <= 0 else post
        tb, ts = torch.from_numpy(np.ascontiguousarray(bx)).to(dev), torch.from_numpy(np.ascontiguousarray(sc)).to(dev)
        keep = torch.full((b, K), -1, dtype=torch.int32, device=dev)
        cnt = torch.zeros((b,), dtype=torch.int32, device=dev)
        wsb = int(L.pp_nms_workspace_bytes(kind, b, n, -1))
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        _lib.check(L.pp_nms_dev(kind, C.c_void_p(tb.data_ptr()), bx.shape[2], C.c_void_p(ts.data_ptr()), None, b, n, -1, post, 0.5,
                                C.c_void_p(keep.data_ptr()), K, C.c_void_p(cnt.data_ptr()), C.c_void_p(ws.data_ptr()), wsb, st))
        torch.cuda.synchronize()
        c = cnt.cpu().numpy()
        k = keep.cpu().numpy()
        return [k[i, :c[i]].tolist() for i in range(b)]
    for post in (-1, 777):
        batched = run(boxes, scores, post)
        for i in range(B):
            single = run(boxes[i:i + 1], scores[i:i + 1], post)[0]
            assert len(single) > 100 and batched[i] == single, f"frame {i}, post_max {post}"
        if post > 0:
            assert all(len(k) == post for k in batched)
