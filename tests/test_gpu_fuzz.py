"""GPU: randomised configurations of the voxelizer / scatter / NMS against the CPU oracle (small sizes)."""
import numpy as np
import pytest

from nms_check import assert_keep_lists_agree

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(24))
def test_voxelizer_random_config(pp, oracle, seed, vox_path):
    rng = np.random.default_rng(1000 + seed)
    D = int(rng.integers(3, 7))
    dt = np.float64 if rng.random() < 0.5 else np.float32
    grid = rng.integers(1, 40, size=3)
    vs = rng.uniform(0.05, 2.0, size=3)
    lo = rng.uniform(-20, 5, size=3)
    # ranges that are exact multiples, half-cell remainders (np.round half-to-even) or arbitrary
    extra = rng.choice([0.0, 0.5, 0.37])
    hi = lo + (grid + extra) * vs
    pcr = np.concatenate([lo, hi])
    N = int(rng.choice([0, 1, 31, 32, 33, 1000, 5000, 20000]))
    P = int(rng.choice([1, 2, 5, 31, 32, 33, 64, 65, 100, 129, 254]))
    cap = int(rng.choice([0, 1, 7, 100, 5000]))
    # clustered points so that some cells overflow max_points and the cap binds
    centers = rng.uniform(lo - 0.2 * (hi - lo), hi + 0.2 * (hi - lo), size=(max(1, N // 50), 3))
    pts = centers[rng.integers(0, centers.shape[0], N)] + rng.normal(0, 1.5, size=(N, 3)) * vs
    if D > 3:
        pts = np.concatenate([pts, rng.random((N, D - 3))], axis=1)
    pts = np.ascontiguousarray(pts.astype(dt))
    rev = bool(rng.random() < 0.5)
    params = [(vs, pcr)]
    if dt == np.float32:
        params.append((vs.astype(np.float32), pcr.astype(np.float32)))  # float32 arithmetic
        params.append((vs.tolist(), pcr.tolist()))                      # lists -> cast to points.dtype
    for v_, r_ in params:
        got = pp.points_to_voxel(pts, v_, r_, P, rev, cap, return_point_slots=True)
        want = oracle.points_to_voxel(pts, v_, r_, P, rev, cap, return_slots=True)
        for a, b in zip(got, want):
            assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b)


@pytest.mark.parametrize("seed", range(8))
def test_scatter_and_nms_random(pp, oracle, synth, seed):
    rng = np.random.default_rng(2000 + seed)
    B, ny, nx, C = int(rng.integers(1, 5)), int(rng.integers(1, 70)), int(rng.integers(1, 90)), int(rng.choice([1, 3, 4, 17, 64, 129]))
    M = int(rng.choice([0, 1, 50, 3000]))
    coords = np.stack([rng.integers(0, B, M), rng.integers(0, 3, M), rng.integers(0, ny, M), rng.integers(0, nx, M)], axis=1).astype(np.int32)
    feats = rng.normal(size=(M, C)).astype(np.float32)
    for layout in ("NCHW", "NHWC"):
        assert np.array_equal(pp.scatter(feats, coords, B, ny, nx, layout), oracle.scatter(feats, coords, B, ny, nx, layout))
    n = int(rng.choice([1, 2, 63, 64, 65, 128, 129, 700, 1500]))
    d = synth.rotated_boxes(n, 3000 + seed, clustered=True)
    thr = float(rng.choice([0.05, 0.3, 0.5, 0.7]))
    pre = rng.choice([None, 10, 100, 128, 129, 1024, 1100])
    post = rng.choice([None, 1, 50, 300])
    pre = None if pre is None else int(pre)
    post = None if post is None else int(post)
    got = pp.rotate_nms_gpu(d, thr, pre_max_size=pre, post_max_size=post)
    want = oracle.rotate_nms_gpu(d, thr, pre, post)
    assert_keep_lists_agree(got, want, d, thr, oracle, tol=1e-6, pre_max_size=pre, post_max_size=post)
    sb = oracle.rbox_to_standup(d[:, :5]) * np.float32(rng.choice([1.0, 20.0]))
    g2 = pp.nms(sb, d[:, 5], pre, post, thr)
    w2 = oracle.nms(sb, d[:, 5], pre, post, thr)
    assert (g2 is None and w2 is None) or g2.tolist() == w2.tolist()


@pytest.mark.parametrize("seed", range(4))
def test_predict_fuzz(pp, oracle, synth, seed):
    """N2: random shapes / options of the predict glue against the oracle (anchor indices bit-exact)."""
    rng = np.random.default_rng(1000 + seed)
    A = int(rng.integers(1, 3000))
    B = int(rng.integers(1, 5))
    nc = int(rng.integers(1, 4))
    an = synth.anchors_stride(synth.D435)[rng.choice(10240, A, replace=A > 10240)]
    bp = rng.normal(0, 0.2, (B, A, 7)).astype(np.float32)
    cl = rng.normal(-1, 1.5, (B, A, nc)).astype(np.float32)
    dr = rng.normal(0, 1, (B, A, 2)).astype(np.float32)
    mask = (rng.random((B, A)) < rng.uniform(0.05, 1.0)).astype(np.uint8) if rng.random() < 0.7 else None
    rect = rng.normal(0, 1, (B, 4, 4)).astype(np.float32)
    trv = rng.normal(0, 1, (B, 4, 4)).astype(np.float32)
    opts = dict(top_k=int(rng.integers(1, 129)), nms_pre_max_size=int(rng.integers(1, 200)), nms_post_max_size=int(rng.integers(1, 80)),
                nms_iou_threshold=float(rng.uniform(0.05, 0.8)), nms_score_threshold=float(rng.choice([0.0, 0.1, 0.3])),
                rotated=bool(rng.integers(0, 2)))
    lid, cam, sc, lab, idx, cnt = pp.predict_arrays(bp, cl, dr, an, mask, rect, trv, num_class=nc, **opts)
    for b in range(B):
        want = oracle.predict_frame(bp[b], cl[b], dr[b], an, None if mask is None else mask[b], rect[b], trv[b],
                                    top_k=opts["top_k"], pre_max_size=opts["nms_pre_max_size"], post_max_size=opts["nms_post_max_size"],
                                    iou_threshold=opts["nms_iou_threshold"], score_threshold=opts["nms_score_threshold"],
                                    rotated=opts["rotated"])
        k = int(cnt[b])
        if want["box3d_lidar"] is None:
            assert k == 0
            continue
        assert np.array_equal(idx[b, :k], want["anchor_index"]), (seed, b, opts)
        assert np.array_equal(lab[b, :k], want["label_preds"])
        np.testing.assert_allclose(lid[b, :k], want["box3d_lidar"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(cam[b, :k], want["box3d_camera"], rtol=1e-5, atol=1e-4)
        np.testing.assert_allclose(sc[b, :k], want["scores"], rtol=1e-6, atol=0)


@pytest.mark.parametrize("seed", range(6))
def test_predict_large_selection(pp, oracle, synth, seed):
    """N2 with min(top_k, nms_pre_max_size) > 128 (KITTI-style 1000 / 300): the general decode + NMS path behind the same call."""
    rng = np.random.default_rng(7000 + seed)
    A = int(rng.choice([150, 3000, 10240, 40000]))
    B = int(rng.integers(1, 4))
    nc = int(rng.integers(1, 4))
    base = synth.anchors_stride(synth.D435)
    an = base[rng.choice(10240, A, replace=A > 10240)]
    per_frame = bool(seed % 3 == 2)
    if per_frame:
        an = np.stack([an[rng.permutation(A)] for _ in range(B)])
    bp = rng.normal(0, 0.2, (B, A, 7)).astype(np.float32)
    cl = rng.normal(-1, 1.5, (B, A, nc)).astype(np.float32)
    dr = rng.normal(0, 1, (B, A, 2)).astype(np.float32)
    mask = (rng.random((B, A)) < rng.uniform(0.05, 1.0)).astype(np.uint8) if rng.random() < 0.6 else None
    rect = rng.normal(0, 1, (B, 4, 4)).astype(np.float32)
    trv = rng.normal(0, 1, (B, 4, 4)).astype(np.float32)
    top_k = int(rng.choice([129, 500, 1000, 1025, 3000]))
    pre = int(rng.choice([-1, 129, 1000, 2000]))
    post = int(rng.choice([-1, 50, 300]))
    opts = dict(top_k=top_k, nms_pre_max_size=pre, nms_post_max_size=post, nms_iou_threshold=float(rng.uniform(0.05, 0.8)),
                nms_score_threshold=float(rng.choice([0.0, 0.05, 0.3])), rotated=bool(seed % 2))
    lid, cam, sc, lab, idx, cnt = pp.predict_arrays(bp, cl, dr, an, mask, rect, trv, num_class=nc, **opts)
    for b in range(B):
        want = oracle.predict_frame(bp[b], cl[b], dr[b], an[b] if per_frame else an, None if mask is None else mask[b], rect[b],
                                    trv[b], top_k=top_k, pre_max_size=pre, post_max_size=post,
                                    iou_threshold=opts["nms_iou_threshold"], score_threshold=opts["nms_score_threshold"],
                                    rotated=opts["rotated"])
        k = int(cnt[b])
        if want["box3d_lidar"] is None:
            assert k == 0
            continue
        assert np.array_equal(idx[b, :k], want["anchor_index"]), (seed, b, opts)
        assert np.array_equal(lab[b, :k], want["label_preds"])
        np.testing.assert_allclose(lid[b, :k], want["box3d_lidar"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(cam[b, :k], want["box3d_camera"], rtol=1e-5, atol=1e-4)
        np.testing.assert_allclose(sc[b, :k], want["scores"], rtol=1e-6, atol=0)
        assert np.all(idx[b, k:] == -1) and not lid[b, k:].any()


@pytest.mark.parametrize("seed", range(3))
def test_long_score_lists_small_selection(pp, oracle, synth, seed):
    """>= 16 384 scores per frame with <= 128 boxes into NMS (KITTI heads with the reference's top-100): the selection comes
    from the cluster top-k, the rest from the one-launch NMS / the predict kernel; ties and absent anchors included."""
    rng = np.random.default_rng(8000 + seed)
    n = int(rng.choice([16384, 30000, 107136]))
    d = synth.rotated_boxes(n, 8100 + seed, clustered=bool(seed & 1))
    d[:, 5] = np.round(d[:, 5], 2)            # many tied scores: the tie rule (descending index) decides the selection
    thr = float(rng.choice([0.1, 0.5]))
    pre, post = int(rng.choice([100, 128])), int(rng.choice([50, 128]))
    got = pp.rotate_nms_gpu(d, thr, pre_max_size=pre, post_max_size=post)
    want = oracle.rotate_nms_gpu(d, thr, pre, post)
    assert_keep_lists_agree(got, want, d, thr, oracle, tol=1e-6, pre_max_size=pre, post_max_size=post)
    sb = oracle.rbox_to_standup(d[:, :5])
    g2, w2 = pp.nms(sb, d[:, 5], pre, post, thr), oracle.nms(sb, d[:, 5], pre, post, thr)
    assert g2.tolist() == w2.tolist()
    # predict glue on the same number of anchors
    A, B, nc = n, 2, 1
    an = synth.anchors_stride(synth.D435)[rng.choice(10240, A, replace=True)]
    bp = rng.normal(0, 0.2, (B, A, 7)).astype(np.float32)
    cl = np.round(rng.normal(-1, 1.5, (B, A, nc)), 1).astype(np.float32)
    dr = rng.normal(0, 1, (B, A, 2)).astype(np.float32)
    mask = (rng.random((B, A)) < 0.5).astype(np.uint8)
    rect = rng.normal(0, 1, (B, 4, 4)).astype(np.float32)
    trv = rng.normal(0, 1, (B, 4, 4)).astype(np.float32)
    for rotated in (False, True):
        lid, cam, sc, lab, idx, cnt = pp.predict_arrays(bp, cl, dr, an, mask, rect, trv, num_class=nc, top_k=100, nms_pre_max_size=pre,
                                                        nms_post_max_size=post, nms_iou_threshold=thr, rotated=rotated)
        for b in range(B):
            w = oracle.predict_frame(bp[b], cl[b], dr[b], an, mask[b], rect[b], trv[b], top_k=100, pre_max_size=pre, post_max_size=post,
                                     iou_threshold=thr, rotated=rotated)
            k = int(cnt[b])
            assert k > 0 and np.array_equal(idx[b, :k], w["anchor_index"]), (seed, b, rotated)
            np.testing.assert_allclose(lid[b, :k], w["box3d_lidar"], rtol=1e-5, atol=1e-5)
            np.testing.assert_allclose(sc[b, :k], w["scores"], rtol=1e-6, atol=0)


@pytest.mark.parametrize("seed", range(4))
def test_ingest_fuzz(pp, oracle, seed):
    """N3: random record layouts, slices and NaN patterns; reference matrices => bit-identical."""
    from importlib import import_module
    ing = import_module(pp.__name__ + ".ingest")
    rng = np.random.default_rng(2000 + seed)
    n = int(rng.integers(0, 70000))
    ps = int(rng.choice([12, 16, 20, 32]))
    offs = sorted(rng.choice(np.arange(0, ps, 4), 3, replace=False).tolist())
    xyz = rng.normal(0, 3, (n, 3)).astype(np.float32)
    xyz[rng.random(n) < rng.uniform(0, 0.6)] = np.nan
    if n:
        xyz[rng.integers(0, n, 5), rng.integers(0, 3, 5)] = np.inf
    raw = rng.integers(0, 255, (n, ps), dtype=np.uint8)
    for j, o in enumerate(offs):
        raw[:, o:o + 4] = xyz[:, j:j + 1].view(np.uint8)
    start, step = int(rng.integers(0, 6)), int(rng.integers(1, 7))
    got = pp.pointcloud2_to_lidar(raw.tobytes(), start=start, step=step, point_step=ps, offsets=tuple(offs))
    want = oracle.pointcloud2_to_lidar(xyz, (ing.R_Y_NEG90, ing.R_X_POS90), ing.LIFT, start, step)
    assert got.shape == want.shape and np.array_equal(got, want), (seed, n, ps, offs, start, step)


@pytest.mark.parametrize("seed", range(3))
def test_large_n_nms_fuzz(pp, oracle, synth, seed):
    """The alive-stripe path (> 16 384 boxes) against the all-pairs path on the 16 000 best boxes: random size, threshold,
    box scale / aspect, clustering, post_max_size; rotated and standup kinds."""
    rng = np.random.default_rng(4242 + seed)
    n = int(rng.integers(16500, 50000))
    thr = float(rng.uniform(0.02, 0.85))
    d = synth.rotated_boxes(n, 3000 + seed, clustered=bool(rng.integers(0, 2)))
    d[:, :4] *= np.float32(rng.choice([0.3, 1.0, 4.0]))
    if seed % 2 == 0:
        d[:, 2] *= rng.uniform(0.2, 5.0, n).astype(np.float32)
        d[:, 3] *= rng.uniform(0.2, 5.0, n).astype(np.float32)
    post = None if seed == 1 else int(rng.integers(50, 5000))
    top = oracle.argsort_desc(d[:, 5])
    got = pp.rotate_nms_gpu(d, thr, post_max_size=post)
    want = [int(top[i]) for i in pp.rotate_nms_gpu(d[top[:16000]], thr, post_max_size=post)]
    assert got[:len(want)] == want
    sb = oracle.rbox_to_standup(d[:, :5]) * np.float32(10)
    ks = pp.nms(sb, d[:, 5], None, post, thr)
    ka = pp.nms(sb[top[:16000]], d[top[:16000], 5], None, post, thr)
    assert ks[:len(ka)].tolist() == [int(top[i]) for i in ka]
