"""Pin the CPU oracle against the reference's own source executed live (build container only;
skipped where /root/reference is absent, e.g. on the GPU box)."""
import warnings

import numpy as np
import pytest

from oracle import ref_extract

pytestmark = pytest.mark.skipif(not ref_extract.available(), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref():
    warnings.simplefilter("ignore")
    return ref_extract.load()


def test_voxelizer_full_size(ref, oracle, synth):
    d, k = synth.D435, synth.KITTI
    cases = [
        (synth.d435_cloud(0), np.array(d["voxel_size"]), np.array(d["point_cloud_range"]), 50, True, 12000),
        (synth.d435_cloud(1, subsample=True), np.array(d["voxel_size"]), np.array(d["point_cloud_range"]), 50, True, 12000),
        (synth.kitti_cloud(0), np.array(k["voxel_size"]), np.array(k["point_cloud_range"]), 100, True, 12000),
        (synth.kitti_cloud(0, True), np.array(k["voxel_size"]), np.array(k["point_cloud_range"]), 100, True, 12000),
        (synth.kitti_cloud(2, True), k["voxel_size"], k["point_cloud_range"], 100, False, 12000),
    ]
    for args in cases:
        want = ref.points_to_voxel(*args)
        got = oracle.points_to_voxel(*args)
        for w, g in zip(want, got):
            assert w.dtype == g.dtype and np.array_equal(w, g)
    # SURVEY 8d probed numbers for the D435 synthetic cloud
    v, c, n = oracle.points_to_voxel(*cases[0])
    assert v.shape[0] == 8171 and int(n.sum()) == 217921 and set(np.unique(c[:, 0])) == {0, 1}


def test_rotated_iou_random(ref, oracle, synth):
    for seed, clustered in ((21, True), (22, False)):
        d = synth.rotated_boxes(400, seed, clustered)
        for crit in (-1, 0, 1, 2):
            want = ref.rotate_iou_matrix(d[:, :5].copy(), d[:128, :5].copy(), crit)
            got = oracle.rotate_iou_gpu_eval(d[:, :5], d[:128, :5], crit)
            off = ~np.eye(400, 128, dtype=bool)  # identical boxes are chaotic in the reference (F10)
            np.testing.assert_allclose(got[off], want[off], rtol=0, atol=1e-6)
        keep, iou_all = ref.rotate_nms(d, 0.3)
        if int((np.abs(iou_all - np.float32(0.3)) < 1e-6).sum()) == 0:
            assert oracle.rotate_nms_gpu(d, 0.3) == keep


def test_decode_and_standup(ref, oracle, synth):
    an = synth.anchors_stride(synth.D435)
    be, sc = synth.rpn_standin(an.shape[0], 5)
    np.testing.assert_allclose(oracle.second_box_decode(be, an), ref.second_box_decode(be, an), rtol=1e-6, atol=1e-6)
    d = synth.rotated_boxes(500, 3)
    want = ref.corner_to_standup_nd_jit(ref.center_to_corner_box2d(d[:, :2], d[:, 2:4], d[:, 4]))
    np.testing.assert_allclose(oracle.rbox_to_standup(d[:, :5]), want, rtol=1e-6, atol=1e-5)
    dets = np.concatenate([want, d[:, 5:6]], axis=1)
    keep, _ = ref.standup_nms(dets, 0.5)
    assert oracle.nms(want, d[:, 5], None, None, 0.5).tolist() == keep


def test_anchors_and_mask(ref, oracle, synth):
    for cfg, fs in ((synth.D435, [1, 64, 80]), (synth.KITTI, [1, 248, 216])):
        ra = ref.create_anchors_3d_stride(fs, cfg["anchor_sizes"], cfg["anchor_strides"], cfg["anchor_offsets"],
                                          cfg["anchor_rotations"]).reshape(-1, 7)
        an = synth.anchors_stride(cfg)
        assert ra.dtype == an.dtype and np.array_equal(ra, an)
        vs, pcr = np.array(cfg["voxel_size"]), np.array(cfg["point_cloud_range"])
        pts = synth.d435_cloud(9, True) if cfg["name"] == "d435i" else synth.kitti_cloud(9, True)
        _, c, _ = oracle.points_to_voxel(pts, vs, pcr, cfg["max_points"], True, cfg["max_voxels"])
        assert np.array_equal(ref.rbbox2d_to_near_bbox(ra[:, [0, 1, 3, 4, 6]]), oracle.rbbox2d_to_near_bbox(an[:, [0, 1, 3, 4, 6]]))
        for thr in (1, 0, 4):
            wa, wm = ref.anchors_mask(c, ra, vs, pcr, thr)
            ga, gm = oracle.anchors_mask(c, an, vs, pcr, thr)
            assert np.array_equal(wa, ga) and np.array_equal(wm, gm)
    # anchors shifted left / down by less than one grid width: the upper index stays negative after the one-sided
    # clip and numba wraps it around once -- the oracle reproduces that (anything farther out is undefined
    # behaviour in the reference and is not run here)
    cfg = synth.D435
    vs, pcr = np.array(cfg["voxel_size"]), np.array(cfg["point_cloud_range"])
    _, c, _ = oracle.points_to_voxel(synth.d435_cloud(9, True), vs, pcr, cfg["max_points"], True, cfg["max_voxels"])
    an = synth.anchors_stride(cfg)[::7].copy()
    an[::2, 0] -= 4.0
    an[1::3, 1] -= 3.0
    wa, wm = ref.anchors_mask(c, an, vs, pcr, 1)
    ga, gm = oracle.anchors_mask(c, an, vs, pcr, 1)
    assert np.array_equal(wa, ga) and np.array_equal(wm, gm)


def test_d3_box_overlap(ref, oracle, synth):
    b, q = synth.camera_boxes(300, 7), synth.camera_boxes(180, 8)
    for crit in (-1, 0, 1, 2):
        np.testing.assert_allclose(oracle.d3_box_overlap(b, q, crit), ref.d3_box_overlap(b, q, crit), rtol=0, atol=1e-6)


def test_decorate_scatter_full_size(ref, oracle, synth):
    """a4 / a5 at full size: the reference's PillarFeatureNet.call (143-203) / PointPillarsScatter.call (285-341)
    bodies over the numpy TF stand-in vs the C oracle -- decoration to 1e-6, canvas bit for bit."""
    for cfg, pts in ((synth.D435, synth.d435_cloud(5)), (synth.KITTI, synth.kitti_cloud(5))):
        vs, pcr = np.array(cfg["voxel_size"]), np.array(cfg["point_cloud_range"])
        v, c, n = oracle.points_to_voxel(pts, vs, pcr, cfg["max_points"], True, cfg["max_voxels"])
        v = v.astype(np.float32)
        c4 = np.concatenate([np.zeros((c.shape[0], 1), np.int32), c], axis=1)
        want = ref.pillar_decorate(v, n, c4, cfg["voxel_size"], cfg["point_cloud_range"])
        got = oracle.decorate(v, n, c4, vs[0], vs[1], vs[0] / 2 + pcr[0], vs[1] / 2 + pcr[1])
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-6)
        nx, ny, _ = synth.grid_size(cfg)
        c4b = np.concatenate([c4, np.concatenate([np.ones((c.shape[0], 1), np.int32), c], axis=1)])
        feats = synth.pfn_standin(c4b.shape[0], cfg["num_filters"], 2)
        assert np.array_equal(oracle.scatter(feats, c4b, 2, ny, nx),
                              ref.pointpillars_scatter(feats, c4b, 2, cfg["num_filters"], ny, nx))


def test_live_nms_wrapper(ref, oracle, synth):
    """a7 wrapper: nms(bboxes, scores, pre_max_size, post_max_size, iou_threshold), eval_helper_functions.py:463-492,
    run as written (its numba.cuda kernel replaced by the CPU run of the same body): index arrays and the None sentinel."""
    d = synth.rotated_boxes(600, 9, clustered=True)
    sb = oracle.rbox_to_standup(d[:, :5]) * np.float32(10)
    sc = d[:, 5].copy()
    for pre, post, thr in ((100, 50, 0.5), (None, None, 0.3), (600, 7, 0.7), (17, 100, 0.1)):
        want = ref.nms(sb, sc, pre_max_size=pre, post_max_size=post, iou_threshold=thr)
        got = oracle.nms(sb, sc, pre, post, thr)
        assert got.dtype == np.int64 and np.array_equal(got, np.asarray(want, np.int64)), (pre, post, thr)
    assert ref.nms(sb[:0], sc[:0], pre_max_size=None, post_max_size=50, iou_threshold=0.5) is None
    assert oracle.nms(sb[:0], sc[:0], None, 50, 0.5) is None
