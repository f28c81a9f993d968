"""Generate tests/golden/*.npz from the REFERENCE's own source (build container only).

Run:  python tests/golden/make_golden.py
Needs /root/reference (read-only) and numba; see oracle/ref_extract.py for how the reference
functions are loaded without copying them.  The fixtures are committed; the GPU box never runs
this script.  Inputs are stored beside the outputs so the fixtures do not depend on the synthetic
generators staying unchanged.
"""
import importlib
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_extract  # noqa: E402

synth = importlib.import_module("3d-object-detection-for-autonomous-navigation_b200.synth")
OUT = os.path.dirname(os.path.abspath(__file__))


def voxel_cases(ref):
    d, k = synth.D435, synth.KITTI
    cases = {}

    def run(name, pts, vs, pcr, P, rev, cap):
        v, c, n = ref.points_to_voxel(pts, vs, pcr, P, rev, cap)
        cases[name] = dict(points=pts, voxel_size=np.asarray(vs, np.float64),
                           coors_range=np.asarray(pcr, np.float64),
                           params_are_lists=np.array(not isinstance(vs, np.ndarray)),
                           max_points=np.array(P), reverse_index=np.array(rev),
                           max_voxels=np.array(cap), voxels=v.copy(), coors=c.copy(), num=n.copy())

    vs_d, pcr_d = np.array(d["voxel_size"]), np.array(d["point_cloud_range"])
    vs_k, pcr_k = np.array(k["voxel_size"]), np.array(k["point_cloud_range"])
    cloud = synth.d435_cloud(0)
    run("d435_f64_sub16", np.ascontiguousarray(cloud[3::16]), vs_d, pcr_d, 50, True, 12000)
    run("d435_f64_cap", np.ascontiguousarray(cloud[5::16]), vs_d, pcr_d, 5, True, 700)
    run("d435_f32_sub16", np.ascontiguousarray(cloud[7::16].astype(np.float32)), vs_d, pcr_d, 50, True, 12000)
    kc = synth.kitti_cloud(1)
    run("kitti_ring_cap", np.ascontiguousarray(kc[::4]), vs_k, pcr_k, 100, True, 3000)
    ks = synth.kitti_cloud(1, shuffled=True)
    run("kitti_shuf_cap", np.ascontiguousarray(ks[::4]), vs_k, pcr_k, 100, True, 3000)
    run("kitti_listparams_f32arith_xyz", np.ascontiguousarray(ks[1::4]), k["voxel_size"],
        k["point_cloud_range"], 8, False, 20000)
    run("uniform_break_early", synth.uniform_cloud(20000, k, 3), vs_k, pcr_k, 100, True, 500)
    run("empty", np.zeros((0, 4), np.float32), vs_k, pcr_k, 100, True, 100)
    run("all_outside", (synth.uniform_cloud(100, k, 4) + 1000).astype(np.float32), vs_k, pcr_k, 100, True, 100)
    return cases


PREDICT_CONFIG = {"num_class": 1, "model": {"second": {
    "num_class": 1, "encode_background_as_zeros": True, "use_direction_classifier": True, "use_multi_class_nms": False,
    "nms_score_threshold": 0.0, "nms_pre_max_size": 100, "nms_post_max_size": 50, "nms_iou_threshold": 0.5,
    "use_sigmoid_score": True}}}  # configs/train.yaml:121-179


def predict_case(ref):
    """Three D435 frames through the reference's own predict(): random-init-like RPN outputs, a 60 % anchor
    mask, one frame with an empty mask (the None branch), a camera calibration with small random perturbation."""
    an = synth.anchors_stride(synth.D435)
    A, B = an.shape[0], 3
    rng = np.random.default_rng(5)
    bp = rng.normal(0, 0.1, (B, A, 7)).astype(np.float32)
    cl = rng.normal(-2, 1, (B, A, 1)).astype(np.float32)
    dr = rng.normal(0, 1, (B, A, 2)).astype(np.float32)
    mask = (rng.random((B, A)) < 0.6).astype(np.uint8)
    mask[2, 1000:] = 0   # few candidates: fewer than 100 survive the mask on a 80-anchor window
    mask[2, :920] = 0
    rect = np.tile(np.eye(4, dtype=np.float32), (B, 1, 1))
    trv = np.tile(np.array([[0, -1, 0, 0.01], [0, 0, -1, -0.07], [1, 0, 0, -0.27], [0, 0, 0, 1]], np.float32), (B, 1, 1))
    trv = (trv + rng.normal(0, 1e-3, (B, 4, 4))).astype(np.float32)
    anchors = np.tile(an, (B, 1, 1))
    example = [None, None, None, rect, trv, rect, anchors, mask, np.arange(B)]
    res = ref.predict(example, {"box_preds": bp, "cls_preds": cl, "dir_cls_preds": dr}, PREDICT_CONFIG)
    out = dict(box_preds=bp, cls_preds=cl, dir_cls_preds=dr, anchors=anchors, anchors_mask=mask, rect=rect, Trv2c=trv)
    for b, r in enumerate(res):
        out[f"count{b}"] = np.array(0 if r["box3d_lidar"] is None else r["box3d_lidar"].shape[0])
        for k in ("box3d_lidar", "box3d_camera", "scores", "label_preds", "bbox"):
            if r[k] is not None:
                out[f"{k}{b}"] = np.asarray(r[k])
        print("predict frame", b, int(out[f"count{b}"]))
    return out


def decorate_scatter_case(ref):
    out = {}
    for cfg, pts in ((synth.D435, synth.d435_cloud(2, True)[::6]), (synth.KITTI, synth.kitti_cloud(2)[::12])):
        n_ = cfg["name"]
        vs, pcr = np.array(cfg["voxel_size"]), np.array(cfg["point_cloud_range"])
        v, c, num = ref.points_to_voxel(pts, vs, pcr, cfg["max_points"], True, cfg["max_voxels"])
        v = v[:600].astype(np.float32); c = c[:600]; num = num[:600]
        c4 = np.concatenate([np.zeros((c.shape[0], 1), np.int32), c], axis=1)
        dec = ref.pillar_decorate(v, num, c4, cfg["voxel_size"], cfg["point_cloud_range"])
        nx, ny, _ = synth.grid_size(cfg)
        half = c.shape[0] // 2
        c4b = np.concatenate([c4, np.concatenate([np.ones((half, 1), np.int32), c[:half]], axis=1)])
        c4b[-1, 2:] = c4b[-2, 2:]          # an exact duplicate (b,y,x): scatter_nd adds
        feats = synth.pfn_standin(c4b.shape[0], 8, 1)
        canvas = ref.pointpillars_scatter(feats, c4b, 2, 8, ny, nx)
        out.update({f"{n_}_voxels": v, f"{n_}_num": num.copy(), f"{n_}_coors": c4, f"{n_}_decorated": dec,
                    f"{n_}_voxel_size": vs, f"{n_}_range": pcr, f"{n_}_feats": feats, f"{n_}_coords": c4b,
                    f"{n_}_canvas": canvas})
        print("decorate/scatter", n_, dec.shape, canvas.shape, int((canvas != 0).sum()))
    return out


def main():
    warnings.simplefilter("ignore")
    ref = ref_extract.load()

    for name, c in voxel_cases(ref).items():
        np.savez_compressed(os.path.join(OUT, f"voxel_{name}.npz"), **c)
        print("voxel", name, c["voxels"].shape, int(c["num"].sum()))

    # rotated IoU / NMS (second/core/non_max_suppression/nms_gpu.py)
    d = synth.rotated_boxes(320, 11, clustered=True)
    q = synth.rotated_boxes(150, 12, clustered=True)
    q[:, :2] = d[:150, :2] + np.random.default_rng(5).normal(0, 0.7, size=(150, 2)).astype(np.float32)
    out = dict(boxes=d[:, :5].copy(), query=q[:, :5].copy(), dets=d)
    for crit in (-1, 0, 1, 2):
        out[f"iou_crit{crit}"] = ref.rotate_iou_matrix(out["boxes"], out["query"], crit)
    for thr in (0.5, 0.1, 0.01):
        keep, iou_all = ref.rotate_nms(d, thr)
        out[f"keep_thr{thr}"] = np.asarray(keep, np.int64)
        # pairs whose IoU is within 1e-6 of the threshold are excluded from index parity
        out[f"near_thr{thr}"] = np.array(int((np.abs(iou_all - np.float32(thr)) < 1e-6).sum()))
    table = np.array([
        [1, 1, 2, 1, 0, 2, 1, 2, 1, 0], [1, 1, 2, 1, 0, 1.5, 1.5, 2, 1, 0],
        [0, 0, 2, 1, 0, 0, 0, 2, 1, np.pi / 2], [0, 0, 2, 2, 0, 0, 0, 2, 2, np.pi / 4],
        [0, 0, 4, 4, 0.1, 0, 0, 1, 1, 0.5], [0, 0, 1, 1, 0.2, 5, 5, 1, 1, 0.7],
        [1, 1, 2, 1, 0, 1, 1, 2, 1, 0]], np.float32)
    out["table"] = table
    out["table_iou"] = np.array([ref.rotate_iou_matrix(t[None, :5].copy(), t[None, 5:].copy(), -1)[0, 0]
                                 for t in table], np.float32)
    np.savez_compressed(os.path.join(OUT, "rotated.npz"), **out)
    print("rotated", {k: v.shape for k, v in out.items() if k.startswith("keep")})

    # live standup NMS (libraries/eval_helper_functions.py:463-598) + prep (load_data.py:1525-1594)
    corners = ref.center_to_corner_box2d(d[:, :2], d[:, 2:4], d[:, 4])
    standup = ref.corner_to_standup_nd_jit(corners)
    so = dict(rboxes=d[:, :5].copy(), standup=standup, scores=d[:, 5].copy())
    for scale in (1.0, 10.0):
        dets = np.concatenate([standup * np.float32(scale), d[:, 5:6]], axis=1).astype(np.float32)
        for thr in (0.5, 0.3):
            keep, iou_all = ref.standup_nms(dets, thr)
            so[f"keep_s{scale}_thr{thr}"] = np.asarray(keep, np.int64)
    np.savez_compressed(os.path.join(OUT, "standup.npz"), **so)

    # decode (libraries/eval_helper_functions.py:388-461)
    an = synth.anchors_stride(synth.D435)[::7][:1500].copy()
    be, _ = synth.rpn_standin(an.shape[0], 3)
    be[:, 3:6] *= 8  # exercise exp over a wider range
    np.savez_compressed(os.path.join(OUT, "decode.npz"), box_encodings=be, anchors=an,
                        decoded=ref.second_box_decode(be, an))
    # anchor mask, "next" row N1 (load_data.py:3043-3072)
    am = {}
    for cfg, fs, pts, stride in ((synth.D435, [1, 64, 80], synth.d435_cloud(3, True)[:30000], 1),
                                 (synth.KITTI, [1, 248, 216], synth.kitti_cloud(3), 16)):
        vs, pcr = np.array(cfg["voxel_size"]), np.array(cfg["point_cloud_range"])
        anchors = ref.create_anchors_3d_stride(fs, cfg["anchor_sizes"], cfg["anchor_strides"], cfg["anchor_offsets"],
                                               cfg["anchor_rotations"]).reshape(-1, 7)[::stride].copy()
        _, coors, _ = ref.points_to_voxel(pts, vs, pcr, cfg["max_points"], True, cfg["max_voxels"])
        area, mask = ref.anchors_mask(coors, anchors, vs, pcr, 1)
        n = cfg["name"]
        am.update({f"{n}_anchors": anchors, f"{n}_coors": coors.copy(), f"{n}_voxel_size": vs, f"{n}_range": pcr,
                   f"{n}_area": area, f"{n}_mask": mask})
        print("anchor mask", n, anchors.shape, float(mask.mean()))
    np.savez_compressed(os.path.join(OUT, "anchor_mask.npz"), **am)
    # 3-D overlap of the KITTI eval, "next" row N4 (second/utils/eval.py:131-163)
    b, q = synth.camera_boxes(260, 1), synth.camera_boxes(140, 2)
    d3 = dict(boxes=b, query=q)
    for crit in (-1, 0, 1, 2):
        d3[f"d3_crit{crit}"] = ref.d3_box_overlap(b, q, crit)
    np.savez_compressed(os.path.join(OUT, "d3_overlap.npz"), **d3)
    # decoration + scatter: the reference's own method bodies over oracle/tf_shim.py (model/pointpillars.py:128-203, 285-341)
    np.savez_compressed(os.path.join(OUT, "decorate_scatter.npz"), **decorate_scatter_case(ref))
    # VoxelNet.predict post-network half, "next" row N2 (model/voxelnet.py:1060-1389)
    np.savez_compressed(os.path.join(OUT, "predict.npz"), **predict_case(ref))
    print("done")


if __name__ == "__main__":
    main()
