"""GPU: zero-copy device entry points (DLPack capsules, __dlpack__, __cuda_array_interface__)."""
import importlib

import numpy as np
import pytest

from conftest import PKG

pytestmark = pytest.mark.gpu


class _CAI:
    """Minimal __cuda_array_interface__ exporter (what numba device arrays look like)."""

    def __init__(self, t):
        self._t = t
        self.__cuda_array_interface__ = t.__cuda_array_interface__


def test_dlpack_device_path_matches_oracle(oracle, synth):
    import torch
    from torch.utils.dlpack import to_dlpack
    iop = importlib.import_module(PKG + ".interop")
    cfg = synth.KITTI
    vs, pcr = cfg["voxel_size"], cfg["point_cloud_range"]
    frames = [synth.kitti_cloud(60), synth.kitti_cloud(61, True)[:70000]]
    pts = np.concatenate(frames)
    off = np.array([0, frames[0].shape[0], pts.shape[0]], np.int64)
    d_pts = torch.from_numpy(pts).cuda()
    out = iop.points_to_voxel(to_dlpack(d_pts), vs, pcr, 40, True, 9000, frame_offsets=_CAI(torch.from_numpy(off).cuda()),
                              decorate=True, max_frame_points=max(f.shape[0] for f in frames))
    vb = out["voxel_base"].cpu().numpy()
    M = int(vb[-1])
    vx, vy = vs[:2]
    xo, yo = vx / 2 + pcr[0], vy / 2 + pcr[1]
    nx, ny, _ = synth.grid_size(cfg)
    for b, f in enumerate(frames):
        ov, oc, on = oracle.points_to_voxel(f, np.array(vs), np.array(pcr), 40, True, 9000)
        lo, hi = vb[b], vb[b + 1]
        assert np.array_equal(out["voxels"][lo:hi].cpu().numpy(), ov)
        assert np.array_equal(out["coors"][lo:hi, 1:].cpu().numpy(), oc)
        assert np.array_equal(out["num_points"][lo:hi].cpu().numpy(), on)
    # standalone decoration on the produced tensors == fused decoration
    dec = iop.pillar_decorate(out["voxels"][:M], out["num_points"][:M], out["coors"][:M], vx, vy, xo, yo)
    np.testing.assert_allclose(dec.cpu().numpy(), out["decorated"][:M].cpu().numpy(), rtol=1e-5, atol=1e-3)
    # scatter with the row count left on the device
    feats = torch.from_numpy(synth.pfn_standin(out["coors"].shape[0], 32, 1)).cuda()
    canvas = iop.scatter(feats, out["coors"], 2, ny, nx, "NCHW", num_rows=out["voxel_base"][2:3])
    want = oracle.scatter(feats[:M].cpu().numpy(), out["coors"][:M].cpu().numpy(), 2, ny, nx)
    assert np.array_equal(canvas.cpu().numpy(), want)
    # decode (one anchor set for the batch) + batched rotated NMS on decoded boxes, in place
    an = synth.anchors_stride(synth.D435)
    A = an.shape[0]
    box = np.stack([synth.rpn_standin(A, 5)[0], synth.rpn_standin(A, 6)[0]])
    sco = np.stack([synth.rpn_standin(A, 5)[1], synth.rpn_standin(A, 6)[1]])
    dec_boxes = iop.second_box_decode(torch.from_numpy(box).cuda(), torch.from_numpy(an).cuda())
    keep, cnt = iop.nms(dec_boxes, torch.from_numpy(sco).cuda(), 100, 50, 0.5, rotated=True)
    for b in range(2):
        ob = oracle.second_box_decode(box[b], an)
        dets = np.concatenate([ob[:, [0, 1, 3, 4, 6]], sco[b][:, None]], axis=1)
        want = oracle.rotate_nms_gpu(dets, 0.5, 100, 50)
        assert keep[b, :int(cnt[b])].cpu().numpy().tolist() == want
    iou = iop.rotate_iou(torch.from_numpy(dets[:50, :5].copy()).cuda(), torch.from_numpy(dets[:20, :5].copy()).cuda())
    np.testing.assert_allclose(iou.cpu().numpy(), oracle.rotate_iou_gpu_eval(dets[:50, :5], dets[:20, :5]), atol=1e-5)
    with pytest.raises(ValueError):
        iop.points_to_voxel(torch.from_numpy(pts), vs, pcr, 40, True, 9000)  # host tensor: no silent CPU path


def test_tf_adaptor_with_torch_standing_in_for_tensorflow(oracle, synth):
    """The TF shim (model/voxelnet.py:867, 881 call sites) exchanges tensors through
    `tf.experimental.dlpack.to_dlpack / from_dlpack`; TensorFlow is not installed here, so a namespace with the same
    two functions backed by torch's DLPack stands in for it.  Results must be the device path's (== the oracle's)."""
    import types
    import torch
    tfa = importlib.import_module(PKG + ".tf_adaptor")
    fake_tf = types.SimpleNamespace(experimental=types.SimpleNamespace(dlpack=types.SimpleNamespace(
        to_dlpack=lambda t: torch.utils.dlpack.to_dlpack(t), from_dlpack=lambda c: torch.utils.dlpack.from_dlpack(c))))
    tfa.use_tf_module(fake_tf)
    cfg = synth.D435
    vs, pcr = np.array(cfg["voxel_size"]), np.array(cfg["point_cloud_range"])
    v, c, n = oracle.points_to_voxel(synth.d435_cloud(4, subsample=True), vs, pcr, cfg["max_points"], True, cfg["max_voxels"])
    v = v.astype(np.float32)
    c4 = np.concatenate([np.zeros((c.shape[0], 1), np.int32), c], axis=1)
    vx, vy = cfg["voxel_size"][:2]
    xo, yo = vx / 2 + pcr[0], vy / 2 + pcr[1]
    dev = torch.device("cuda", 0)
    tv, tn, tc = torch.from_numpy(v).to(dev), torch.from_numpy(n).to(dev), torch.from_numpy(c4).to(dev)
    dec = tfa.pillar_decorate(tv, tn, tc, vx, vy, xo, yo)
    want = oracle.decorate(v, n, c4, vx, vy, xo, yo)
    np.testing.assert_allclose(dec.cpu().numpy(), want, rtol=1e-5, atol=1e-5 * (float(np.abs(v).max()) + 1))
    nx, ny, _ = synth.grid_size(cfg)
    feats = synth.pfn_standin(c4.shape[0], cfg["num_filters"], 1)
    config = {"model": {"second": {"voxel_feature_extractor": {"num_filters": cfg["num_filters"]},
                                   "voxel_generator": {"voxel_size": cfg["voxel_size"], "point_cloud_range": cfg["point_cloud_range"]}}},
              "train_input_reader": {"batch_size": 2}, "eval_input_reader": {"batch_size": 1}}
    layer = tfa.PointPillarsScatter(config, training=False)
    assert (layer.batch_size, layer.ny, layer.nx, layer.nchannels) == (1, ny, nx, cfg["num_filters"])
    canvas = layer(torch.from_numpy(feats).to(dev), tc)
    assert np.array_equal(canvas.cpu().numpy(), oracle.scatter(feats, c4, 1, ny, nx))
