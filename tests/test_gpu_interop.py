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
