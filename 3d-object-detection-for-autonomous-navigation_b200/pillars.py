"""Drop-ins for the two pillar stages of model/pointpillars.py in the reference.

  * `pillar_decorate` / `PillarFeatureNet.decorate`: lines 143-203 of PillarFeatureNet.call (the
    Dense/BN/ReLU/max at 211-225 stay in the host framework).
  * `PointPillarsScatter.call`: lines 285-341.
numpy in, numpy out; the device-resident variants live in pipeline.py.
"""
from __future__ import annotations

import numpy as np

from . import _lib


def pillar_decorate(voxels, num_points, coors, vx, vy, x_offset, y_offset, device=None):
    """voxels [M,P,D] f32, num_points [M] i32, coors [M,4] (batch,z,y,x) -> [M,P,D+5] f32."""
    voxels = np.ascontiguousarray(voxels, np.float32)
    num_points = np.ascontiguousarray(num_points, np.int32)
    coors = np.ascontiguousarray(coors, np.int32)
    if voxels.ndim != 3 or coors.shape != (voxels.shape[0], 4) or num_points.shape != (voxels.shape[0],):
        raise ValueError("expected voxels [M,P,D], num_points [M], coors [M,4]")
    M, P, D = voxels.shape
    out = np.empty((M, P, D + 5), np.float32)
    c = _lib.ctx(device)
    _lib.check(_lib.lib().pp_decorate_host(c.handle, _lib.ptr(voxels), _lib.ptr(num_points), _lib.ptr(coors),
                                           M, P, D, float(vx), float(vy), float(x_offset), float(y_offset),
                                           _lib.ptr(out)))
    return out


class PillarFeatureNet:
    """Holds the constants of model/pointpillars.py:121-124 and exposes the decoration."""

    def __init__(self, config):
        vg = config["model"]["second"]["voxel_generator"]
        self.vx = vg["voxel_size"][0]
        self.vy = vg["voxel_size"][1]
        self.x_offset = self.vx / 2 + vg["point_cloud_range"][0]
        self.y_offset = self.vy / 2 + vg["point_cloud_range"][1]

    def decorate(self, voxels, num_points, coors):
        return pillar_decorate(voxels, num_points, coors, self.vx, self.vy, self.x_offset, self.y_offset)


class PointPillarsScatter:
    """model/pointpillars.py:240-341: __init__ fixes batch_size / ny / nx / nchannels from the
    config exactly as the reference does (254-275); call(voxel_features, coords) -> [B,C,ny,nx]."""

    def __init__(self, config, training=False, layout="NCHW"):
        sec = config["model"]["second"]
        self.nchannels = sec["voxel_feature_extractor"]["num_filters"]
        self.batch_size = (config["train_input_reader"] if training else config["eval_input_reader"])["batch_size"]
        voxel_size = np.array(sec["voxel_generator"]["voxel_size"])
        pcr = np.array(sec["voxel_generator"]["point_cloud_range"])
        grid = np.round((pcr[3:] - pcr[:3]) / voxel_size).astype(np.int64)
        self.nx, self.ny = int(grid[0]), int(grid[1])
        self.layout = layout

    def call(self, voxel_features, coords):
        return scatter(voxel_features, coords, self.batch_size, self.ny, self.nx, self.layout)

    __call__ = call


def scatter(voxel_features, coords, batch_size, ny, nx, layout="NCHW", device=None):
    f = np.ascontiguousarray(voxel_features, np.float32)
    co = np.ascontiguousarray(coords, np.int32)
    if f.ndim != 2 or co.shape != (f.shape[0], 4):
        raise ValueError("expected voxel_features [M,C] and coords [M,4]")
    M, Cc = f.shape
    nhwc = layout == "NHWC"
    out = np.empty((batch_size, ny, nx, Cc) if nhwc else (batch_size, Cc, ny, nx), np.float32)
    c = _lib.ctx(device)
    _lib.check(_lib.lib().pp_scatter_host(c.handle, _lib.ptr(f), _lib.ptr(co), M, Cc, int(batch_size), int(ny),
                                          int(nx), _lib.PP_LAYOUT_NHWC if nhwc else _lib.PP_LAYOUT_NCHW,
                                          _lib.ptr(out)))
    return out
