"""Seeded synthetic inputs for the hot path (SURVEY 8d generators).

No dataset or checkpoint is reachable, so tests and bench.py use these clouds, pillar features
and RPN stand-ins.  Pure numpy; shapes and dtypes follow the reference's call sites.
"""
from __future__ import annotations

import numpy as np

# configs/train.yaml:112,115,117,120 (voxel_generator), 125 (num_filters), 176-179 (nms),
# 186-194 (anchor generator) of the reference.
D435 = dict(
    name="d435i",
    point_cloud_range=[0.0, -2.56, -3.0, 6.40, 2.56, 3.0],
    voxel_size=[0.08, 0.08, 4.0],
    max_points=50,
    max_voxels=12000,
    num_point_features=3,
    point_dtype="float64",
    num_filters=128,
    anchor_sizes=[0.6, 0.8, 1.73],
    anchor_strides=[0.08, 0.08, 0.0],
    anchor_offsets=[0.08, -2.56, -1.465],
    anchor_rotations=[0, 1.57],
    nms_pre_max_size=100,
    nms_post_max_size=50,
    nms_iou_threshold=0.5,
)

# canonical PointPillars KITTI car config (BASELINE.json configs[2]; SURVEY 8 shorthand)
KITTI = dict(
    name="kitti",
    point_cloud_range=[0.0, -39.68, -3.0, 69.12, 39.68, 1.0],
    voxel_size=[0.16, 0.16, 4.0],
    max_points=100,
    max_voxels=12000,
    num_point_features=4,
    point_dtype="float32",
    num_filters=64,
    anchor_sizes=[1.6, 3.9, 1.56],
    anchor_strides=[0.32, 0.32, 0.0],
    anchor_offsets=[0.16, -39.52, -1.78],
    anchor_rotations=[0, 1.57],
    nms_pre_max_size=1000,
    nms_post_max_size=300,
    nms_iou_threshold=0.5,
)


def grid_size(cfg):
    """np.round((hi-lo)/vs).astype(int32): load_data.py:612-615 -> [nx, ny, nz]."""
    r = np.asarray(cfg["point_cloud_range"], np.float64)
    v = np.asarray(cfg["voxel_size"], np.float64)
    return np.round((r[3:] - r[:3]) / v).astype(np.int32).tolist()


def d435_cloud(seed=0, subsample=False):
    """848x480 depth image, depth~U(0.3,8), pinhole 87x58 deg -> float64 [407040,3].
    `subsample` applies the production path's [1::4] (load_data.py:2434)."""
    W, H = 848, 480
    rng = np.random.default_rng(seed)
    depth = rng.uniform(0.3, 8.0, size=(H, W))
    fx = (W / 2) / np.tan(np.deg2rad(87.0) / 2)
    fy = (H / 2) / np.tan(np.deg2rad(58.0) / 2)
    u, v = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    x = depth
    y = -(u - W / 2) / fx * depth
    z = -(v - H / 2) / fy * depth + 1.0
    pts = np.stack([x, y, z], axis=-1).reshape(-1, 3).astype(np.float64)
    return np.ascontiguousarray(pts[1::4]) if subsample else pts


def d435_sensor_cloud(seed=0, invalid=0.1):
    """The same scene as d435_cloud as the sensor publishes it (sensor_msgs/PointCloud2 xyz, float32, camera
    optical frame, invalid pixels NaN): float32 [407040,3] such that the reference's ingest
    (load_data.py:2434-2443: finite rows, [1::4], two rotations, +1 m) maps it into the D435 grid."""
    lidar = d435_cloud(seed)
    x, y, z = lidar[:, 0], lidar[:, 1], lidar[:, 2] - 1.0
    # inverse of p @ R_y(-90) @ R_x(+90): lidar (x,y,z) = (s2, s0, -s1) -> sensor (s0,s1,s2) = (y, -z, x)
    sensor = np.stack([y, -z, x], axis=1).astype(np.float32)
    rng = np.random.default_rng(seed + 77_000)
    sensor[rng.random(sensor.shape[0]) < invalid] = np.nan
    return np.ascontiguousarray(sensor)


def kitti_cloud(seed=0, shuffled=False):
    """HDL-64-like: 64 rings x 1875 azimuth steps, float32 [120000,4] ring-major."""
    rng = np.random.default_rng(seed)
    rings, steps = 64, 1875
    elev = np.deg2rad(np.linspace(-24.8, 2.0, rings))[:, None]
    az = np.linspace(-np.pi, np.pi, steps, endpoint=False)[None, :]
    free = rng.uniform(5.0, 80.0, size=(rings, steps))
    with np.errstate(divide="ignore"):
        ground = np.where(np.sin(elev) < 0, 1.73 / -np.sin(elev), np.inf)
    r = np.minimum(ground, free)
    x = r * np.cos(elev) * np.cos(az)
    y = r * np.cos(elev) * np.sin(az)
    z = r * np.sin(elev)
    inten = rng.uniform(0.0, 1.0, size=(rings, steps))
    pts = np.stack([x, y, z, inten], axis=-1).reshape(-1, 4).astype(np.float32)
    if shuffled:
        pts = pts[np.random.default_rng(seed + 10_000).permutation(pts.shape[0])]
    return np.ascontiguousarray(pts)


def uniform_cloud(n, cfg, seed=0, dtype=np.float32, margin=0.05):
    """Uniform points over the range inflated by `margin` (some fall outside)."""
    rng = np.random.default_rng(seed)
    r = np.asarray(cfg["point_cloud_range"], np.float64)
    lo, hi = r[:3], r[3:]
    ext = (hi - lo) * margin
    D = cfg["num_point_features"]
    pts = rng.uniform(lo - ext, hi + ext, size=(n, 3))
    if D > 3:
        pts = np.concatenate([pts, rng.uniform(0, 1, size=(n, D - 3))], axis=1)
    return np.ascontiguousarray(pts.astype(dtype))


def anchors_stride(cfg):
    """Strided anchors [nz(=1)*ny'*nx'*rot, 7] (x,y,z,w,l,h,r), the layout the reference's
    create_anchors_3d_stride (load_data.py:1598-1638) + reshape (1663) produce.  Feature map =
    BEV grid / (layer_strides[0] // upsample_strides[0]) = BEV grid for the reference config."""
    nx, ny, _ = grid_size(cfg)
    sx, sy, sz = cfg["anchor_strides"]
    ox, oy, oz = cfg["anchor_offsets"]
    step = max(1, int(round(sx / cfg["voxel_size"][0])))
    fx, fy = nx // step, ny // step
    xc = np.arange(fx, dtype=np.float32) * np.float32(sx) + np.float32(ox)
    yc = np.arange(fy, dtype=np.float32) * np.float32(sy) + np.float32(oy)
    zc = np.arange(1, dtype=np.float32) * np.float32(sz) + np.float32(oz)
    rot = np.asarray(cfg["anchor_rotations"], np.float32)
    size = np.asarray(cfg["anchor_sizes"], np.float32)
    a = np.zeros((1, fy, fx, rot.shape[0], 7), np.float32)
    a[..., 0] = xc[None, None, :, None]
    a[..., 1] = yc[None, :, None, None]
    a[..., 2] = zc[:, None, None, None]
    a[..., 3:6] = size
    a[..., 6] = rot[None, None, None, :]
    return a.reshape(-1, 7)


def rpn_standin(num_anchors, seed=0):
    """box_preds ~ N(0,0.1) [A,7], cls logits ~ N(-2,1) [A]."""
    rng = np.random.default_rng(seed + 777)
    box = rng.normal(0.0, 0.1, size=(num_anchors, 7)).astype(np.float32)
    cls = rng.normal(-2.0, 1.0, size=(num_anchors,)).astype(np.float32)
    scores = (1.0 / (1.0 + np.exp(-cls.astype(np.float64)))).astype(np.float32)
    return box, scores


def pfn_standin(num_rows, channels, seed=0):
    """Stand-in for the PFN Dense/BN/ReLU/max output: N(0,1) float32 [num_rows, C]."""
    rng = np.random.default_rng(seed + 4242)
    return rng.normal(0.0, 1.0, size=(num_rows, channels)).astype(np.float32)


def rotated_boxes(n, seed=0, clustered=False, cfg=KITTI):
    """[n,6] (x,y,w,l,angle,score): car-sized boxes, distinct scores, no duplicate boxes."""
    rng = np.random.default_rng(seed)
    r = cfg["point_cloud_range"]
    if clustered:
        k = max(1, n // 40)
        cx = rng.uniform(r[0], r[3], size=k)
        cy = rng.uniform(r[1], r[4], size=k)
        which = rng.integers(0, k, size=n)
        x = cx[which] + rng.normal(0, 1.0, size=n)
        y = cy[which] + rng.normal(0, 1.0, size=n)
    else:
        x = rng.uniform(r[0], r[3], size=n)
        y = rng.uniform(r[1], r[4], size=n)
    w = rng.uniform(1.4, 1.9, size=n)
    l = rng.uniform(3.2, 4.8, size=n)
    a = rng.uniform(-np.pi, np.pi, size=n)
    s = rng.permutation(np.linspace(0.0, 1.0, n))
    return np.stack([x, y, w, l, a, s], axis=1).astype(np.float32)


def camera_boxes(n, seed=0):
    """KITTI-eval style camera boxes [n,7] float64 (x,y,z,l,h,w,ry) in loose clusters (so pairs overlap)."""
    r = np.random.default_rng(seed)
    k = max(1, n // 25)
    c = r.integers(0, k, n)
    cx = r.uniform(-10, 10, k)[c] + r.normal(0, 1.2, n)
    cz = r.uniform(5, 40, k)[c] + r.normal(0, 1.2, n)
    y = r.uniform(1.0, 2.0, n)
    return np.stack([cx, y, cz, r.uniform(3.2, 4.8, n), r.uniform(1.3, 1.9, n), r.uniform(1.4, 1.9, n),
                     r.uniform(-np.pi, np.pi, n)], axis=1)
