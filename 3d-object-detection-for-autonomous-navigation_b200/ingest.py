"""Drop-in for the sensor ingest of the production path ("next" row N3), load_data.py:2434-2443:

    points = ros_numpy.point_cloud2.pointcloud2_to_xyz_array(self.production_pc)[1::4]
    r = R.from_euler('y', -90, degrees=True).as_dcm(); r2 = R.from_euler('x', 90, degrees=True).as_dcm()
    points = np.dot(points, r); points = np.dot(points, r2)
    points = points + [0.0, 0.0, 1.0]

`pointcloud2_to_lidar(cloud, ...)` takes the PointCloud2 payload (a structured array with float32 fields
x, y, z such as ros_numpy.numpify(msg) returns, raw bytes + point_step, or a plain [N,3] float32 array) and
returns the float64 [M,3] array `points_to_voxel` is called with.  One H2D of the raw cloud, the finite-row
compaction, the slice, both rotations and the lift on the device.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

# scipy's R.from_euler('y', -90, degrees=True).as_matrix() and R.from_euler('x', 90, ...): exact values
# (2**-52 where cos(90 deg) would be); `as_dcm` is the same function under its pre-1.4 name
_E = 2.0 ** -52
R_Y_NEG90 = np.array([[_E, -0.0, -1.0], [0.0, 1.0, -0.0], [1.0, 0.0, _E]], np.float64)
R_X_POS90 = np.array([[1.0, 0.0, 0.0], [0.0, _E, -1.0], [0.0, 1.0, _E]], np.float64)
LIFT = np.array([0.0, 0.0, 1.0], np.float64)


def _layout(cloud, point_step, offsets):
    """-> (contiguous buffer array, n_in, point_step, (ox, oy, oz))"""
    if isinstance(cloud, (bytes, bytearray, memoryview)):
        if not point_step:
            raise ValueError("raw PointCloud2 bytes need point_step")
        buf = np.frombuffer(cloud, np.uint8)
        return buf, buf.shape[0] // point_step, int(point_step), tuple(offsets or (0, 4, 8))
    a = np.asarray(cloud)
    if a.dtype.fields is not None:  # structured PointCloud2 records
        for f in "xyz":
            if a.dtype.fields[f][0] != np.float32:
                raise ValueError("PointCloud2 x, y, z must be float32 fields")
        a = np.ascontiguousarray(a).reshape(-1)
        return a, a.shape[0], a.dtype.itemsize, tuple(int(a.dtype.fields[f][1]) for f in "xyz")
    if a.ndim != 2 or a.shape[1] < 3:
        raise ValueError("cloud must be [N,>=3] float32, a structured PointCloud2 array, or raw bytes")
    a = np.ascontiguousarray(a, np.float32)
    return a, a.shape[0], 4 * a.shape[1], (0, 4, 8)


def pointcloud2_to_lidar(cloud, rotations=(R_Y_NEG90, R_X_POS90), translation=LIFT, start=1, step=4,
                         point_step=None, offsets=None, device=None):
    """load_data.py:2434-2443 -> float64 [M,3]."""
    buf, n_in, ps, (ox, oy, oz) = _layout(cloud, point_step, offsets)
    rot = np.ascontiguousarray(np.stack([np.asarray(r, np.float64).reshape(3, 3) for r in rotations])
                               if len(rotations) else np.zeros((0, 3, 3)), np.float64)
    tr = None if translation is None else np.ascontiguousarray(translation, np.float64).reshape(3)
    cap = max(0, (n_in - start + step - 1) // step)
    out = np.empty((cap, 3), np.float64)
    n = C.c_int32(0)
    c = _lib.ctx(device)
    _lib.check(_lib.lib().pp_ingest_host(c.handle, _lib.ptr(buf), n_in, ps, ox, oy, oz, int(start), int(step),
                                         _lib.ptr(rot), rot.shape[0], None if tr is None else _lib.ptr(tr),
                                         _lib.ptr(out), cap, C.byref(n)))
    return out[:n.value]
