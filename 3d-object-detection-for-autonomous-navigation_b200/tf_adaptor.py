"""TensorFlow side of the drop-in (last bullet): the two Keras layers of the reference that sit on
the hot path keep their `call` signatures and hand their tensors to the device entry points over DLPack, without a
host round trip.

    model/voxelnet.py:867   voxel_features = self.voxel_feature_extractor(voxels, num_points, coors)
                            -> the decoration half of PillarFeatureNet.call (model/pointpillars.py:143-203)
    model/voxelnet.py:881   spatial_features = self.middle_feature_extractor(voxel_features, coors)
                            -> PointPillarsScatter.call (model/pointpillars.py:285-341)

TensorFlow is not part of this image, so the module never imports it at load time: `tf` is looked up on first use
(or injected with `use_tf_module`, which is how the tests stand torch's DLPack in for TensorFlow's).  Inside a
`tf.function` the calls are wrapped with `tf.py_function`, as any eager-only op is.
"""
from __future__ import annotations

from . import interop as _interop

_tf = None


def use_tf_module(module):
    """Inject the module that provides `experimental.dlpack.to_dlpack / from_dlpack` (tensorflow itself by default)."""
    global _tf
    _tf = module


def _tfm():
    global _tf
    if _tf is None:
        try:
            import tensorflow as tf  # noqa: PLC0415
        except ImportError as e:  # pragma: no cover - TensorFlow is absent in this image
            raise ImportError("tf_adaptor needs TensorFlow (or use_tf_module(...)): " + str(e)) from e
        _tf = tf
    return _tf


def from_tf(x):
    """tf.Tensor on the GPU -> torch view of the same memory (no copy)."""
    return _interop.as_device_tensor(_tfm().experimental.dlpack.to_dlpack(x))


def to_tf(t):
    """torch CUDA tensor -> tf.Tensor over the same memory (no copy)."""
    import torch  # noqa: PLC0415
    return _tfm().experimental.dlpack.from_dlpack(torch.utils.dlpack.to_dlpack(t))


def pillar_decorate(voxels, num_points, coors, vx, vy, x_offset, y_offset):
    """Lines 143-203 of PillarFeatureNet.call on tf tensors: [M,P,D] f32, [M] i32, [M,4] i32 -> tf [M,P,D+5] f32 (what
    the layer hands to its pfn_layers)."""
    return to_tf(_interop.pillar_decorate(from_tf(voxels), from_tf(num_points), from_tf(coors), vx, vy, x_offset, y_offset))


class PointPillarsScatter:
    """Same constructor `(config, training)` -- batch_size / ny / nx / nchannels fixed from the config as at
    model/pointpillars.py:254-275 -- and the same `call(voxel_features, coords)` -> [batch_size, nchannels, ny, nx]
    float32 (NCHW) as the reference layer, on tf tensors."""

    def __init__(self, config, training=False, layout="NCHW"):
        from .pillars import PointPillarsScatter as _HostLayer  # noqa: PLC0415  (the config parsing lives there)
        h = _HostLayer(config, training, layout)
        self.batch_size, self.ny, self.nx, self.nchannels, self.layout = h.batch_size, h.ny, h.nx, h.nchannels, layout

    def call(self, voxel_features, coords):
        return to_tf(_interop.scatter(from_tf(voxel_features), from_tf(coords), self.batch_size, self.ny, self.nx, self.layout))

    __call__ = call
