"""numpy stand-in for the handful of TensorFlow ops the reference's decoration and scatter call.

TEST INFRASTRUCTURE (build container only).  TensorFlow 2.2 cannot be installed here, so
`model/pointpillars.py:PillarFeatureNet.call` (143-203) and `PointPillarsScatter.call` (285-341) cannot run
as shipped.  To pin the oracle to the reference's OWN op sequence -- which columns, which constants, the order of
concat / mask / scatter, duplicate handling -- `oracle/ref_extract.py` executes those two method bodies unmodified
with this module bound to the name `tf`.  Semantics mirrored from TensorFlow:
  * python scalars / lists combined with a tensor take the tensor's dtype (`c * self.vx`, `x @ [[1, 0]]`);
  * `tf.zeros` defaults to float32; `tf.scatter_nd` ADDS duplicate indices; `tf.boolean_mask` keeps row order.
What this cannot pin is TensorFlow's internal float32 summation order inside `reduce_sum` (numpy adds the P slots
in slot order) -- the 1e-5 relative tolerance of the parity tests covers it.
"""
import numpy as np

float32, float64, int32, int64 = np.float32, np.float64, np.int32, np.int64


class Tensor(np.ndarray):
    """ndarray whose binary ops convert python lists to its own dtype, like tf.Tensor.__matmul__ does."""

    def __matmul__(self, other):
        if not isinstance(other, np.ndarray):
            other = np.asarray(other, dtype=self.dtype)
        return np.matmul(np.asarray(self), np.asarray(other)).view(Tensor)


def _t(a):
    return np.asarray(a).view(Tensor)


def convert(a):
    return _t(np.ascontiguousarray(a))


def reduce_sum(x, axis=None, keepdims=False):
    return _t(np.sum(np.asarray(x), axis=axis, keepdims=keepdims, dtype=np.asarray(x).dtype))


def reshape(x, shape):
    return _t(np.reshape(np.asarray(x), shape))


def cast(x, dtype):
    return _t(np.asarray(x).astype(dtype))


def shape(x):
    return np.array(np.shape(x), np.int32)


def zeros(shp, dtype=np.float32):
    return _t(np.zeros(tuple(int(v) for v in np.atleast_1d(shp)), dtype))


def expand_dims(x, axis):
    return _t(np.expand_dims(np.asarray(x), axis))


def concat(xs, axis):
    return _t(np.concatenate([np.asarray(x) for x in xs], axis=axis))


def range(n, dtype=np.int32):  # noqa: A001  (the reference calls tf.range)
    return _t(np.arange(int(n), dtype=dtype))


def boolean_mask(x, mask):
    return _t(np.asarray(x)[np.asarray(mask)])


def transpose(x, perm=None):
    return _t(np.transpose(np.asarray(x), perm))


def constant(v, dtype=None):
    return _t(np.asarray(v, dtype=dtype))


def scatter_nd(indices, updates, shp):
    out = np.zeros(tuple(int(v) for v in np.asarray(shp)), np.asarray(updates).dtype)
    np.add.at(out, tuple(np.asarray(indices).T), np.asarray(updates))  # duplicates accumulate, in row order
    return _t(out)


def stack(xs, axis=0):
    return _t(np.stack([np.asarray(x) for x in xs], axis=axis))


def squeeze(x):
    return _t(np.squeeze(np.asarray(x)))


class math:  # noqa: N801
    @staticmethod
    def reduce_max(x, axis=None, keepdims=False):
        return _t(np.max(np.asarray(x), axis=axis, keepdims=keepdims))
