"""NumPy stand-in for the handful of TensorFlow calls made by the reference's hot-path files.
TEST INFRASTRUCTURE ONLY (see ../README.md).  Tensors are plain numpy arrays."""
import types

import numpy as np

float32, float64, int32, int64, complex128, complex64 = (np.float32, np.float64, np.int32, np.int64, np.complex128,
                                                         np.complex64)
__version__ = "2.8.1-numpy-shim"


def convert_to_tensor(x, dtype=None, name=None):
    return np.asarray(x, dtype=dtype)


def cast(x, dtype):
    return np.asarray(x).astype(dtype)


def sqrt(x):
    return np.sqrt(x)


def pad(tensor, paddings, mode="CONSTANT", constant_values=0):
    assert mode.upper() == "CONSTANT"
    return np.pad(np.asarray(tensor), [(int(a), int(b)) for a, b in paddings], mode="constant",
                  constant_values=constant_values)


def transpose(a, perm=None):
    return np.transpose(np.asarray(a), perm)


def repeat(x, repeats, axis=None):
    return np.repeat(np.asarray(x), repeats, axis=axis)


def expand_dims(x, axis):
    return np.expand_dims(np.asarray(x), axis)


def squeeze(x, axis=None):
    return np.squeeze(np.asarray(x), axis=axis)


def stack(values, axis=0):
    return np.stack([np.asarray(v) for v in values], axis=axis)


def range(start, limit=None, delta=1, dtype=None):  # noqa: A001  (tf.range)
    if limit is None:
        start, limit = 0, start
    return np.arange(start, limit, delta, dtype=dtype)


def meshgrid(*args, indexing="xy"):
    return np.meshgrid(*args, indexing=indexing)


def reduce_sum(x, axis=None):
    return np.sum(np.asarray(x), axis=axis)


def reduce_min(x, axis=None):
    return np.min(np.asarray(x), axis=axis)


def reduce_max(x, axis=None):
    return np.max(np.asarray(x), axis=axis)


math = types.SimpleNamespace(ceil=np.ceil, floor=np.floor, cos=np.cos, sin=np.sin, real=np.real, sqrt=np.sqrt,
                             reduce_sum=reduce_sum, reduce_min=reduce_min, reduce_max=reduce_max)
signal = types.SimpleNamespace(fft=lambda x: np.fft.fft(x, axis=-1), ifft=lambda x: np.fft.ifft(x, axis=-1))
