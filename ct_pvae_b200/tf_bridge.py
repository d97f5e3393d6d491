"""TensorFlow binding of libctradon (north_star: "wrapped in a tf.custom_gradient, so the
forward call dispatches the projector and the gradient dispatches the adjoint").

EXPERIMENTAL: TensorFlow is not installable in this image, so this module has never run under a real
TensorFlow.  What does run (tests/test_tf_bridge.py, on the GPU) is this exact code against a small torch-backed
stand-in for the handful of ``tf`` calls it makes; the torch binding (ops.py) is the one under full test.

How it is wired (and why):
* ``tf.custom_gradient`` wraps the op; forward and gradient each run inside ``tf.py_function``, so the body sees
  EAGER tensors even when the caller is a ``@tf.function`` graph (the reference's ``train_step``,
  main_ct_vae.py:463, where ``theta`` is the symbolic ``tf.gather(theta, angles_i)``): ``to_dlpack`` and the plan
  lookup need concrete values.
* Tensors are exchanged zero-copy through ``tf.experimental.dlpack`` and the ``*_dl`` entry points of the C ABI.
* Streams: TensorFlow's GPU compute stream is a private non-blocking stream that Python cannot name, so it is NOT
  ordered with the stream the kernels are launched on.  The binding therefore synchronises the device after the
  inputs / outputs exist and again after the kernels (``ctr_device_synchronize``): correct, at the price of two
  device syncs per call.  A C++ custom op that receives ``ctx->eigen_gpu_device().stream()`` would remove them.
* The device comes from the input tensor; plans come from the package's plan cache (nothing is leaked per call).
"""
from __future__ import annotations

import ctypes
import re

import numpy as np

from . import _lib

try:  # pragma: no cover - TensorFlow is absent in the build image
    import tensorflow as tf
except Exception:  # noqa: BLE001
    tf = None

_cap = ctypes.pythonapi.PyCapsule_GetPointer
_cap.restype, _cap.argtypes = ctypes.c_void_p, [ctypes.py_object, ctypes.c_char_p]


def available() -> bool:
    return tf is not None


def _dl(t):
    """Borrowed DLTensor* of an eager tensor; the capsule (kept alive by the caller) owns the export."""
    cap = tf.experimental.dlpack.to_dlpack(t)
    return cap, ctypes.c_void_p(_cap(cap, b"dltensor"))


def _device_index(t) -> int:
    m = re.search(r"(\d+)$", str(t.device))
    return int(m.group(1)) if m else 0


def _theta64(theta) -> np.ndarray:
    th = theta.numpy() if hasattr(theta, "numpy") else np.asarray(theta)
    return np.ascontiguousarray(np.asarray(th, np.float64).reshape(-1))


def _forward_eager(img, theta, pad, iid):
    """img [B,X,Y] float32 eager GPU tensor -> sino [B,A,W]."""
    dev = _device_index(img)
    B, X, Y = (int(v) for v in img.shape)
    plan = _lib.get_plan(_theta64(theta), X, Y, bool(pad), dev)
    with tf.device(img.device):
        img = tf.identity(tf.cast(img, tf.float32))
        sino = tf.zeros([B, plan.A, plan.W], tf.float32)
        ws = tf.zeros([max(plan.forward_workspace_bytes(B), 256)], tf.uint8)
    (c1, a), (c2, b), (c3, w) = _dl(img), _dl(sino), _dl(ws)
    L = _lib.lib()
    _lib.check(L.ctr_device_synchronize(dev))          # TF's stream has produced img and zero-filled sino / ws
    _lib.check(L.ctr_radon_forward_dl(plan.handle, a, b, iid, w, None))
    _lib.check(L.ctr_device_synchronize(dev))          # ... and the result is complete before TF reads it
    del c1, c2, c3
    return sino


def _adjoint_eager(dsino, theta, X, Y, pad, iid, mid):
    dev = _device_index(dsino)
    B = int(dsino.shape[0])
    plan = _lib.get_plan(_theta64(theta), int(X), int(Y), bool(pad), dev)
    with tf.device(dsino.device):
        dsino = tf.identity(tf.cast(dsino, tf.float32))
        dimg = tf.zeros([B, plan.X, plan.Y], tf.float32)
        ws = tf.zeros([max(plan.adjoint_workspace_bytes(B), 256)], tf.uint8)
    (c1, p), (c2, q), (c3, r) = _dl(dsino), _dl(dimg), _dl(ws)
    L = _lib.lib()
    _lib.check(L.ctr_device_synchronize(dev))
    _lib.check(L.ctr_radon_adjoint_dl(plan.handle, p, q, iid, mid, r, None))
    _lib.check(L.ctr_device_synchronize(dev))
    del c1, c2, c3
    return dimg


def project_tf_fast(phantom, theta, pad=False, dim=3, integrate_vae=False, *, interpolation="nearest", adjoint="exact"):
    """Same signature and layouts as the reference (ctvae/forward_functions.py:80-123); tensors stay on the TF GPU
    device; differentiable with respect to ``phantom`` (the gradient is the adjoint kernel)."""
    if tf is None:
        raise RuntimeError("TensorFlow is not installed; use ct_pvae_b200.forward_functions (torch / NumPy) instead")
    if interpolation not in ("nearest", "bilinear") or adjoint not in ("exact", "tf_compat"):
        raise ValueError("interpolation must be 'nearest' or 'bilinear', adjoint 'exact' or 'tf_compat'")
    iid = _lib.INTERP_NEAREST if interpolation == "nearest" else _lib.INTERP_BILINEAR
    mid = _lib.ADJOINT_EXACT if adjoint == "exact" else _lib.ADJOINT_TF_COMPAT
    len(theta)                                   # the reference calls len(theta) (:90)
    x = tf.convert_to_tensor(phantom)
    if integrate_vae:
        bxy = x[..., 0]
    else:
        if dim == 2:
            x = x[..., None]
        bxy = tf.transpose(x, [2, 0, 1])
    X, Y = int(bxy.shape[1]), int(bxy.shape[2])
    theta_t = tf.convert_to_tensor(theta)

    @tf.custom_gradient
    def op(img):
        sino = tf.py_function(lambda i, t: _forward_eager(i, t, pad, iid), [img, theta_t], tf.float32)

        def grad(dsino):
            return tf.py_function(lambda d, t: _adjoint_eager(d, t, X, Y, pad, iid, mid), [dsino, theta_t], tf.float32)

        return sino, grad

    s = op(bxy)
    return s[..., None] if integrate_vae else tf.transpose(s, [1, 2, 0])
