"""TensorFlow binding of libctradon (north_star: "wrapped in a tf.custom_gradient, so the
forward call dispatches the projector and the gradient dispatches the adjoint").

TensorFlow is NOT installable in this image, so this module is import-guarded and is
exercised only where TF exists; the torch binding (ops.py) is the one under test here.
It uses the same C ABI with tf.experimental.dlpack for zero-copy tensor exchange and keeps
the reference signature of ``project_tf_fast`` (ctvae/forward_functions.py:80-123).
"""
from __future__ import annotations

import ctypes

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


def _dl(t):  # pragma: no cover
    cap = tf.experimental.dlpack.to_dlpack(t)
    return cap, ctypes.c_void_p(_cap(cap, b"dltensor"))


def project_tf_fast(phantom, theta, pad=False, dim=3, integrate_vae=False, *, interpolation="nearest", adjoint="exact"):  # pragma: no cover
    """Same signature and layouts as the reference; tensors stay on the TF GPU device."""
    if tf is None:
        raise RuntimeError("TensorFlow is not installed; use ct_pvae_b200.forward_functions (torch / NumPy) instead")
    iid = _lib.INTERP_NEAREST if interpolation == "nearest" else _lib.INTERP_BILINEAR
    mid = _lib.ADJOINT_EXACT if adjoint == "exact" else _lib.ADJOINT_TF_COMPAT
    th = np.ascontiguousarray(np.asarray(theta, np.float64).reshape(-1))
    x = tf.convert_to_tensor(phantom)
    if integrate_vae:
        bxy = x[..., 0]
    else:
        if dim == 2:
            x = x[..., None]
        bxy = tf.transpose(x, [2, 0, 1])
    B, X, Y = int(bxy.shape[0]), int(bxy.shape[1]), int(bxy.shape[2])
    plan = _lib.get_plan(th, X, Y, bool(pad), 0)
    L = _lib.lib()

    @tf.custom_gradient
    def op(img):
        img = tf.identity(tf.cast(img, tf.float32))
        sino = tf.zeros([B, plan.A, plan.W], tf.float32)
        ws = tf.zeros([max(plan.forward_workspace_bytes(B), 256)], tf.uint8)
        (c1, a), (c2, b), (c3, w) = _dl(img), _dl(sino), _dl(ws)
        _lib.check(L.ctr_radon_forward_dl(plan.handle, a, b, iid, w, None))

        def grad(dsino):
            dsino = tf.identity(tf.cast(dsino, tf.float32))
            dimg = tf.zeros([B, X, Y], tf.float32)
            ws2 = tf.zeros([max(plan.adjoint_workspace_bytes(B), 256)], tf.uint8)
            (d1, p), (d2, q), (d3, r) = _dl(dsino), _dl(dimg), _dl(ws2)
            _lib.check(L.ctr_radon_adjoint_dl(plan.handle, p, q, iid, mid, r, None))
            return dimg

        return sino, grad

    s = op(bxy)
    return s[..., None] if integrate_vae else tf.transpose(s, [1, 2, 0])
