"""Drop-in for ``ctvae/forward_functions.py`` of vganapati/CT_PVAE.

Same function names, positional arguments, layouts and angle conventions as the
reference (``/root/reference/ctvae/forward_functions.py``):

  pad_phantom(phantom, dim=3, integrate_vae=False)                       :18-46
  project_tf_low_mem(phantom, theta, pad=False)                          :49-78
  project_tf_fast(phantom, theta, pad=False, dim=3, integrate_vae=False) :80-123

but the rotate + row-sum graph (pad -> transpose -> repeat -> tfa.image.rotate ->
reduce_sum -> transpose) is ONE fused ray-driven kernel on the B200, and the gradient
is the matched gather adjoint.  Inputs may be torch tensors (CUDA: zero-copy; CPU:
staged through the GPU) or NumPy arrays; the result comes back in the same kind,
device and dtype.  Keyword-only extras keep the reference's defaults:

  interpolation  "nearest" for project_tf_fast (tfa.image.rotate's default, :113),
                 "bilinear" for project_tf_low_mem (:70-74)
  adjoint        "exact" (true transpose, <Ax,y> == <x,A^T y>) or "tf_compat"
                 (TensorFlow's registered gradient of the reference graph)
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, hostpipe, ops

__all__ = ["pad_phantom", "project_tf_low_mem", "project_tf_fast", "backproject", "num_proj_pix"]


def num_proj_pix(img_size_x: int, img_size_y: int) -> int:
    """Detector width pad_phantom pads to (forward_functions.py:29-30)."""
    return int(_lib.lib().ctr_num_proj_pix(int(img_size_x), int(img_size_y)))


def _as_tensor(x):
    """-> (torch tensor, was_numpy)"""
    if isinstance(x, torch.Tensor):
        return x, False
    arr = np.asarray(x)
    if arr.dtype == np.float16 or not np.issubdtype(arr.dtype, np.floating):
        if np.issubdtype(arr.dtype, np.integer) or arr.dtype == np.bool_:
            arr = arr.astype(np.float32)
        else:
            raise TypeError(f"unsupported image dtype {arr.dtype}")
    return torch.from_numpy(np.ascontiguousarray(arr)), True


def _compute_device(t: torch.Tensor) -> torch.device:
    if t.is_cuda:
        return t.device
    if not torch.cuda.is_available():
        raise RuntimeError("ct_pvae_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _finish(out: torch.Tensor, like: torch.Tensor, was_numpy: bool):
    dtype = like.dtype if like.dtype.is_floating_point else torch.float32
    if out.is_cuda and like.device.type == "cpu" and like.is_pinned() and not out.requires_grad:
        # pinned in -> pinned out: the device->host copy runs at full PCIe rate
        host = torch.empty(out.shape, dtype=dtype, pin_memory=True)
        host.copy_(out.to(dtype), non_blocking=True)
        torch.cuda.current_stream(out.device).synchronize()
        return host
    out = out.to(device=like.device, dtype=dtype)
    return out.detach().numpy() if was_numpy else out


def pad_phantom(phantom, dim=3, integrate_vae=False):
    """Zero-pad rows and columns to the diagonal detector size (reference :18-46).
    Pure layout; the projector itself never materialises this copy."""
    t, was_numpy = _as_tensor(phantom)
    if integrate_vae:
        X, Y = t.shape[1], t.shape[2]
    else:
        X, Y = t.shape[0], t.shape[1]
    P = num_proj_pix(X, Y)
    padx, pady = (P - X) // 2, (P - Y) // 2
    odd_x, odd_y = (P - X) % 2, (P - Y) % 2
    if integrate_vae:
        pads = (0, 0, pady, pady + odd_y, padx, padx + odd_x, 0, 0)
    elif dim == 3:
        pads = (0, 0, pady, pady + odd_y, padx, padx + odd_x)
    elif dim == 2:
        pads = (pady, pady + odd_y, padx, padx + odd_x)
    else:
        raise ValueError("dim must be 2 or 3")
    out = torch.nn.functional.pad(t, pads, mode="constant", value=0)
    return out.numpy() if was_numpy else out


def _project_bxy(img_bxy: torch.Tensor, theta, pad: bool, interpolation: str, adjoint: str, async_op: bool = False, out=None):
    """[B,X,Y] (any float dtype, any device) -> [B,A,W] float32 on the compute device."""
    if interpolation not in ops.INTERP:
        raise ValueError(f"interpolation must be 'nearest' or 'bilinear', got {interpolation!r}")
    if adjoint not in ops.ADJOINT:
        raise ValueError(f"adjoint must be 'exact' or 'tf_compat', got {adjoint!r}")
    dev = _compute_device(img_bxy)
    th = ops.theta_to_host(theta)
    plan = _lib.get_plan(th, int(img_bxy.shape[1]), int(img_bxy.shape[2]), bool(pad), dev.index or 0)
    if hostpipe.eligible(img_bxy):
        # host (pinned) batch: overlap copy-in / kernels / copy-out chunk by chunk; result stays on the host
        iid = ops.INTERP[interpolation]
        return hostpipe.forward_host(plan, img_bxy, iid, async_op=async_op, out=out)
    if async_op or out is not None:
        raise ValueError("async_op / out need a contiguous float32 host batch of >= 32 images")
    x = img_bxy.to(device=dev, dtype=torch.float32, non_blocking=True)
    return ops.project(x, plan, ops.INTERP[interpolation], ops.ADJOINT[adjoint])


def project_tf_fast(phantom, theta, pad=False, dim=3, integrate_vae=False, *, interpolation="nearest",
                    adjoint="exact", async_op=False, out=None):
    """Parallel-beam Radon transform of every image / channel (reference :80-123).

    phantom is ``[X,Y,Z]`` (dim=3), ``[X,Y]`` (dim=2) or, with integrate_vae,
    ``[B,X,Y,1]``.  Returns ``[A,P,Z]``, ``[A,P,1]`` or ``[B,A,P,1]``: bin ``j`` of angle
    ``a`` is the sum over rows of the image rotated by ``-theta[a]`` about its centre.

    Host batches (``integrate_vae`` with >= 32 float32 images in a CPU tensor or NumPy array) run through the
    library's chunked copy / compute pipeline.  ``async_op=True`` then returns ``(result, handle)``: the page-locked
    result may be read after ``handle.wait()``; ``out=`` (a float32 CPU tensor ``[B,A,P]``, ideally pinned) receives
    the result instead of a new buffer.
    """
    t, was_numpy = _as_tensor(phantom)
    if not t.dtype.is_floating_point:
        raise TypeError(f"unsupported image dtype {t.dtype}")
    len(theta)  # the reference calls len(theta): scalars are not accepted (:90)
    if integrate_vae:
        if t.dim() != 4 or t.shape[3] != 1:
            raise ValueError("integrate_vae expects [batch, x, y, 1]")
        if async_op:
            # host batch: returns (result, handle); the result may be read after handle.wait()
            sino, handle = _project_bxy(t[..., 0], theta, pad, interpolation, adjoint, async_op=True, out=out)
            sino = sino.unsqueeze(-1)
            return (sino.numpy() if was_numpy else sino), handle
        sino = _project_bxy(t[..., 0], theta, pad, interpolation, adjoint, out=out)      # [B,A,W]
        out = sino.unsqueeze(-1)
    else:
        if async_op:
            raise ValueError("async_op=True is only available with integrate_vae=True (batched host input)")
        if dim == 2:
            if t.dim() != 2:
                raise ValueError("dim=2 expects [x, y]")
            t3 = t.unsqueeze(-1)
        elif dim == 3:
            if t.dim() != 3:
                raise ValueError("dim=3 expects [x, y, z]")
            t3 = t
        else:
            raise ValueError("dim must be 2 or 3")
        sino = _project_bxy(t3.permute(2, 0, 1), theta, pad, interpolation, adjoint)  # [Z,A,W]
        out = sino.permute(1, 2, 0)
    return _finish(out, t, was_numpy)


def project_tf_low_mem(phantom, theta, pad=False, *, interpolation="bilinear", adjoint="exact"):
    """Per-angle variant of the reference (:49-78): ``[X,Y,Z] -> [A,P,Z]``, bilinear."""
    t, was_numpy = _as_tensor(phantom)
    if t.dim() != 3:
        raise ValueError("project_tf_low_mem expects [x, y, z]")
    len(theta)
    sino = _project_bxy(t.permute(2, 0, 1), theta, pad, interpolation, adjoint)
    return _finish(sino.permute(1, 2, 0), t, was_numpy)


def backproject(sinogram, theta, x_size, y_size, pad=False, *, interpolation="nearest", adjoint="exact",
                async_op=False, out=None):
    """Adjoint of ``project_tf_fast(..., integrate_vae=True)`` as a function:
    ``[B,A,P,1]`` (or ``[B,A,P]``) -> ``[B,x_size,y_size,1]`` (or ``[B,x_size,y_size]``).
    This is what autograd calls; exposed for matched iterative solvers and tests."""
    t, was_numpy = _as_tensor(sinogram)
    squeeze = t.dim() == 4
    s3 = t[..., 0] if squeeze else t
    dev = _compute_device(s3)
    th = ops.theta_to_host(theta)
    plan = _lib.get_plan(th, int(x_size), int(y_size), bool(pad), dev.index or 0)
    if hostpipe.eligible(s3):
        iid, mid = ops.INTERP[interpolation], ops.ADJOINT[adjoint]
        if async_op:
            g, handle = hostpipe.adjoint_host(plan, s3, iid, mid, async_op=True, out=out)
            g = g.unsqueeze(-1) if squeeze else g
            return (g.numpy() if was_numpy else g), handle
        g = hostpipe.adjoint_host(plan, s3, iid, mid, out=out)
        return _finish(g.unsqueeze(-1) if squeeze else g, t, was_numpy)
    if async_op or out is not None:
        raise ValueError("async_op / out need a contiguous float32 host batch of >= 32 sinograms")
    y = s3.to(device=dev, dtype=torch.float32)
    g = ops.radon_adjoint(y, plan, ops.INTERP[interpolation], ops.ADJOINT[adjoint])
    return _finish(g.unsqueeze(-1) if squeeze else g, t, was_numpy)
