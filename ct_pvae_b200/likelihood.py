"""Measurement log-likelihood on top of the projector.

Mirrors ``calculate_log_prob_M_given_R`` of ``ctvae/helper_functions.py:336-368``:

    theta, mask, proj_sample are gathered at ``angles_i``                    (:350-357)
    proj        = project_tf_fast(output_sample, theta, pad, dim=2, integrate_vae=True)   (:359)
    proj_masked = proj * mask[:, :, None, None]                               (:360)
    Normal(proj_masked, sqrt_reg + sqrt(proj_masked / pnm + sqrt_reg)).log_prob(proj_sample)  (:364-368)

``calculate_log_prob_M_given_R`` keeps the reference signature and returns the full
``[B,A',P,1]`` tensor.  ``log_prob_M_given_R_sum`` is the fused form of what the loss
actually consumes (``reduce_sum`` over all axes, :305-311): the projector kernel's
epilogue evaluates the log-probability and its derivative per ray, so the sinogram is
never written and the backward pass is a single adjoint launch.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib, ops
from .forward_functions import _compute_device, project_tf_fast

__all__ = ["calculate_log_prob_M_given_R", "log_prob_M_given_R_sum"]


def _gather_theta(theta, angles_i):
    th = ops.theta_to_host(theta)
    if angles_i is None:
        return th, None
    idx = np.asarray(angles_i.detach().cpu() if isinstance(angles_i, torch.Tensor) else angles_i).astype(np.int64).reshape(-1)
    # the reference casts the gathered angles to float32 before the projector sees them (:355)
    return th[idx].astype(np.float32).astype(np.float64), idx


def calculate_log_prob_M_given_R(output_sample, mask, proj_sample, poisson_noise_multiplier, sqrt_reg, theta=None,
                                 angles_i=None, pad=True, *, interpolation="nearest", adjoint="exact"):
    """Drop-in (reference :336-368): returns log p(M | R) per ray, ``[B,A',P,1]``."""
    th, idx = _gather_theta(theta, angles_i)
    if idx is None or not (isinstance(output_sample, torch.Tensor) and output_sample.is_cuda):
        if idx is not None:
            sel = torch.as_tensor(idx, device=mask.device)
            mask = mask.index_select(1, sel)
            proj_sample = proj_sample.index_select(1, sel.to(proj_sample.device))
        proj = project_tf_fast(output_sample, th, pad=pad, dim=2, integrate_vae=True, interpolation=interpolation, adjoint=adjoint)
    else:
        # device tensors + an angle minibatch: one plan over all the angles, the subset as an index list
        dev = output_sample.device
        th_all = ops.theta_to_host(theta).astype(np.float32).astype(np.float64)
        plan = _lib.get_plan(th_all, int(output_sample.shape[1]), int(output_sample.shape[2]), bool(pad), dev.index or 0)
        sel = torch.as_tensor(idx, dtype=torch.int32, device=dev)
        mask = mask.to(dev).index_select(1, sel.long())
        proj_sample = proj_sample.to(dev).index_select(1, sel.long())
        proj = ops.project(output_sample[..., 0].to(torch.float32), plan, ops.INTERP[interpolation], ops.ADJOINT[adjoint], sel).unsqueeze(-1)
    pm = proj * mask.to(proj.device)[:, :, None, None]
    scale = sqrt_reg + torch.sqrt(pm / poisson_noise_multiplier + sqrt_reg)
    y = proj_sample.to(proj.device).unsqueeze(-1)
    return -0.5 * ((y - pm) / scale) ** 2 - torch.log(scale) - 0.5 * math.log(2 * math.pi)


def log_prob_M_given_R_sum(output_sample, mask, proj_sample, poisson_noise_multiplier, sqrt_reg, theta=None,
                           angles_i=None, pad=True, *, interpolation="nearest", adjoint="exact", per_image=False):
    """Fused ``reduce_sum(calculate_log_prob_M_given_R(...))`` (scalar, or ``[B]`` with
    per_image=True), differentiable with respect to ``output_sample`` ``[B,X,Y,1]``."""
    if output_sample.dim() != 4 or output_sample.shape[3] != 1:
        raise ValueError("output_sample must be [batch, x, y, 1]")
    dev = _compute_device(output_sample)
    # ONE plan over all the angles serves every angle minibatch: the subset travels as an index list (no plan is
    # created, uploaded or freed on the iteration path).  The reference casts the gathered angles to float32 (:355);
    # the plan's transform table is built from float32(-theta) either way, so the table rows are the same.
    th_all = ops.theta_to_host(theta)
    if angles_i is not None:
        th_all = th_all.astype(np.float32).astype(np.float64)
    plan = _lib.get_plan(th_all, int(output_sample.shape[1]), int(output_sample.shape[2]), bool(pad), dev.index or 0)
    img = output_sample[..., 0].to(device=dev, dtype=torch.float32)
    mask_d = mask.to(device=dev, dtype=torch.float32)
    meas_d = proj_sample.to(device=dev, dtype=torch.float32)
    sel = None
    if angles_i is not None:
        sel = torch.as_tensor(angles_i).reshape(-1).to(device=dev, dtype=torch.int32)
    iid, mid = ops.INTERP[interpolation], ops.ADJOINT[adjoint]
    if img.requires_grad and torch.is_grad_enabled():
        ll = ops.LoglikFunction.apply(img, plan, mask_d, meas_d, None, float(poisson_noise_multiplier), float(sqrt_reg), iid, mid, sel)
    else:
        ll, _ = ops.radon_loglik(img, plan, mask_d, meas_d, None, float(poisson_noise_multiplier), float(sqrt_reg), iid, sel)
    return ll if per_image else ll.sum()
