"""Device-level operators: torch CUDA tensors in, torch CUDA tensors out.

torch is plumbing here (device memory, the current stream, autograd bookkeeping);
every FLOP runs in libctradon's sm_100a kernels, reached through the C ABI with
zero-copy DLPack tensors.  CPU tensors are rejected -- there is no fallback.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

INTERP = {"nearest": _lib.INTERP_NEAREST, "bilinear": _lib.INTERP_BILINEAR}
ADJOINT = {"exact": _lib.ADJOINT_EXACT, "tf_compat": _lib.ADJOINT_TF_COMPAT}


def _require_cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: ct_pvae_b200 has no CPU path")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


def _stream_ptr(device: torch.device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


def _workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    # caller-owned scratch from torch's caching allocator (512-byte aligned blocks)
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def theta_to_host(theta) -> np.ndarray:
    """Angles as the float64 host vector the plan is keyed on.  float32 inputs (the
    training path casts theta to float32, helper_functions.py:355) widen exactly."""
    if isinstance(theta, torch.Tensor):
        theta = theta.detach().cpu().numpy()
    th = np.asarray(theta)
    if th.ndim > 1:
        raise ValueError("angles should have rank 0 or 1.")
    return np.ascontiguousarray(th.reshape(-1), dtype=np.float64)


def _check_sel(sel, plan: _lib.Plan, device) -> torch.Tensor:
    """Angle subset as the device int32 index list the *_sel entry points take."""
    if not isinstance(sel, torch.Tensor):
        sel = torch.as_tensor(np.asarray(sel).reshape(-1), dtype=torch.int32)
    sel = sel.to(device=device, dtype=torch.int32).contiguous().reshape(-1)
    if sel.numel() < 1 or sel.numel() > plan.A:
        raise ValueError("the angle subset must have between 1 and A entries")
    return sel


def radon_forward(img: torch.Tensor, plan: _lib.Plan, interp: int, sel=None) -> torch.Tensor:
    """img [B,X,Y] float32 CUDA -> sino [B,A,W] float32 CUDA (ctr_radon_forward_dl).
    sel: angle subset (indices into the plan's angles) -> sino [B,len(sel),W] (ctr_radon_forward_sel)."""
    img = _require_cuda_f32(img, "img")
    if img.dim() != 3 or img.shape[1] != plan.X or img.shape[2] != plan.Y:
        raise ValueError(f"img must be [B,{plan.X},{plan.Y}], got {tuple(img.shape)}")
    B = img.shape[0]
    if sel is not None:
        sel = _check_sel(sel, plan, img.device)
        sino = torch.empty((B, sel.numel(), plan.W), dtype=torch.float32, device=img.device)
        if B == 0:
            return sino
        ws = _workspace(plan.forward_workspace_bytes(B), img.device)
        _lib.check(_lib.lib().ctr_radon_forward_sel(plan.handle, img.data_ptr(), sino.data_ptr(), B, interp, sel.data_ptr(), sel.numel(),
                                                    ws.data_ptr(), ws.numel(), _stream_ptr(img.device)))
        return sino
    sino = torch.empty((B, plan.A, plan.W), dtype=torch.float32, device=img.device)
    if B == 0:
        return sino
    ws = _workspace(plan.forward_workspace_bytes(B), img.device)
    vi, vs, vw = _lib.DLView(img), _lib.DLView(sino), _lib.DLView(ws)
    _lib.check(_lib.lib().ctr_radon_forward_dl(plan.handle, vi.ptr, vs.ptr, interp, vw.ptr, _stream_ptr(img.device)))
    return sino


def radon_adjoint(dsino: torch.Tensor, plan: _lib.Plan, interp: int, mode: int, sel=None) -> torch.Tensor:
    """dsino [B,A,W] float32 CUDA -> dimg [B,X,Y] float32 CUDA (ctr_radon_adjoint_dl).
    sel: dsino is [B,len(sel),W], the rows of an angle subset (ctr_radon_adjoint_sel)."""
    if sel is not None:
        return radon_adjoint_scaled(dsino, plan, interp, mode, 1.0, sel)
    dsino = _require_cuda_f32(dsino, "dsino")
    if dsino.dim() != 3 or dsino.shape[1] != plan.A or dsino.shape[2] != plan.W:
        raise ValueError(f"dsino must be [B,{plan.A},{plan.W}], got {tuple(dsino.shape)}")
    B = dsino.shape[0]
    dimg = torch.empty((B, plan.X, plan.Y), dtype=torch.float32, device=dsino.device)
    if B == 0:
        return dimg
    ws = _workspace(plan.adjoint_workspace_bytes(B), dsino.device)
    vi, vo, vw = _lib.DLView(dsino), _lib.DLView(dimg), _lib.DLView(ws)
    _lib.check(_lib.lib().ctr_radon_adjoint_dl(plan.handle, vi.ptr, vo.ptr, interp, mode, vw.ptr, _stream_ptr(dsino.device)))
    return dimg


def fbp(sino: torch.Tensor, plan: _lib.FbpPlan) -> torch.Tensor:
    """sino [B,A,P] float32 CUDA -> recon [B,x_size,y_size] float32 CUDA (ctr_fbp_dl)."""
    sino = _require_cuda_f32(sino, "sinogram")
    B = sino.shape[0]
    out = torch.empty((B, plan.x_size, plan.y_size), dtype=torch.float32, device=sino.device)
    if B == 0:
        return out
    ws = _workspace(plan.workspace_bytes(B), sino.device)
    vi, vo, vw = _lib.DLView(sino), _lib.DLView(out), _lib.DLView(ws)
    _lib.check(_lib.lib().ctr_fbp_dl(plan.handle, vi.ptr, vo.ptr, vw.ptr, _stream_ptr(sino.device)))
    return out


class RadonFunction(torch.autograd.Function):
    """Differentiable projector: forward dispatches K1, backward dispatches the adjoint
    (exact transpose by default, TensorFlow's gradient with adjoint="tf_compat").
    Mirrors what tf.GradientTape does around project_tf_fast (main_ct_vae.py:471-481)."""

    @staticmethod
    def forward(ctx, img, plan, interp, mode, sel=None):
        ctx.plan, ctx.interp, ctx.mode, ctx.sel = plan, interp, mode, sel
        return radon_forward(img, plan, interp, sel)

    @staticmethod
    def backward(ctx, dsino):
        return radon_adjoint(dsino.contiguous(), ctx.plan, ctx.interp, ctx.mode, ctx.sel), None, None, None, None


def project(img: torch.Tensor, plan: _lib.Plan, interp: int, mode: int, sel=None) -> torch.Tensor:
    if sel is not None:
        sel = _check_sel(sel, plan, img.device)
    if img.requires_grad and torch.is_grad_enabled():
        return RadonFunction.apply(img, plan, interp, mode, sel)
    return radon_forward(img, plan, interp, sel)


def radon_adjoint_scaled(dsino: torch.Tensor, plan: _lib.Plan, interp: int, mode: int, scale: float, sel=None) -> torch.Tensor:
    """``scale * A^T dsino`` in one launch (ctr_radon_adjoint_scaled / ctr_radon_adjoint_sel)."""
    dsino = _require_cuda_f32(dsino, "dsino")
    B = dsino.shape[0]
    dimg = torch.empty((B, plan.X, plan.Y), dtype=torch.float32, device=dsino.device)
    if sel is not None:
        sel = _check_sel(sel, plan, dsino.device)
        if dsino.dim() != 3 or dsino.shape[1] != sel.numel() or dsino.shape[2] != plan.W:
            raise ValueError(f"dsino must be [B,{sel.numel()},{plan.W}], got {tuple(dsino.shape)}")
    elif dsino.dim() != 3 or dsino.shape[1] != plan.A or dsino.shape[2] != plan.W:
        raise ValueError(f"dsino must be [B,{plan.A},{plan.W}], got {tuple(dsino.shape)}")
    if B == 0:
        return dimg
    ws = _workspace(plan.adjoint_workspace_bytes(B), dsino.device)
    if sel is not None:
        _lib.check(_lib.lib().ctr_radon_adjoint_sel(plan.handle, dsino.data_ptr(), dimg.data_ptr(), B, interp, mode, float(scale),
                                                    sel.data_ptr(), sel.numel(), ws.data_ptr(), ws.numel(), _stream_ptr(dsino.device)))
        return dimg
    _lib.check(_lib.lib().ctr_radon_adjoint_scaled(plan.handle, dsino.data_ptr(), dimg.data_ptr(), B, interp, mode, float(scale),
                                                   ws.data_ptr(), ws.numel(), _stream_ptr(dsino.device)))
    return dimg


def radon_loglik(img: torch.Tensor, plan: _lib.Plan, mask: torch.Tensor, meas: torch.Tensor, angle_map, pnm: float,
                 sqrt_reg: float, interp: int, sel=None):
    """Fused projector + measurement log-likelihood (ctr_radon_loglik).

    img [B,X,Y], mask [B,A_all], meas [B,A_all,W] float32 CUDA; angle_map int32 CUDA [A] or None.
    Returns (loglik [B], dproj [B,A,W]) -- dproj is d loglik / d proj, the adjoint's cotangent.
    sel: the plan covers ALL A_all angles and the call projects the subset sel (ctr_radon_loglik_sel);
    dproj is then [B,len(sel),W]."""
    img = _require_cuda_f32(img, "img")
    mask = _require_cuda_f32(mask, "mask")
    meas = _require_cuda_f32(meas, "proj_sample")
    B, A_all = img.shape[0], mask.shape[1]
    if mask.shape[0] != B or meas.shape[0] != B or meas.shape[1] != A_all or meas.shape[2] != plan.W:
        raise ValueError("mask must be [B,A_all] and proj_sample [B,A_all,num_proj_pix]")
    if sel is not None:
        if A_all != plan.A:
            raise ValueError("with an angle subset the mask must cover exactly the plan's angles")
        sel = _check_sel(sel, plan, img.device)
        loglik = torch.empty((B,), dtype=torch.float32, device=img.device)
        dproj = torch.empty((B, sel.numel(), plan.W), dtype=torch.float32, device=img.device)
        ws = _workspace(plan.loglik_workspace_bytes(B), img.device)
        _lib.check(_lib.lib().ctr_radon_loglik_sel(plan.handle, img.data_ptr(), mask.data_ptr(), meas.data_ptr(), sel.data_ptr(),
                                                   sel.numel(), float(pnm), float(sqrt_reg), loglik.data_ptr(), dproj.data_ptr(), B,
                                                   interp, ws.data_ptr(), ws.numel(), _stream_ptr(img.device)))
        return loglik, dproj
    if angle_map is None:
        if A_all != plan.A:
            raise ValueError("without angles_i the mask must cover exactly the plan's angles")
        amap_ptr = None
    else:
        if angle_map.dtype != torch.int32 or not angle_map.is_cuda or angle_map.numel() != plan.A:
            raise ValueError("angle_map must be an int32 CUDA tensor with one entry per plan angle")
        angle_map = angle_map.contiguous()
        amap_ptr = angle_map.data_ptr()
    loglik = torch.empty((B,), dtype=torch.float32, device=img.device)
    dproj = torch.empty((B, plan.A, plan.W), dtype=torch.float32, device=img.device)
    ws = _workspace(plan.loglik_workspace_bytes(B), img.device)
    _lib.check(_lib.lib().ctr_radon_loglik(plan.handle, img.data_ptr(), mask.data_ptr(), meas.data_ptr(), amap_ptr, A_all,
                                           float(pnm), float(sqrt_reg), loglik.data_ptr(), dproj.data_ptr(), B, interp,
                                           ws.data_ptr(), ws.numel(), _stream_ptr(img.device)))
    return loglik, dproj


class LoglikFunction(torch.autograd.Function):
    """sum_b loglik[b] with the gradient d/d img = A^T (d loglik / d proj): the forward pass
    already produced the cotangent, so backward is a single (scaled) adjoint launch."""

    @staticmethod
    def forward(ctx, img, plan, mask, meas, angle_map, pnm, sqrt_reg, interp, mode, sel=None):
        loglik, dproj = radon_loglik(img, plan, mask, meas, angle_map, pnm, sqrt_reg, interp, sel)
        ctx.plan, ctx.interp, ctx.mode, ctx.sel = plan, interp, mode, sel
        ctx.save_for_backward(dproj)
        return loglik

    @staticmethod
    def backward(ctx, grad_loglik):
        (dproj,) = ctx.saved_tensors
        g = grad_loglik.reshape(-1)
        # The upstream weight is one scalar for a summed loss: it can ride along as the adjoint's `scale` argument --
        # but reading it costs a device->host sync, which stalls the launch queue and is illegal under CUDA-graph
        # capture.  Scaling the cotangent on the device is one small elementwise launch and needs neither.
        if torch.cuda.is_current_stream_capturing() or g.numel() > 1:
            cot = dproj * g.view(-1, 1, 1)
            dimg = radon_adjoint(cot, ctx.plan, ctx.interp, ctx.mode, ctx.sel)
        else:
            dimg = radon_adjoint_scaled(dproj, ctx.plan, ctx.interp, ctx.mode, float(g[0]), ctx.sel)
        return (dimg,) + (None,) * 9
