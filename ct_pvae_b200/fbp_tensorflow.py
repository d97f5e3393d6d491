"""Drop-in for ``ctvae/fbp_tensorflow.py`` of vganapati/CT_PVAE.

``iradon(sinogram, theta, x_size, y_size, filter_1d)`` keeps the reference signature
(``/root/reference/ctvae/fbp_tensorflow.py:14-75``): ``sinogram`` is
``[batch, num_angles, num_proj_pix]``, ``filter_1d`` the length-``num_proj_pix``
frequency-domain filter in FFT order, and the result is float64 ``[batch, x_size,
y_size]``.  On the GPU the FFT product becomes the equivalent circular convolution
with ``real(ifft(filter_1d))`` done in shared memory, followed by the linear
interpolating back-projection with tfp's edge clamp, scaled by ``pi / (2 A)``.
Arithmetic is float32 on the device (geometry in float64); the float64 return dtype
of the reference is preserved by widening.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, ops
from .forward_functions import _as_tensor, _compute_device

__all__ = ["iradon", "get_fourier_filter"]


def get_fourier_filter(size: int, filter_name="ramp") -> np.ndarray:
    """The ``filter_1d`` the reference intended to pass (skimage's
    ``_get_fourier_filter``, named at main_ct_vae.py:22,183): length ``size``, FFT order.

    Restates ``skimage.transform.radon_transform._get_fourier_filter`` of scikit-image (Copyright (C) the
    scikit-image team, BSD-3-Clause; see THIRD_PARTY_NOTICES.md) so that the values are the ones the reference's
    callers would have computed; scikit-image itself is not a dependency."""
    n = np.concatenate((np.arange(1, size / 2 + 1, 2, dtype=int), np.arange(size / 2 - 1, 0, -2, dtype=int)))
    f = np.zeros(size)
    f[0] = 0.25
    f[1::2] = -1 / (np.pi * n) ** 2
    fourier_filter = 2 * np.real(np.fft.fft(f))
    if filter_name == "ramp":
        pass
    elif filter_name == "shepp-logan":
        omega = np.pi * np.fft.fftfreq(size)[1:]
        fourier_filter[1:] *= np.sin(omega) / omega
    elif filter_name == "cosine":
        fourier_filter *= np.fft.fftshift(np.sin(np.linspace(0, np.pi, size, endpoint=False)))
    elif filter_name == "hamming":
        fourier_filter *= np.fft.fftshift(np.hamming(size))
    elif filter_name == "hann":
        fourier_filter *= np.fft.fftshift(np.hanning(size))
    elif filter_name is None:
        fourier_filter[:] = 1
    else:
        raise ValueError(f"Unknown filter: {filter_name}")
    return fourier_filter


def iradon(sinogram, theta, x_size, y_size, filter_1d, *, fused=None):
    """Filtered back-projection (reference :14-75).

    ``fused`` (keyword-only extra): ``None`` / ``False`` -- the library default, two kernels (row filter, then the
    gather); ``True`` -- ONE thread-block-cluster kernel that filters the rows in shared memory and back-projects them
    from there (images up to 128 x 128; same values)."""
    t, was_numpy = _as_tensor(sinogram)
    if t.dim() != 3:
        raise ValueError("sinogram must be [batch, num_angles, num_proj_pix]")
    num_angles = len(theta)
    if num_angles != t.shape[1]:
        raise ValueError("The given ``theta`` does not match the number of "
                         "projections in ``radon_image``.")
    dev = _compute_device(t)
    if isinstance(filter_1d, torch.Tensor):
        filter_1d = filter_1d.detach().cpu().numpy()
    plan = _lib.get_fbp_plan(ops.theta_to_host(theta), int(t.shape[2]), int(x_size), int(y_size),
                             np.asarray(filter_1d), dev.index or 0)
    if fused is not None:
        plan.set_fused(bool(fused))
    try:
        rec = ops.fbp(t.to(device=dev, dtype=torch.float32), plan)
    finally:
        if fused is not None:
            plan.set_fused(False)
    out = rec.to(device=t.device, dtype=torch.float64)
    return out.numpy() if was_numpy else out
