"""On-disk formats and evaluation helpers either side of the Radon path.

  scripts/images_to_sinograms.py:74-79  x_train_sinograms.npy [N,A,P], dataset_parameters.npy (pickled
                                        object array [theta, num_proj_pix]), x_size.npy, y_size.npy
  ctvae/helper_functions.py:50-56       get_sinograms
  ctvae/helper_functions.py:394-430     compare (MSE / SSIM / PSNR, skimage.metrics) and crop
  scripts/images_to_sinograms.py:61-68  images -> sinograms (tomopy.project there; the B200 projector here)

Host-side glue only (numpy / scipy); nothing here is on the GPU hot path.
"""
from __future__ import annotations

import os

import numpy as np


def save_dataset(save_path: str, x_train_sinograms, theta, x_size: int, y_size: int) -> None:
    os.makedirs(save_path, exist_ok=True)
    s = np.asarray(x_train_sinograms)
    s = np.where(s < 0, 0, s)                                                   # images_to_sinograms.py:72
    np.save(os.path.join(save_path, "x_train_sinograms.npy"), s)
    params = np.empty(2, dtype=object)
    params[0], params[1] = np.asarray(theta), int(s.shape[-1])
    np.save(os.path.join(save_path, "dataset_parameters.npy"), params, allow_pickle=True)
    np.save(os.path.join(save_path, "x_size.npy"), x_size)
    np.save(os.path.join(save_path, "y_size.npy"), y_size)


def get_sinograms(save_path: str):
    """(x_train_sinograms [N,A,P], theta [A], num_proj_pix) -- helper_functions.py:50-56."""
    theta, num_proj_pix = np.load(os.path.join(save_path, "dataset_parameters.npy"), allow_pickle=True)
    return np.load(os.path.join(save_path, "x_train_sinograms.npy")), np.asarray(theta, np.float64), int(num_proj_pix)


def reconstruction_size(num_proj_pix: int, no_pad: bool = False):
    """main_ct_vae.py:156-161: image side the driver reconstructs for a detector width."""
    if no_pad:
        return num_proj_pix, num_proj_pix
    n = int(np.floor(num_proj_pix / np.sqrt(2) - 2))
    return n, n


def images_to_sinograms(images, theta, pad=True, save_path=None, interpolation="bilinear"):
    """scripts/images_to_sinograms.py:61-79 on the B200 projector: [N,X,Y] -> [N,A,P] (+ files)."""
    import torch

    from .forward_functions import project_tf_fast

    t = torch.as_tensor(np.asarray(images, np.float32))
    s = project_tf_fast(t.unsqueeze(-1), theta, pad=pad, dim=2, integrate_vae=True, interpolation=interpolation)[..., 0]
    s = s.cpu().numpy()
    if save_path is not None:
        save_dataset(save_path, s, theta, t.shape[1], t.shape[2])
    return s


def crop(img_2d, final_x, final_y, ignore_dim_0=False):
    """helper_functions.py:420-430: centred crop."""
    x, y = img_2d.shape[-2:]
    rx, ry = final_x % 2, final_y % 2
    sl = (slice(x // 2 - final_x // 2, x // 2 + final_x // 2 + rx), slice(y // 2 - final_y // 2, y // 2 + final_y // 2 + ry))
    return img_2d[(slice(None),) + sl] if ignore_dim_0 else img_2d[sl]


def _ssim(im0, im1, data_range, win_size=None):
    """skimage.metrics.structural_similarity defaults (uniform window 7, K1=.01, K2=.03, sample covariance).
    Restates the published algorithm of scikit-image's ``structural_similarity`` (Copyright (C) the scikit-image
    team, BSD-3-Clause; see THIRD_PARTY_NOTICES.md), which the reference calls at helper_functions.py:394-418."""
    from scipy.ndimage import uniform_filter

    win = 7 if win_size is None else win_size
    a, b = np.asarray(im0, np.float64), np.asarray(im1, np.float64)
    npx = win ** a.ndim
    cov_norm = npx / (npx - 1)
    ux, uy = uniform_filter(a, win), uniform_filter(b, win)
    uxx, uyy, uxy = uniform_filter(a * a, win), uniform_filter(b * b, win), uniform_filter(a * b, win)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux ** 2 + uy ** 2 + c1) * (vx + vy + c2))
    pad = (win - 1) // 2
    return float(s[tuple(slice(pad, n - pad) for n in s.shape)].mean())


def compare(recon0, recon1, verbose=False):
    """helper_functions.py:394-418: (MSE, SSIM, PSNR) with data_range = range of recon0."""
    a, b = np.asarray(recon0, np.float64), np.asarray(recon1, np.float64)
    mse = float(np.mean((a - b) ** 2))
    small = min(a.shape)
    win = (small if small % 2 else small - 1) if small < 7 else None
    dr = float(a.max() - a.min())
    ssim = _ssim(a, b, dr, win)
    psnr = float(10 * np.log10(dr ** 2 / mse)) if mse > 0 else float("inf")
    if verbose:
        print("MSE: {:.8f}, SSIM: {:.3f}, PSNR: {:.3f}".format(mse, ssim, psnr))
    return mse, ssim, psnr
