"""CPU oracle for CT_PVAE's parallel-beam projector path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  Nothing under ``ct_pvae_b200/``
does, and the product path has no CPU fallback.

PARITY PARTLY PINNED: the reference ships no tests or golden vectors for this path and
TensorFlow / tensorflow-addons / tensorflow-probability are not installable here, so
the third-party arithmetic (tfa 0.17.1 ``rotate``, TF 2.8.1
``ImageProjectiveTransformV3`` and its gradient, tfp 0.14 ``interp_regular_1d_grid``)
is restated from their published algorithms.  The rotation is pinned by tensorflow-addons'
own known-answer tests (``test_rotate_even`` / ``test_rotate_odd`` / ``test_bilinear``, transcribed in
``tests/test_tfa_known_answers.py``): nearest exactly, bilinear to that test's 1e-3; TF's gradient and
tfp's interpolation stay UNPINNED.  Further anchors: the toy dataset's closed
form sinograms (reference ``scripts/images_to_sinograms.py:54-59``), theta=0 column
sums, mass conservation, the explicit sparse matrix, and an independent bilinear
implementation (``torch.nn.functional.grid_sample``) -- see ``tests/test_oracle.py``.

Two restatements live here on purpose:
  * ``radon_oracle.c`` (compiled, OpenMP) -- the one the GPU parity tests use;
  * the ``*_np`` functions below (vectorised numpy float32) -- written separately and
    used to cross-check the C file.
Reference anchors: ``ctvae/forward_functions.py:18-46`` (pad_phantom), ``:80-123``
(project_tf_fast), ``:49-78`` (project_tf_low_mem), ``ctvae/fbp_tensorflow.py:14-75``
(iradon).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

NEAREST, BILINEAR = 0, 1
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)


def build(force: bool = False) -> str:
    """Compile radon_oracle.c -> libradon_oracle.so (gcc, no FMA contraction)."""
    so = os.path.join(_HERE, "libradon_oracle.so")
    src = os.path.join(_HERE, "radon_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        i = ctypes.c_int
        L.orc_num_proj_pix.argtypes = [i, i]
        L.orc_num_proj_pix.restype = i
        L.orc_max_threads.restype = i
        L.orc_set_num_threads.argtypes = [i]
        L.orc_set_num_threads.restype = None
        L.orc_make_transforms.argtypes = [_f64p, i, i, i, _f32p]
        L.orc_invert_transforms.argtypes = [_f32p, i, _f32p]
        for name in ("orc_forward", "orc_forward_dataflow", "orc_adjoint_exact", "orc_adjoint_tf"):
            getattr(L, name).argtypes = [_f32p, i, i, i, i, i, i, i, _f32p, i, i, _f32p]
            getattr(L, name).restype = None
        L.orc_iradon_backproject.argtypes = [_f64p, _f64p, i, i, i, i, i, _f64p]
        L.orc_iradon_backproject.restype = None
        _LIB = L
    return _LIB


def _p32(a):
    return a.ctypes.data_as(_f32p)


def _p64(a):
    return a.ctypes.data_as(_f64p)


# ----------------------------------------------------------------------------- geometry
def num_proj_pix(X: int, Y: int) -> int:
    """pad_phantom's detector size (forward_functions.py:29-30)."""
    return int(np.ceil((np.sqrt(np.float64(X * X + Y * Y)) + 2) / 2.0) * 2)


def frame_of(X: int, Y: int, pad: bool):
    """(H, W, padx, pady): the frame project_tf_fast rotates and where the image sits in it."""
    if not pad:
        return X, Y, 0, 0
    P = num_proj_pix(X, Y)
    return P, P, (P - X) // 2, (P - Y) // 2


def make_transforms(theta, H: int, W: int) -> np.ndarray:
    """[A,8] float32 table of tfa.image.rotate(images, -theta) (C restatement)."""
    th = np.ascontiguousarray(np.asarray(theta, dtype=np.float64).reshape(-1))
    t = np.empty((th.size, 8), np.float32)
    lib().orc_make_transforms(_p64(th), th.size, H, W, _p32(t))
    return t


def make_transforms_np(theta, H: int, W: int) -> np.ndarray:
    """Same table in numpy float32 (independent restatement)."""
    ang = (-np.asarray(theta, dtype=np.float64).reshape(-1)).astype(np.float32)
    c, s = np.cos(ang), np.sin(ang)  # float32 in, float32 out
    wm1, hm1 = np.float32(W - 1), np.float32(H - 1)
    x_off = (wm1 - (c * wm1 - s * hm1)) / np.float32(2)
    y_off = (hm1 - (s * wm1 + c * hm1)) / np.float32(2)
    z = np.zeros_like(c)
    return np.stack([c, -s, x_off, s, c, y_off, z, z], axis=1).astype(np.float32)


def invert_transforms(t: np.ndarray) -> np.ndarray:
    t = np.ascontiguousarray(t, np.float32)
    out = np.empty_like(t)
    lib().orc_invert_transforms(_p32(t), t.shape[0], _p32(out))
    return out


# ----------------------------------------------------------------------------- C oracle wrappers
def _call(fn, src, shape_out, B, X, Y, H, W, padx, pady, t, interp):
    src = np.ascontiguousarray(src, np.float32)
    t = np.ascontiguousarray(t, np.float32)
    out = np.empty(shape_out, np.float32)
    fn(_p32(src), B, X, Y, H, W, padx, pady, _p32(t), t.shape[0], int(interp), _p32(out))
    return out


def forward(img, theta, pad: bool, interp: int, dataflow: bool = False) -> np.ndarray:
    """img [B,X,Y] float32 -> sino [B,A,W]; restates project_tf_fast (interp=NEAREST) /
    project_tf_low_mem (interp=BILINEAR)."""
    img = np.asarray(img, np.float32)
    B, X, Y = img.shape
    H, W, padx, pady = frame_of(X, Y, pad)
    t = make_transforms(theta, H, W)
    fn = lib().orc_forward_dataflow if dataflow else lib().orc_forward
    return _call(fn, img, (B, t.shape[0], W), B, X, Y, H, W, padx, pady, t, interp)


def adjoint_exact(y, theta, X: int, Y: int, pad: bool, interp: int) -> np.ndarray:
    """y [B,A,W] -> g [B,X,Y]: the exact transpose of ``forward``."""
    y = np.asarray(y, np.float32)
    B = y.shape[0]
    H, W, padx, pady = frame_of(X, Y, pad)
    assert y.shape[2] == W
    t = make_transforms(theta, H, W)
    return _call(lib().orc_adjoint_exact, y, (B, X, Y), B, X, Y, H, W, padx, pady, t, interp)


def adjoint_tf(y, theta, X: int, Y: int, pad: bool, interp: int) -> np.ndarray:
    """y [B,A,W] -> g [B,X,Y]: TensorFlow's registered gradient of the projector graph."""
    y = np.asarray(y, np.float32)
    B = y.shape[0]
    H, W, padx, pady = frame_of(X, Y, pad)
    assert y.shape[2] == W
    tinv = invert_transforms(make_transforms(theta, H, W))
    return _call(lib().orc_adjoint_tf, y, (B, X, Y), B, X, Y, H, W, padx, pady, tinv, interp)


# ----------------------------------------------------------------------------- numpy restatement
def _coords_np(t_row, H, W):
    """float32 input coordinates for every output pixel (i rows, j cols), TF expression order."""
    ox = np.arange(W, dtype=np.float32)[None, :]
    oy = np.arange(H, dtype=np.float32)[:, None]
    t = t_row.astype(np.float32)
    x = (t[0] * ox + t[1] * oy) + t[2]
    y = (t[3] * ox + t[4] * oy) + t[5]
    return x.astype(np.float32), y.astype(np.float32)


def _round_half_away(v):
    return np.where(v >= 0, np.floor(v + np.float32(0.5)), np.ceil(v - np.float32(0.5)))


def _round_half_away_exact(v):
    """std::round for float32 without the v+0.5 double-rounding trap."""
    v64 = v.astype(np.float64)
    return np.sign(v64) * np.floor(np.abs(v64) + 0.5)


def _read(padded, yy, xx):
    H, W = padded.shape[-2:]
    ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
    yyc = np.clip(yy, 0, H - 1).astype(np.int64)
    xxc = np.clip(xx, 0, W - 1).astype(np.int64)
    return np.where(ok, padded[..., yyc, xxc], np.float32(0))


def forward_np(img, theta, pad: bool, interp: int) -> np.ndarray:
    """Vectorised numpy restatement of project_tf_fast's dataflow (pad -> rotate -> row sum)."""
    img = np.asarray(img, np.float32)
    B, X, Y = img.shape
    H, W, padx, pady = frame_of(X, Y, pad)
    padded = np.zeros((B, H, W), np.float32)
    padded[:, padx:padx + X, pady:pady + Y] = img
    t = make_transforms_np(theta, H, W)
    out = np.empty((B, t.shape[0], W), np.float32)
    for a in range(t.shape[0]):
        x, y = _coords_np(t[a], H, W)
        if interp == NEAREST:
            rot = _read(padded, _round_half_away_exact(y), _round_half_away_exact(x))
        else:
            yf, xf = np.floor(y), np.floor(x)
            yc, xc = yf + np.float32(1), xf + np.float32(1)
            v_f = (xc - x) * _read(padded, yf, xf) + (x - xf) * _read(padded, yf, xc)
            v_c = (xc - x) * _read(padded, yc, xf) + (x - xf) * _read(padded, yc, xc)
            rot = (yc - y) * v_f + (y - yf) * v_c
        out[:, a, :] = rot.astype(np.float64).sum(axis=1).astype(np.float32)
    return out


def build_matrix(theta, X: int, Y: int, pad: bool, interp: int):
    """Explicit sparse A [(A*W) x (X*Y)] float64 of the forward map (small sizes only)."""
    import scipy.sparse as sp

    H, W, padx, pady = frame_of(X, Y, pad)
    t = make_transforms_np(theta, H, W)
    rows, cols, vals = [], [], []
    for a in range(t.shape[0]):
        x, y = _coords_np(t[a], H, W)
        jj = np.broadcast_to(np.arange(W)[None, :], (H, W))
        if interp == NEAREST:
            taps = [(_round_half_away_exact(y), _round_half_away_exact(x), np.ones((H, W)))]
        else:
            yf, xf = np.floor(y), np.floor(x)
            yc, xc = yf + np.float32(1), xf + np.float32(1)
            wxf, wxc = (xc - x).astype(np.float64), (x - xf).astype(np.float64)
            wyf, wyc = (yc - y).astype(np.float64), (y - yf).astype(np.float64)
            taps = [(yf, xf, wyf * wxf), (yf, xc, wyf * wxc), (yc, xf, wyc * wxf), (yc, xc, wyc * wxc)]
        for ty, tx, w in taps:
            r = ty.astype(np.int64) - padx
            c = tx.astype(np.int64) - pady
            ok = (ty >= 0) & (ty < H) & (tx >= 0) & (tx < W) & (r >= 0) & (r < X) & (c >= 0) & (c < Y)
            rows.append((a * W + jj)[ok])
            cols.append((r * Y + c)[ok])
            vals.append(np.asarray(w, np.float64)[ok])
    return sp.csr_matrix(
        (np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
        shape=(t.shape[0] * W, X * Y),
    )


# ----------------------------------------------------------------------------- FBP (fbp_tensorflow.py)
def get_fourier_filter(size: int, filter_name):
    """Restatement of skimage.transform.radon_transform._get_fourier_filter (the
    ``filter_1d`` the reference fed to iradon, main_ct_vae.py:22,183, commented out)."""
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
        freq = np.linspace(0, np.pi, size, endpoint=False)
        fourier_filter *= np.fft.fftshift(np.sin(freq))
    elif filter_name == "hamming":
        fourier_filter *= np.fft.fftshift(np.hamming(size))
    elif filter_name == "hann":
        fourier_filter *= np.fft.fftshift(np.hanning(size))
    elif filter_name is None:
        fourier_filter[:] = 1
    else:
        raise ValueError(f"unknown filter {filter_name!r}")
    return fourier_filter


def iradon(sinogram, theta, x_size: int, y_size: int, filter_1d) -> np.ndarray:
    """float64 restatement of ctvae/fbp_tensorflow.py:14-75 (numpy FFT + C back-projection)."""
    sinogram = np.asarray(sinogram)
    theta = np.ascontiguousarray(np.asarray(theta, np.float64).reshape(-1))
    if theta.size != sinogram.shape[1]:
        raise ValueError("The given ``theta`` does not match the number of projections in ``radon_image``.")
    B, A, P = sinogram.shape
    projection = np.fft.fft(sinogram.astype(np.complex128), axis=-1) * np.asarray(filter_1d)
    rf = np.ascontiguousarray(np.real(np.fft.ifft(projection, axis=-1)), np.float64)
    out = np.empty((B, x_size, y_size), np.float64)
    lib().orc_iradon_backproject(_p64(rf), _p64(theta), B, A, P, x_size, y_size, _p64(out))
    return out


def iradon_np(sinogram, theta, x_size: int, y_size: int, filter_1d) -> np.ndarray:
    """Pure-numpy restatement of iradon following the reference line by line."""
    sinogram = np.asarray(sinogram)
    theta = np.asarray(theta, np.float64).reshape(-1)
    num_angles = len(theta)
    P = sinogram.shape[2]
    if num_angles != sinogram.shape[1]:
        raise ValueError("The given ``theta`` does not match the number of projections in ``radon_image``.")
    projection = np.fft.fft(sinogram.astype(np.complex128), axis=-1) * np.asarray(filter_1d)
    rf = np.real(np.fft.ifft(projection, axis=-1))
    xpr, ypr = np.meshgrid(np.arange(x_size, dtype=np.float64) - x_size / 2,
                           np.arange(y_size, dtype=np.float64) - y_size / 2, indexing="ij")
    coords = np.arange(P, dtype=np.float64) - P / 2
    lo, hi = coords.min(), coords.max()
    rec = np.zeros((sinogram.shape[0], x_size, y_size), np.float64)
    for a in range(num_angles):
        t = ypr * np.cos(theta[a]) - xpr * np.sin(theta[a])
        idx = np.clip((t - lo) / (hi - lo) * (P - 1), 0, P - 1)
        below = np.floor(idx)
        above = np.minimum(below + 1, P - 1)
        below = np.maximum(above - 1, 0)
        alpha = idx - below
        row = rf[:, a, :]
        rec += alpha * row[:, above.astype(np.int64)] + (1 - alpha) * row[:, below.astype(np.int64)]
    return rec * np.pi / (2 * num_angles)


# ----------------------------------------------------------------------------- measurement log-likelihood
def log_prob_M_given_R(img, mask, proj_sample, pnm, sqrt_reg, theta, angles_i=None, pad=True, interp=NEAREST):
    """float64 restatement of calculate_log_prob_M_given_R (ctvae/helper_functions.py:336-368)
    on top of the float32 projector oracle.  Returns (logp [B,A',P], d logp / d proj [B,A',P])."""
    theta = np.asarray(theta, np.float64).reshape(-1)
    mask = np.asarray(mask, np.float64)
    y = np.asarray(proj_sample, np.float64)
    if angles_i is not None:
        idx = np.asarray(angles_i, np.int64)
        theta = theta[idx].astype(np.float32).astype(np.float64)      # tf.cast(tf.gather(theta, angles_i), float32), :355
        mask, y = mask[:, idx], y[:, idx]
    proj = forward(np.asarray(img, np.float32), theta, pad, interp).astype(np.float64)
    m = mask[:, :, None]
    pm = proj * m
    sr = np.sqrt(pm / pnm + sqrt_reg)
    sc = sqrt_reg + sr
    z = (y - pm) / sc
    logp = -0.5 * z * z - np.log(sc) - 0.5 * np.log(2 * np.pi)
    dsc = 0.5 / (pnm * sr)
    dproj = (z / sc + (z * z - 1.0) / sc * dsc) * m
    return logp, dproj


# ----------------------------------------------------------------------------- synthetic inputs
def synthetic_foam(n: int, size: int, seed: int = 0) -> np.ndarray:
    """Stand-in for xdesign.Foam (scripts/create_foam_images.py:27-40): a unit disk
    with non-overlapping zero-valued circular pores.  xdesign itself is unavailable."""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.linspace(-1, 1, size), np.linspace(-1, 1, size), indexing="ij")
    out = np.zeros((n, size, size), np.float32)
    for k in range(n):
        img = (xx * xx + yy * yy <= 1.0).astype(np.float32)
        target = rng.random() * 0.6
        pores, area = [], 0.0
        for _ in range(400):
            if area >= target * np.pi:
                break
            r = rng.uniform(0.01, 0.2)
            ang, rad = rng.uniform(0, 2 * np.pi), np.sqrt(rng.random()) * (1 - r)
            cx, cy = rad * np.cos(ang), rad * np.sin(ang)
            if all((cx - px) ** 2 + (cy - py) ** 2 >= (r + pr) ** 2 for px, py, pr in pores):
                pores.append((cx, cy, r))
                area += np.pi * r * r
                img[(xx - cx) ** 2 + (yy - cy) ** 2 <= r * r] = 0.0
        out[k] = img
    return out
