"""tfa.image.rotate restated in numpy (tfa 0.17.1 transform_ops.py -> TF 2.8.1 ImageProjectiveTransformV3).

Written independently of oracle/radon_oracle.{c,py}; only fill_mode="constant" is implemented
(the only mode the reference uses: forward_functions.py:70-74 and :113).
"""
import numpy as np

_f32 = np.float32


def angles_to_projective_transforms(angles, image_height, image_width):
    ang = np.asarray(angles, dtype=np.float32)       # tfa casts the angles to float32 first
    if ang.ndim == 0:
        ang = ang[None]
    elif ang.ndim != 1:
        raise ValueError("angles should have rank 0 or 1.")
    h, w = _f32(image_height), _f32(image_width)
    co, si = np.cos(ang), np.sin(ang)                 # float32 in -> float32 out
    x_offset = ((w - _f32(1)) - (co * (w - _f32(1)) - si * (h - _f32(1)))) / _f32(2.0)
    y_offset = ((h - _f32(1)) - (si * (w - _f32(1)) + co * (h - _f32(1)))) / _f32(2.0)
    z = np.zeros_like(co)
    return np.stack([co, -si, x_offset, si, co, y_offset, z, z], axis=1).astype(np.float32)


def _read_with_fill(img, yy, xx, fill):
    """img [H,W,C]; integer index arrays [H_out,W_out]; out-of-range taps read `fill`."""
    H, W = img.shape[:2]
    ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
    v = img[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)]
    return np.where(ok[..., None], v, np.asarray(fill, dtype=img.dtype))


def _transform_one(img, t, interpolation, fill_value):
    H, W, _ = img.shape
    xo = np.arange(W, dtype=np.float32)[None, :]
    yo = np.arange(H, dtype=np.float32)[:, None]
    proj = t[6] * xo + t[7] * yo + _f32(1)
    x = (t[0] * xo + t[1] * yo + t[2]) / proj          # float32, each product and sum rounded (no FMA)
    y = (t[3] * xo + t[4] * yo + t[5]) / proj
    assert x.dtype == np.float32 and y.dtype == np.float32
    if interpolation == "NEAREST":
        rnd = lambda v: (np.sign(v) * np.floor(np.abs(v.astype(np.float64)) + 0.5)).astype(np.int64)  # std::round
        return _read_with_fill(img, rnd(y), rnd(x), fill_value)
    xf, yf = np.floor(x), np.floor(y)
    xc, yc = xf + _f32(1), yf + _f32(1)
    T = img.dtype
    xi, yi = xf.astype(np.int64), yf.astype(np.int64)
    v_yf = ((xc - x).astype(T)[..., None] * _read_with_fill(img, yi, xi, fill_value)
            + (x - xf).astype(T)[..., None] * _read_with_fill(img, yi, xi + 1, fill_value))
    v_yc = ((xc - x).astype(T)[..., None] * _read_with_fill(img, yi + 1, xi, fill_value)
            + (x - xf).astype(T)[..., None] * _read_with_fill(img, yi + 1, xi + 1, fill_value))
    return (yc - y).astype(T)[..., None] * v_yf + (y - yf).astype(T)[..., None] * v_yc


def rotate(images, angles, interpolation="nearest", fill_mode="constant", name=None, fill_value=0.0):
    if fill_mode.lower() != "constant":
        raise NotImplementedError("shim: only fill_mode='constant'")
    x = np.asarray(images)
    if x.dtype not in (np.float16, np.float32, np.float64, np.uint8, np.int32, np.int64):
        raise TypeError("unsupported image dtype")
    nd = x.ndim
    x4 = x[None, :, :, None] if nd == 2 else (x[None] if nd == 3 else x)
    if x4.ndim != 4:
        raise ValueError("images must have rank 2, 3 or 4")
    N, H, W, _ = x4.shape
    ts = angles_to_projective_transforms(angles, H, W)
    if ts.shape[0] not in (1, N):
        raise ValueError("number of angles must be 1 or the batch size")
    out = np.stack([_transform_one(x4[n], ts[n if ts.shape[0] > 1 else 0], interpolation.upper(), fill_value)
                    for n in range(N)])
    return out[0, :, :, 0] if nd == 2 else (out[0] if nd == 3 else out)
