"""NumPy stand-in for the tfp calls of the reference's fbp_tensorflow.py (tfp 0.14.0).  TEST INFRASTRUCTURE ONLY."""
import types

import numpy as np


def interp_regular_1d_grid(x, x_ref_min, x_ref_max, y_ref, axis=-1, fill_value="constant_extension",
                           fill_value_below=None, fill_value_above=None, grid_regularizing_transform=None, name=None):
    """Linear interpolation on a regular grid; output shape y_ref.shape[:axis] + x.shape + y_ref.shape[axis+1:]."""
    if fill_value != "constant_extension" or fill_value_below is not None or fill_value_above is not None:
        raise NotImplementedError("shim: only constant_extension")
    x = np.asarray(x)
    y_ref = np.moveaxis(np.asarray(y_ref), axis, -1)
    ny = y_ref.shape[-1]
    idx_unclipped = (x - x_ref_min) / (x_ref_max - x_ref_min) * (ny - 1)
    idx = np.clip(idx_unclipped, 0.0, ny - 1.0)
    below = np.floor(idx)
    above = np.minimum(below + 1, ny - 1)
    below = np.maximum(above - 1, 0)
    t = (idx - below).astype(y_ref.dtype)
    yb = y_ref[..., below.astype(np.int64)]           # [..., *x.shape]
    ya = y_ref[..., above.astype(np.int64)]
    y = t * ya + (1 - t) * yb
    y = np.where(x < x_ref_min, y_ref[..., :1].reshape(y_ref.shape[:-1] + (1,) * x.ndim), y)
    y = np.where(x > x_ref_max, y_ref[..., -1:].reshape(y_ref.shape[:-1] + (1,) * x.ndim), y)
    lead = y_ref.ndim - 1
    ax = axis if axis >= 0 else axis + lead + 1
    # move the x dims to where `axis` was
    order = list(range(ax)) + list(range(lead, lead + x.ndim)) + list(range(ax, lead))
    return np.transpose(y, order)


math = types.SimpleNamespace(interp_regular_1d_grid=interp_regular_1d_grid)
distributions = types.SimpleNamespace()   # forward_functions.py:16 only binds the name
