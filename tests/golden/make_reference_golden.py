"""Runs the reference's OWN, unmodified hot-path files and stores what they return.

  /root/reference/ctvae/forward_functions.py   pad_phantom, project_tf_fast, project_tf_low_mem
  /root/reference/ctvae/fbp_tensorflow.py      iradon

TensorFlow / tensorflow-addons / tensorflow-probability are not installable in this image, so the three
packages are replaced by the NumPy stand-ins under oracle/tf_shim/ (see its README for exactly what that
does and does not pin).  The reference cannot travel to the GPU box, hence the committed fixture
tests/golden/reference_shim_golden.npz.          Run (in the build container only):

    python tests/golden/make_reference_golden.py
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("CTR_REFERENCE", "/root/reference")


def load_reference():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_shim"))
    mods = {}
    for name in ("forward_functions", "fbp_tensorflow"):
        spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, "ctvae", name + ".py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        mods[name] = m
    return mods["forward_functions"], mods["fbp_tensorflow"]


def ramp_filter(size):
    """skimage's _get_fourier_filter(size, 'ramp') (the filter main_ct_vae.py:22 refers to), restated."""
    n = np.concatenate((np.arange(1, size / 2 + 1, 2, dtype=int), np.arange(size / 2 - 1, 0, -2, dtype=int)))
    f = np.zeros(size)
    f[0] = 0.25
    f[1::2] = -1 / (np.pi * n) ** 2
    return 2 * np.real(np.fft.fft(f))


def main():
    ff, fbp = load_reference()
    rng = np.random.default_rng(20261019)
    out = {}

    # pad_phantom in its three layouts (forward_functions.py:18-46)
    out["pad2_in"] = rng.random((5, 7)).astype(np.float32)
    out["pad2_out"] = ff.pad_phantom(out["pad2_in"], dim=2)
    out["pad3_in"] = rng.random((6, 4, 2)).astype(np.float32)
    out["pad3_out"] = ff.pad_phantom(out["pad3_in"], dim=3)
    out["padv_in"] = rng.random((2, 5, 5, 1)).astype(np.float32)
    out["padv_out"] = ff.pad_phantom(out["padv_in"], integrate_vae=True)

    # project_tf_fast, VAE layout [B,X,Y,1] -> [B,A,P,1], padded, tfa default interpolation (helper_functions.py:359)
    th12 = np.linspace(0, np.pi, 12, endpoint=False)
    out["theta12"] = th12
    out["vae_in"] = rng.random((3, 20, 20, 1)).astype(np.float32)
    out["vae_out"] = ff.project_tf_fast(out["vae_in"], th12, pad=True, dim=2, integrate_vae=True)
    # float32 tensor theta, as helper_functions.py:355 passes it
    out["vae_out_theta32"] = ff.project_tf_fast(out["vae_in"], th12.astype(np.float32), pad=True, dim=2, integrate_vae=True)
    # [X,Y] NumPy image, fp64 theta (main_ct_vae.py:523-524)
    out["xy_in"] = rng.random((16, 24)).astype(np.float32)
    out["xy_out"] = ff.project_tf_fast(out["xy_in"], th12, pad=True, dim=2)
    # [X,Y,Z] float64, no padding (tomopy_forward_compare.py:52)
    out["xyz_in"] = rng.random((18, 18, 3))
    out["xyz_out"] = ff.project_tf_fast(out["xyz_in"], th12, pad=False, dim=3)
    # project_tf_low_mem: bilinear, [X,Y,Z] (tomopy_forward_compare.py:56), float32 and float64
    out["lm_out32"] = ff.project_tf_low_mem(out["pad3_in"], th12, pad=True)
    out["lm_out64"] = ff.project_tf_low_mem(out["xyz_in"], th12, pad=True)
    # toy dataset: 2x2, no padding, theta = [0, pi/2] (scripts/images_to_sinograms.py:28-31, toy_mcmc_v2_functions.py:41)
    toy = (np.array([[1, 2], [3, 4]], np.float32) / 10)[None, :, :, None]
    out["toy_theta"] = np.array([0, np.pi / 2])
    out["toy_in"] = toy
    out["toy_out"] = ff.project_tf_fast(toy, out["toy_theta"], pad=False, dim=2, integrate_vae=True)

    # iradon (fbp_tensorflow.py:14-75): ramp filter and the all-ones filter iradon_all uses for the mask
    P = out["vae_out"].shape[2]
    sino = ff.project_tf_low_mem(np.transpose(out["vae_in"][..., 0], (1, 2, 0)), th12, pad=True)   # [A,P,B]
    out["fbp_sino"] = np.ascontiguousarray(np.transpose(sino, (2, 0, 1))).astype(np.float64)
    out["fbp_ramp"] = ramp_filter(P)
    out["fbp_out_ramp"] = fbp.iradon(out["fbp_sino"], th12, 20, 20, out["fbp_ramp"])
    out["fbp_out_none"] = fbp.iradon(out["fbp_sino"], th12, 20, 20, np.ones(P))
    out["fbp_out_rect"] = fbp.iradon(out["fbp_sino"], th12, 14, 22, out["fbp_ramp"])
    try:
        fbp.iradon(out["fbp_sino"], th12[:-1], 20, 20, out["fbp_ramp"])
        raise AssertionError("reference iradon accepted a theta of the wrong length")
    except ValueError:
        pass

    np.savez_compressed(os.path.join(HERE, "reference_shim_golden.npz"), **out)
    print({k: (np.asarray(v).shape, str(np.asarray(v).dtype)) for k, v in out.items()})


if __name__ == "__main__":
    main()
