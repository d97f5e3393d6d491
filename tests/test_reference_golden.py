"""The reference's own hot-path files, executed under the NumPy TF stand-ins (oracle/tf_shim, generator
tests/golden/make_reference_golden.py), against (1) the CPU oracle and (2) the CUDA path through the
drop-in Python API.  This pins what the reference's own code decides (padding, layouts, -theta sign,
summed axis, interpolation defaults, iradon conventions); the arithmetic inside the third-party ops is a
restatement on both sides (parity unpinned there).

Tolerances: float32 forward rel-L2 <= 1e-5 (north_star); iradon (float32 values on the GPU vs float64) 1e-5 too.
"""
import os

import numpy as np
import pytest

from oracle import radon_oracle as orc

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_shim_golden.npz"))


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


# ------------------------------------------------------------------------------------ CPU: oracle vs reference
def test_toy_known_answer():
    # scripts/images_to_sinograms.py:54-59: theta=0 -> column sums, theta=pi/2 -> reversed row sums
    np.testing.assert_allclose(G["toy_out"][0, :, :, 0], [[0.4, 0.6], [0.7, 0.3]], rtol=0, atol=1e-6)


def test_oracle_pad_rule():
    for k, shape_xy in (("pad2", (5, 7)), ("pad3", (6, 4)), ("padv", (5, 5))):
        P = orc.num_proj_pix(*shape_xy)
        out = G[k + "_out"]
        ax = 1 if k == "padv" else 0
        assert out.shape[ax] == P and out.shape[ax + 1] == P
        H, W, padx, pady = orc.frame_of(shape_xy[0], shape_xy[1], True)
        sl = [slice(None)] * out.ndim
        sl[ax], sl[ax + 1] = slice(padx, padx + shape_xy[0]), slice(pady, pady + shape_xy[1])
        np.testing.assert_array_equal(out[tuple(sl)], G[k + "_in"])
        assert float(np.abs(out).sum()) == pytest.approx(float(np.abs(G[k + "_in"]).sum()), rel=1e-6)


def test_oracle_forward_vae_layout_nearest_default():
    s = orc.forward(G["vae_in"][..., 0], G["theta12"], True, orc.NEAREST)
    assert rel(s, G["vae_out"][..., 0]) <= 1e-6
    assert rel(s, G["vae_out_theta32"][..., 0]) <= 1e-6
    # and the default really is nearest: bilinear differs visibly
    assert rel(orc.forward(G["vae_in"][..., 0], G["theta12"], True, orc.BILINEAR), G["vae_out"][..., 0]) > 1e-3


def test_oracle_forward_xy_and_xyz_layouts():
    s = orc.forward(G["xy_in"][None], G["theta12"], True, orc.NEAREST)            # [1,A,P]
    assert rel(s[0], G["xy_out"][:, :, 0]) <= 1e-6
    s = orc.forward(np.transpose(G["xyz_in"], (2, 0, 1)).astype(np.float32), G["theta12"], False, orc.NEAREST)
    assert rel(np.transpose(s, (1, 2, 0)), G["xyz_out"]) <= 1e-6                 # float32 oracle vs float64 reference


def test_oracle_low_mem_is_bilinear():
    s = orc.forward(np.transpose(G["pad3_in"], (2, 0, 1)), G["theta12"], True, orc.BILINEAR)
    assert rel(np.transpose(s, (1, 2, 0)), G["lm_out32"]) <= 1e-6
    s = orc.forward(np.transpose(G["xyz_in"], (2, 0, 1)).astype(np.float32), G["theta12"], True, orc.BILINEAR)
    assert rel(np.transpose(s, (1, 2, 0)), G["lm_out64"]) <= 1e-6


def test_oracle_iradon():
    for key, xs, ys, filt in (("fbp_out_ramp", 20, 20, G["fbp_ramp"]), ("fbp_out_none", 20, 20, np.ones(32)),
                              ("fbp_out_rect", 14, 22, G["fbp_ramp"])):
        r = orc.iradon(G["fbp_sino"], G["theta12"], xs, ys, filt)
        assert rel(r, G[key]) <= 1e-10, key
    np.testing.assert_allclose(orc.get_fourier_filter(32, "ramp").reshape(-1), G["fbp_ramp"], rtol=1e-12, atol=1e-14)


# ------------------------------------------------------------------------------------ GPU: drop-in API vs reference
@pytest.mark.gpu
def test_gpu_pad_phantom_layouts():
    import ct_pvae_b200 as cp
    np.testing.assert_array_equal(cp.pad_phantom(G["pad2_in"], dim=2), G["pad2_out"])
    np.testing.assert_array_equal(cp.pad_phantom(G["pad3_in"], dim=3), G["pad3_out"])
    np.testing.assert_array_equal(cp.pad_phantom(G["padv_in"], integrate_vae=True), G["padv_out"])


@pytest.mark.gpu
def test_gpu_project_tf_fast_layouts():
    import torch

    import ct_pvae_b200 as cp
    th = G["theta12"]
    out = cp.project_tf_fast(G["vae_in"], th, pad=True, dim=2, integrate_vae=True)
    assert out.shape == G["vae_out"].shape and rel(out, G["vae_out"]) <= 1e-5
    out = cp.project_tf_fast(torch.from_numpy(G["vae_in"]).cuda(), torch.from_numpy(th.astype(np.float32)), pad=True,
                             dim=2, integrate_vae=True)
    assert rel(out.cpu().numpy(), G["vae_out_theta32"]) <= 1e-5
    out = cp.project_tf_fast(G["xy_in"], th, pad=True, dim=2)
    assert out.shape == G["xy_out"].shape and rel(out, G["xy_out"]) <= 1e-5
    out = cp.project_tf_fast(G["xyz_in"], th, pad=False, dim=3)
    assert out.shape == G["xyz_out"].shape and out.dtype == np.float64 and rel(out, G["xyz_out"]) <= 1e-5
    out = cp.project_tf_fast(G["toy_in"], G["toy_theta"], pad=False, dim=2, integrate_vae=True)
    np.testing.assert_allclose(out, G["toy_out"], rtol=0, atol=1e-6)


@pytest.mark.gpu
def test_gpu_project_tf_low_mem():
    import ct_pvae_b200 as cp
    out = cp.project_tf_low_mem(G["pad3_in"], G["theta12"], pad=True)
    assert out.shape == G["lm_out32"].shape and rel(out, G["lm_out32"]) <= 1e-5
    out = cp.project_tf_low_mem(G["xyz_in"], G["theta12"], pad=True)
    assert out.shape == G["lm_out64"].shape and rel(out, G["lm_out64"]) <= 1e-5


@pytest.mark.gpu
def test_gpu_iradon():
    import ct_pvae_b200 as cp
    for key, xs, ys, filt in (("fbp_out_ramp", 20, 20, G["fbp_ramp"]), ("fbp_out_none", 20, 20, np.ones(32)),
                              ("fbp_out_rect", 14, 22, G["fbp_ramp"])):
        r = cp.iradon(G["fbp_sino"], G["theta12"], xs, ys, filt)
        assert np.asarray(r).shape == G[key].shape and rel(r, G[key]) <= 1e-5, key
    with pytest.raises(ValueError):
        cp.iradon(G["fbp_sino"], G["theta12"][:-1], 20, 20, G["fbp_ramp"])


# ------------------------------------------------------------------------------------ fixture freshness (build container only)
@pytest.mark.skipif(not os.path.isdir("/root/reference/ctvae"), reason="the reference tree exists only in the build container")
def test_fixture_regenerates_identically(tmp_path):
    """Re-runs the reference's files under the shim in a subprocess and checks the committed fixture bit for bit."""
    import subprocess
    import sys
    gen = os.path.join(os.path.dirname(__file__), "golden", "make_reference_golden.py")
    code = ("import runpy, numpy as np, sys; m = runpy.run_path(%r); ff, fbp = m['load_reference']();"
            "g = np.load(%r); th = g['theta12'];"
            "a = ff.project_tf_fast(g['vae_in'], th, pad=True, dim=2, integrate_vae=True);"
            "b = ff.project_tf_low_mem(g['xyz_in'], th, pad=True);"
            "c = fbp.iradon(g['fbp_sino'], th, 20, 20, g['fbp_ramp']);"
            "ok = np.array_equal(a, g['vae_out']) and np.array_equal(b, g['lm_out64']) and np.array_equal(c, g['fbp_out_ramp']);"
            "sys.exit(0 if ok else 1)") % (gen, os.path.join(os.path.dirname(__file__), "golden", "reference_shim_golden.npz"))
    assert subprocess.run([sys.executable, "-c", code], cwd=str(tmp_path)).returncode == 0
