"""The TensorFlow binding (ct_pvae_b200/tf_bridge.py) executed on the GPU against a torch-backed stand-in for the
``tf`` calls it makes (tests/tf_torch_shim): TensorFlow itself cannot be installed in this image, so this is the
closest the binding's own code gets to running here.  What it proves: the DLPack capsule handling, the plan cache and
device lookup, the two device synchronisations around the launch, and that the custom gradient is the adjoint kernel."""
import importlib
import os
import sys

import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu
SHIM = os.path.join(os.path.dirname(__file__), "tf_torch_shim")


@pytest.fixture()
def bridge():
    import ct_pvae_b200.tf_bridge as tb

    if tb.available() and "tf_torch_shim" not in getattr(tb.tf, "__file__", ""):
        pytest.skip("a real TensorFlow is installed: run the binding under it instead")
    sys.path.insert(0, SHIM)
    sys.modules.pop("tensorflow", None)
    try:
        tb = importlib.reload(tb)
        assert tb.available() and "tf_torch_shim" in tb.tf.__file__
        yield tb
    finally:
        sys.path.remove(SHIM)
        sys.modules.pop("tensorflow", None)
        importlib.reload(tb)


@pytest.mark.parametrize("interp,adjoint", [("nearest", "exact"), ("bilinear", "exact"), ("bilinear", "tf_compat")])
def test_tf_binding_forward_and_gradient(bridge, orc, interp, adjoint):
    import torch

    rng = np.random.default_rng(12)
    B, X, A = 6, 24, 9
    th = np.linspace(0, np.pi, A, endpoint=False)
    img = rng.random((B, X, X, 1), dtype=np.float32)
    iid = 0 if interp == "nearest" else 1
    x = torch.from_numpy(img).cuda().requires_grad_(True)
    s = bridge.project_tf_fast(x, torch.from_numpy(th), pad=True, dim=2, integrate_vae=True, interpolation=interp, adjoint=adjoint)
    want = orc.forward(img[..., 0], th, True, iid)
    assert tuple(s.shape) == (B, A, want.shape[2], 1)
    assert rel_l2(s[..., 0].detach().cpu().numpy(), want) <= 1e-5
    cot = rng.random(want.shape, dtype=np.float32)
    (s[..., 0] * torch.from_numpy(cot).cuda()).sum().backward()
    fn = orc.adjoint_exact if adjoint == "exact" else orc.adjoint_tf
    assert rel_l2(x.grad[..., 0].cpu().numpy(), fn(cot, th, X, X, True, iid)) <= 1e-5


def test_tf_binding_other_layouts(bridge, orc):
    import torch

    rng = np.random.default_rng(13)
    th = np.linspace(0, np.pi, 7, endpoint=False)
    vol = rng.random((20, 20, 3), dtype=np.float32)           # [X,Y,Z], tomopy_forward_compare.py:52
    got = bridge.project_tf_fast(torch.from_numpy(vol).cuda(), th, pad=True)
    want = np.transpose(orc.forward(np.transpose(vol, (2, 0, 1)).copy(), th, True, 0), (1, 2, 0))
    assert tuple(got.shape) == want.shape and rel_l2(got.cpu().numpy(), want) <= 1e-5
    one = bridge.project_tf_fast(torch.from_numpy(vol[:, :, 0].copy()).cuda(), th, pad=False, dim=2)   # toy_mcmc layout
    assert tuple(one.shape) == (7, 20, 1)
    assert rel_l2(one[..., 0].cpu().numpy(), orc.forward(vol[None, :, :, 0].copy(), th, False, 0)[0]) <= 1e-5
