"""The restated callers (ct_pvae_b200/vae.py): network shapes and
distributions on CPU, one training step on the GPU through the fused projector."""
import math

import numpy as np
import pytest
import torch

from ct_pvae_b200 import vae


def test_positive_range_matches_reference_formula():
    x = torch.tensor([-3.0, 0.0, 0.999, 1.0, 2.5])
    eps = float(np.finfo(np.float32).eps)
    want = torch.where(x - 1 < 0, torch.exp(x - 1) + eps, x)
    assert torch.allclose(vae.positive_range(x), want)
    assert (vae.positive_range(torch.randn(1000) * 5) > 0).all()


def test_network_shapes_follow_models_py():
    # defaults of main_ct_vae.py: nfm 20, nfmm 1.1, nb 3, ks 4, se 2, il 2, ik 4; fmm 2 (non-deterministic)
    m = vae.CTVAE(32, 32, num_filters=1)
    x = torch.randn(2, 2, 32, 32)
    skips = m.encode(x)
    assert [tuple(s.shape) for s in skips] == [(2, 4, 32, 32), (2, 40, 16, 16), (2, 44, 8, 8), (2, 48, 4, 4)]
    alpha, beta = m.decode([s.chunk(2, dim=1)[0] for s in skips])
    assert alpha.shape == beta.shape == (2, 1, 32, 32)
    # odd sizes survive the periodic padding / crop arithmetic
    m2 = vae.CTVAE(30, 22, num_filters=1, num_blocks=2)
    sk = m2.encode(torch.randn(1, 2, 30, 22))
    a2, _ = m2.decode([s.chunk(2, dim=1)[0] for s in sk])
    assert a2.shape == (1, 1, 30, 22)


def test_periodic_padding_wraps():
    x = torch.arange(12.0).reshape(1, 1, 3, 4)
    p = vae.periodic_padding(x, (1, 1), (2, 0))
    assert p.shape == (1, 1, 5, 6)
    assert torch.equal(p[0, 0, 0, 2:], x[0, 0, -1]) and torch.equal(p[0, 0, 1:4, :2], x[0, 0, :, -2:])


def test_truncated_normal():
    torch.manual_seed(0)
    loc, scale = torch.full((20000,), 0.3), torch.full((20000,), 0.5)
    d = vae.TruncatedNormal(loc, scale, 0.0, 1e10)
    s = d.sample()
    assert (s >= 0).all() and abs(float(s.mean()) - float(d.mean()[0])) < 0.02
    xs = torch.linspace(0, 6, 6001)
    dd = vae.TruncatedNormal(torch.full_like(xs, 0.3), torch.full_like(xs, 0.5))
    assert abs(float(torch.trapz(dd.log_prob(xs).exp(), xs)) - 1) < 1e-3


def test_create_all_masks_uniform_and_random():
    s = torch.rand(6, 12, 8)
    m, ps = vae.create_all_masks(s, 12, 1e3, num_sparse_angles=4)
    assert m.shape == (6, 12) and torch.allclose(m.sum(dim=1), torch.ones(6))
    assert torch.equal((m[0] > 0).nonzero().flatten(), torch.tensor([0, 3, 6, 9]))
    assert ps.shape == s.shape and float(ps[:, 1].abs().max()) == 0.0
    g = torch.Generator().manual_seed(0)
    mr, _ = vae.create_all_masks(s, 12, 1e3, num_sparse_angles=4, random=True, generator=g)
    assert ((mr > 0).sum(dim=1) == 4).all()


@pytest.mark.gpu
def test_training_step_runs_through_the_fused_projector():
    torch.manual_seed(0)
    dev = torch.device("cuda")
    N, X, A = 4, 32, 24
    theta = np.linspace(0, np.pi, A, endpoint=False)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, X), torch.linspace(-1, 1, X), indexing="ij")
    imgs = ((xx ** 2 + yy ** 2) < 0.6).float()[None].repeat(N, 1, 1).to(dev) * torch.rand(N, 1, 1, device=dev)
    sino = vae.create_sinogram(imgs, theta, pad=True)
    masks, meas = vae.create_all_masks(sino, A, 1e4, num_sparse_angles=6, random=True)
    enc_in = vae.iradon_all(meas, masks, theta, X, X)
    assert enc_in.shape == (N, 2, X, X) and torch.isfinite(enc_in).all()
    model = vae.CTVAE(X, X, num_filters=1).to(dev)
    before = [p.detach().clone() for p in model.parameters()]
    losses = []
    for it in range(3):
        angles_i = torch.randperm(A)[:8]
        loss, _, kl, ll = model.train_step(meas, masks, enc_in, 1e4, theta, angles_i=angles_i, num_samples=2)
        losses.append(float(loss))
    assert all(math.isfinite(v) for v in losses)
    assert any(not torch.equal(a, b) for a, b in zip(before, model.parameters()))
