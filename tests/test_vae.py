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


def _np_positive_range(x, eps):
    xm = x - 1
    return np.where(xm < 0, np.exp(np.clip(xm, -1e10, 10)) + eps, xm + 1)


@pytest.mark.gpu
def test_iradon_all_matches_oracle_restatement(orc):
    """helper_functions.py:491-519 (dose-normalise the masked sinogram, reconstruct it; back-project the mask without
    a filter), with the oracle's float64 iradon in tomopy's place: values, not just shapes."""
    rng = np.random.default_rng(5)
    N, X, A = 5, 32, 24
    theta = np.linspace(0, np.pi, A, endpoint=False)
    P = orc.frame_of(X, X, True)[1]
    eps = float(np.finfo(np.float32).eps)
    masks = np.zeros((N, A), np.float32)
    for i in range(N):
        masks[i, rng.permutation(A)[:6]] = 1.0 / 6.0
    meas = (rng.random((N, A, P), dtype=np.float32) * masks[:, :, None]).astype(np.float32)
    got = vae.iradon_all(torch.from_numpy(meas).cuda(), torch.from_numpy(masks).cuda(), theta, X, X).cpu().numpy()
    m = np.repeat(masks[:, :, None], P, axis=2).astype(np.float64)
    ps = np.where(m > eps, meas / np.maximum(m, eps), meas)
    want0 = orc.iradon(ps, theta, X, X, orc.get_fourier_filter(P, "ramp"))
    want1 = orc.iradon(m, theta, X, X, orc.get_fourier_filter(P, None))
    assert got.shape == (N, 2, X, X)
    rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))  # noqa: E731
    assert rel(got[:, 0], want0) <= 1e-5 and rel(got[:, 1], want1) <= 1e-5


@pytest.mark.gpu
def test_elbo_matches_independent_restatement(orc):
    """find_loss_vae_unsup (helper_functions.py:204-332) against a float64 NumPy / SciPy restatement of its
    arithmetic on the SAME draws: the posterior and truncated-normal samples are redrawn with the same seed and
    call order, everything downstream -- positive_range, the truncated-normal and normal log-densities, the KL terms
    and log p(M|R) through the ORACLE projector (nearest, the reference's mode) at the gathered angles -- is
    recomputed independently."""
    from scipy import stats

    torch.manual_seed(0)
    dev = torch.device("cuda")
    N, X, A, api, ns = 3, 32, 24, 7, 2
    theta = np.linspace(0, np.pi, A, endpoint=False)
    eps = float(np.finfo(np.float32).eps)
    pnm = 1e3
    imgs = torch.rand((N, X, X), device=dev) * 0.5
    sino = vae.create_sinogram(imgs, theta, pad=True, interpolation="nearest")
    masks, meas = vae.create_all_masks(sino, A, pnm, num_sparse_angles=8, random=True)
    enc_in = vae.iradon_all(meas, masks, theta, X, X)
    model = vae.CTVAE(X, X, num_filters=1, num_blocks=2).to(dev).eval()
    angles_i = torch.randperm(A)[:api]
    with torch.no_grad():
        torch.manual_seed(123)
        loss, out_dists, kl, loglik = vae.find_loss_vae_unsup(meas, masks, enc_in, model.encode, model.decode, pnm, eps,
                                                              kl_anneal=0.7, kl_multiplier=1.3, num_samples=ns, theta=theta,
                                                              angles_i=angles_i, pad=True, interpolation="nearest")
        # ---- the same draws, independent arithmetic
        torch.manual_seed(123)
        skips = model.encode(enc_in / 300)
        locs = [sv.chunk(2, dim=1)[0].double().cpu().numpy() for sv in skips]
        scales = [_np_positive_range(sv.chunk(2, dim=1)[1].double().cpu().numpy(), eps) + eps for sv in skips]
        q = [torch.distributions.Normal(sv.chunk(2, dim=1)[0], vae.positive_range(sv.chunk(2, dim=1)[1]) + eps) for sv in skips]
        lp = []
        idx = angles_i.numpy()
        th_sub = theta[idx].astype(np.float32).astype(np.float64)
        m_np, y_np = masks.cpu().numpy(), meas.cpu().numpy()
        for _ in range(ns):
            z = [d.rsample() for d in q]
            alpha, beta = model.decode(z)
            a_np = _np_positive_range(alpha.double().cpu().numpy(), eps)
            b_np = _np_positive_range(beta.double().cpu().numpy(), eps)
            x = vae.TruncatedNormal(vae.positive_range(alpha), vae.positive_range(beta), 0.0, 1e10).sample()
            x_np = x.double().cpu().numpy()
            lp_R = stats.truncnorm.logpdf(x_np, (0.0 - a_np) / b_np, (1e10 - a_np) / b_np, loc=a_np, scale=b_np)
            logp, _ = orc.log_prob_M_given_R(x_np[:, 0].astype(np.float32), m_np, y_np, pnm, eps, theta, idx, True, 0)
            lp.append(logp.sum(axis=(1, 2)) + lp_R.reshape(N, -1).sum(axis=1))
        kl_np = sum((np.log(1.0 / s) + (s ** 2 + l ** 2) / 2.0 - 0.5).reshape(N, -1).sum(axis=1) for l, s in zip(locs[1:], scales[1:]))
        # helper_functions.py:305-306 sums the log-likelihood over the batch axis too (axis=[0,1,2]): one scalar per
        # posterior sample, averaged over the samples (:329); the KL stays per image (:325) and the loss broadcasts
        want_ll = np.mean([v.sum() for v in lp])
        want_loss = 0.7 * 1.3 * kl_np - want_ll
    assert np.allclose(kl.cpu().numpy(), kl_np, rtol=2e-4)
    assert np.allclose(loglik.cpu().numpy(), want_ll, rtol=2e-4), (loglik.cpu().numpy(), want_ll)
    assert np.allclose(loss.cpu().numpy(), want_loss, rtol=2e-4)
    assert th_sub.shape == (api,)


@pytest.mark.gpu
def test_graphed_training_step():
    """The whole iteration as one CUDA graph (vae.GraphedTrainStep): replays with fresh example indices and angle
    minibatches, no plan is created, the weights move and the loss stays finite and goes down on a fixed tiny problem."""
    from ct_pvae_b200 import _lib

    torch.manual_seed(0)
    dev = torch.device("cuda")
    N, X, A, b, api = 8, 32, 24, 4, 6
    theta = np.linspace(0, np.pi, A, endpoint=False)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, X), torch.linspace(-1, 1, X), indexing="ij")
    imgs = ((xx ** 2 + yy ** 2) < 0.6).float()[None].repeat(N, 1, 1).to(dev) * (0.3 + 0.5 * torch.rand(N, 1, 1, device=dev))
    sino = vae.create_sinogram(imgs, theta, pad=True)
    masks, meas = vae.create_all_masks(sino, A, 1e4, num_sparse_angles=8, random=True)
    enc_in = vae.iradon_all(meas, masks, theta, X, X)
    model = vae.CTVAE(X, X, num_filters=1, num_blocks=2, learning_rate=1e-3).to(dev)
    before = [p.detach().clone() for p in model.parameters()]
    step = vae.GraphedTrainStep(model, meas, masks, enc_in, 1e4, theta, batch=b, angles_per_iter=api, num_samples=2)
    made = _lib.PLANS_CREATED
    g = torch.Generator().manual_seed(2)
    losses = []
    for it in range(120):
        loss = step(torch.randint(0, N, (b,), generator=g), torch.randperm(A, generator=g)[:api])
        losses.append(float(loss))
    assert _lib.PLANS_CREATED == made
    assert all(math.isfinite(v) for v in losses)
    assert any(not torch.equal(a, c) for a, c in zip(before, model.parameters()))
    assert np.mean(losses[-20:]) < np.mean(losses[:20])
