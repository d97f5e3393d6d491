"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle.

Tolerance (north_star): relative L2 <= 1e-5 for the float32 projector and its
adjoints; <Ax,y> == <x,A^T y> to float32 rounding.  FBP computes in float32 against a
float64 oracle: relative L2 <= 1e-5 as well.
"""
import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu

TOL = 1e-5
INTERPS = ["nearest", "bilinear"]
IID = {"nearest": 0, "bilinear": 1}

# (B, X, Y, A, pad): ragged batches (B % 4, B % 8 != 0), non-square, no-pad, tiny, multi-strip
SHAPES = [
    (1, 2, 2, 2, False),
    (5, 16, 16, 12, True),
    (3, 16, 16, 12, False),
    (2, 33, 20, 17, True),
    (2, 20, 33, 17, False),
    (9, 64, 64, 24, True),
    (4, 128, 128, 20, True),
    (1, 200, 200, 7, True),
    # B >= 12 with a detector of <= 256 bins: the depth-first forward (16 images per pixel record)
    (13, 16, 16, 12, True),
    (17, 33, 20, 9, True),
    (16, 64, 64, 8, False),
    (37, 128, 128, 6, True),
]


def _theta(A):
    return np.linspace(0, np.pi, A, endpoint=False)


@pytest.fixture(scope="module")
def cp():
    import ct_pvae_b200

    assert torch.cuda.is_available()
    return ct_pvae_b200


@pytest.mark.parametrize("interp", INTERPS)
@pytest.mark.parametrize("B,X,Y,A,pad", SHAPES)
def test_forward_matches_oracle(cp, orc, B, X, Y, A, pad, interp):
    rng = np.random.default_rng(0)
    img = rng.random((B, X, Y), dtype=np.float32)
    th = _theta(A) if (X, Y) != (2, 2) else np.array([0, np.pi / 2])
    want = orc.forward(img, th, pad, IID[interp])
    got = cp.project_tf_fast(torch.from_numpy(img).cuda().unsqueeze(-1), th, pad=pad, dim=2, integrate_vae=True,
                             interpolation=interp)
    assert got.shape == (B, len(th), want.shape[2], 1)
    assert rel_l2(got[..., 0].cpu().numpy(), want) <= TOL


@pytest.mark.parametrize("mode", ["exact", "tf_compat"])
@pytest.mark.parametrize("interp", INTERPS)
@pytest.mark.parametrize("B,X,Y,A,pad", SHAPES)
def test_adjoint_matches_oracle(cp, orc, B, X, Y, A, pad, interp, mode):
    rng = np.random.default_rng(1)
    th = _theta(A) if (X, Y) != (2, 2) else np.array([0, np.pi / 2])
    W = orc.frame_of(X, Y, pad)[1]
    y = rng.random((B, len(th), W), dtype=np.float32)
    fn = orc.adjoint_exact if mode == "exact" else orc.adjoint_tf
    want = fn(y, th, X, Y, pad, IID[interp])
    got = cp.backproject(torch.from_numpy(y).cuda(), th, X, Y, pad=pad, interpolation=interp, adjoint=mode)
    assert got.shape == (B, X, Y)
    assert rel_l2(got.cpu().numpy(), want) <= TOL


@pytest.mark.parametrize("B,X,Y,A,pad", [(1, 1024, 1024, 3, True), (2, 300, 700, 5, False), (3, 700, 300, 5, True)])
def test_large_and_rectangular_frames(cp, orc, B, X, Y, A, pad):
    """Detector wider than one CTA (P = 1452 -> two detector chunks), wide strips, non-square frames."""
    rng = np.random.default_rng(7)
    img = rng.random((B, X, Y), dtype=np.float32)
    th = np.array([0.3, 1.1, 2.5, 0.0, np.pi / 2])[:A]
    W = orc.frame_of(X, Y, pad)[1]
    y = rng.random((B, A, W), dtype=np.float32)
    for interp in INTERPS:
        got = cp.project_tf_fast(torch.from_numpy(img).cuda().unsqueeze(-1), th, pad=pad, dim=2, integrate_vae=True,
                                 interpolation=interp)[..., 0]
        assert rel_l2(got.cpu().numpy(), orc.forward(img, th, pad, IID[interp])) <= TOL
        g = cp.backproject(torch.from_numpy(y).cuda(), th, X, Y, pad=pad, interpolation=interp)
        assert rel_l2(g.cpu().numpy(), orc.adjoint_exact(y, th, X, Y, pad, IID[interp])) <= TOL


@pytest.mark.parametrize("B,X,Y,A,pad,kind", [(16, 300, 300, 24, True, "even"), (20, 200, 420, 10, False, "random"),
                                              (33, 260, 260, 16, True, "even"), (12, 512, 512, 6, True, "random")])
def test_windowed_forward_matches_oracle(cp, orc, B, X, Y, A, pad, kind):
    """Wide detector + batch >= 12: 16-image records with column-windowed strips (per-CTA windows that follow
    the rays).  Far-apart random angles force wide / whole-row windows or the 4-image fallback."""
    from ct_pvae_b200 import _lib
    rng = np.random.default_rng(70 + B)
    th = _theta(A) if kind == "even" else rng.uniform(-4, 4, A)
    img = rng.random((B, X, Y), dtype=np.float32)
    desc = _lib.get_plan(np.asarray(th, np.float64), X, Y, pad, 0).describe(B)
    if kind == "even":
        assert ("images_per_record=32 windowed=1" if B > 16 else "images_per_record=16 windowed=1") in desc, desc
        assert "window_chunks=0/" not in desc, desc
    for interp in INTERPS:
        got = cp.project_tf_fast(torch.from_numpy(img).cuda().unsqueeze(-1), th, pad=pad, dim=2, integrate_vae=True,
                                 interpolation=interp)[..., 0]
        assert rel_l2(got.cpu().numpy(), orc.forward(img, th, pad, IID[interp])) <= TOL, desc


@pytest.mark.parametrize("seed", range(8))
def test_randomised_shapes(cp, orc, seed):
    """Random batch / image / angle-set shapes (ragged against every internal group size)."""
    rng = np.random.default_rng(100 + seed)
    B, X, Y, A = int(rng.integers(1, 41)), int(rng.integers(1, 90)), int(rng.integers(1, 90)), int(rng.integers(1, 14))
    pad = bool(rng.integers(0, 2))
    th = rng.uniform(-4, 4, A)
    img = rng.random((B, X, Y), dtype=np.float32)
    W = orc.frame_of(X, Y, pad)[1]
    y = rng.random((B, A, W), dtype=np.float32)
    for interp in INTERPS:
        got = cp.project_tf_fast(torch.from_numpy(img).cuda().unsqueeze(-1), th, pad=pad, dim=2, integrate_vae=True,
                                 interpolation=interp)[..., 0].cpu().numpy()
        want = orc.forward(img, th, pad, IID[interp])
        assert rel_l2(got, want) <= TOL
        assert np.abs(got - want).max() <= 2e-5 * max(1.0, np.abs(want).max())
        for mode, fn in (("exact", orc.adjoint_exact), ("tf_compat", orc.adjoint_tf)):
            g = cp.backproject(torch.from_numpy(y).cuda(), th, X, Y, pad=pad, interpolation=interp, adjoint=mode).cpu().numpy()
            gw = fn(y, th, X, Y, pad, IID[interp])
            assert rel_l2(g, gw) <= TOL
            assert np.abs(g - gw).max() <= 2e-5 * max(1.0, np.abs(gw).max())


def test_random_angles_and_signs(cp, orc):
    rng = np.random.default_rng(2)
    th = rng.uniform(-7, 7, 23)
    img = rng.random((3, 40, 40), dtype=np.float32)
    for interp in INTERPS:
        want = orc.forward(img, th, True, IID[interp])
        got = cp.project_tf_fast(torch.from_numpy(img).cuda().unsqueeze(-1), th, pad=True, dim=2, integrate_vae=True,
                                 interpolation=interp)
        assert rel_l2(got[..., 0].cpu().numpy(), want) <= TOL


def test_toy_known_answers(cp):
    # scripts/create_toy_images.py:36-37 + scripts/images_to_sinograms.py:54-59
    x0 = np.array([[1, 2], [3, 4]], np.float32) / 10
    th = np.array([0, np.pi / 2])
    for interp in INTERPS:
        s = cp.project_tf_fast(torch.from_numpy(x0).cuda(), th, pad=False, dim=2, integrate_vae=False,
                               interpolation=interp)
        assert s.shape == (2, 2, 1)
        np.testing.assert_allclose(s[..., 0].cpu().numpy(), [[0.4, 0.6], [0.7, 0.3]], rtol=1e-6)


def test_layouts_and_dtypes(cp, orc):
    rng = np.random.default_rng(3)
    th = _theta(9)
    vol = rng.random((24, 24, 3))  # [X,Y,Z] float64 like tomopy_forward_compare.py:52,56
    want = np.transpose(orc.forward(np.transpose(vol, (2, 0, 1)).astype(np.float32), th, True, 0), (1, 2, 0))
    got = cp.project_tf_fast(vol, th, pad=True)       # numpy in -> numpy out, dtype kept
    assert isinstance(got, np.ndarray) and got.dtype == np.float64 and got.shape == want.shape
    assert rel_l2(got, want) <= TOL
    want_b = np.transpose(orc.forward(np.transpose(vol, (2, 0, 1)).astype(np.float32), th, True, 1), (1, 2, 0))
    got_b = cp.project_tf_low_mem(torch.from_numpy(vol), th, pad=True)   # CPU tensor in -> CPU tensor out
    assert got_b.device.type == "cpu" and got_b.dtype == torch.float64
    assert rel_l2(got_b.numpy(), want_b) <= TOL
    one = cp.project_tf_fast(vol[:, :, 0].astype(np.float32), th, pad=True, dim=2)
    assert one.shape == (9, want.shape[1], 1) and rel_l2(one[..., 0], want[..., 0]) <= TOL


@pytest.mark.parametrize("interp", INTERPS)
def test_autograd_is_the_adjoint(cp, orc, interp):
    rng = np.random.default_rng(4)
    th = _theta(15)
    img = torch.from_numpy(rng.random((3, 32, 32, 1), dtype=np.float32)).cuda().requires_grad_(True)
    W = orc.frame_of(32, 32, True)[1]
    y = torch.from_numpy(rng.random((3, 15, W, 1), dtype=np.float32)).cuda()
    s = cp.project_tf_fast(img, th, pad=True, dim=2, integrate_vae=True, interpolation=interp)
    (s * y).sum().backward()
    want = orc.adjoint_exact(y[..., 0].cpu().numpy(), th, 32, 32, True, IID[interp])
    assert rel_l2(img.grad[..., 0].cpu().numpy(), want) <= TOL
    img.grad = None
    s = cp.project_tf_fast(img, th, pad=True, dim=2, integrate_vae=True, interpolation=interp, adjoint="tf_compat")
    (s * y).sum().backward()
    want = orc.adjoint_tf(y[..., 0].cpu().numpy(), th, 32, 32, True, IID[interp])
    assert rel_l2(img.grad[..., 0].cpu().numpy(), want) <= TOL


@pytest.mark.parametrize("interp", INTERPS)
@pytest.mark.parametrize("B,X,A", [(256, 128, 180), (8, 512, 720), (16, 512, 720)])
def test_full_size_properties(cp, B, X, A, interp):
    """BASELINE.json sizes (C2: 256x128^2x180, C4 slice: 512^2x720): size-independent checks."""
    g = torch.Generator(device="cuda").manual_seed(5)
    th = _theta(A)
    img = torch.rand((B, X, X), device="cuda", generator=g)
    s = cp.project_tf_fast(img.unsqueeze(-1), th, pad=True, dim=2, integrate_vae=True, interpolation=interp)[..., 0]
    P = s.shape[2]
    assert P == cp.num_proj_pix(X, X)
    # theta = 0: exact column sums, centred in the detector
    pady = (P - X) // 2
    col = img.double().sum(dim=1)
    assert torch.allclose(s[:, 0, pady:pady + X].double(), col, rtol=1e-5, atol=0)
    assert float(s[:, 0, :pady].abs().max()) == 0.0 and float(s[:, 0, pady + X:].abs().max()) == 0.0
    if interp == "bilinear":
        # mass is conserved up to the rotated lattice not being an exact partition of
        # unity for the bilinear hat (oracle: <= 1e-3 on random 128^2 images)
        mass = img.double().sum(dim=(1, 2))
        assert torch.allclose(s.double().sum(dim=2), mass[:, None].expand(-1, A), rtol=3e-3)
    # linearity
    img2 = torch.rand((B, X, X), device="cuda", generator=g)
    s2 = cp.project_tf_fast(img2.unsqueeze(-1), th, pad=True, dim=2, integrate_vae=True, interpolation=interp)[..., 0]
    s12 = cp.project_tf_fast((img + 2 * img2).unsqueeze(-1), th, pad=True, dim=2, integrate_vae=True,
                             interpolation=interp)[..., 0]
    assert float((s12 - (s + 2 * s2)).norm() / s12.norm()) <= 2e-6
    # adjoint identity <Ax,y> = <x,A^T y>
    y = torch.rand(s.shape, device="cuda", generator=g)
    gimg = cp.backproject(y, th, X, X, pad=True, interpolation=interp, adjoint="exact")
    lhs = float((s.double() * y.double()).sum())
    rhs = float((img.double() * gimg.double()).sum())
    assert abs(lhs - rhs) / abs(lhs) <= 2e-6


# ---- BASELINE.json sizes against the oracle itself (not only through properties) -----------------------------
# configs[1] (C2): all 256 images x 128^2 x 180 angles; configs[3] (C4): an 8-image slice of 512^2 x 720 angles (the
# OpenMP oracle does either in seconds), plus a 32-image slice at 3 angles that runs the production shapes of C4
# (32-image pixel records, windowed strips, reuse march; 32 images per adjoint thread for both adjoints).
FULL = [("c2", 256, 128, 180), ("c4_slice", 8, 512, 720), ("c4_groups", 32, 512, 3)]


@pytest.mark.parametrize("interp", INTERPS)
@pytest.mark.parametrize("name,B,X,A", FULL)
def test_full_size_forward_matches_oracle(cp, orc, name, B, X, A, interp):
    rng = np.random.default_rng(31)
    th = _theta(720)[::240] if name == "c4_groups" else _theta(A)     # 0, 60, 120 degrees: both ray classes
    img = rng.random((B, X, X), dtype=np.float32)
    got = cp.project_tf_fast(torch.from_numpy(img).cuda().unsqueeze(-1), th, pad=True, dim=2, integrate_vae=True,
                             interpolation=interp)[..., 0].cpu().numpy()
    assert rel_l2(got, orc.forward(img, th, True, IID[interp])) <= TOL


@pytest.mark.parametrize("interp", INTERPS)
def test_c4_angle_block_forward_matches_oracle(cp, orc, interp):
    """16 neighbouring angles of configs[3]'s 720 (two blocks, one per ray class) x 40 images (a full and a ragged
    32-image record): the production CTA shapes of C4 with every angle slot filled -- 8-image lanes + reuse march for
    bilinear, two lanes per ray x 16 images with rotated loads (8 angle slots) for nearest."""
    from ct_pvae_b200 import _lib
    rng = np.random.default_rng(33)
    X, B = 512, 40
    th = np.concatenate([_theta(720)[100:108], _theta(720)[300:308]])
    desc = _lib.get_plan(np.asarray(th, np.float64), X, X, True, 0).describe(B)
    assert "images_per_record=32 windowed=1" in desc and "nearest: images_per_lane=16" in desc, desc
    img = rng.random((B, X, X), dtype=np.float32)
    got = cp.project_tf_fast(torch.from_numpy(img).cuda().unsqueeze(-1), th, pad=True, dim=2, integrate_vae=True,
                             interpolation=interp)[..., 0].cpu().numpy()
    assert rel_l2(got, orc.forward(img, th, True, IID[interp])) <= TOL


@pytest.mark.parametrize("mode", ["exact", "tf_compat"])
@pytest.mark.parametrize("interp", INTERPS)
@pytest.mark.parametrize("name,B,X,A", FULL)
def test_full_size_adjoint_matches_oracle(cp, orc, name, B, X, A, interp, mode):
    from ct_pvae_b200 import _lib
    rng = np.random.default_rng(32)
    th = _theta(720)[::240] if name == "c4_groups" else _theta(A)
    W = orc.frame_of(X, X, True)[1]
    y = rng.random((B, len(th), W), dtype=np.float32)
    if name == "c4_groups":   # the dispatch this case exists for (the 32-image TF-compat path had no value check)
        desc = _lib.get_plan(np.asarray(th, np.float64), X, X, True, 0).describe(B)
        assert "exact=32 tf_compat=32" in desc, desc
    fn = orc.adjoint_exact if mode == "exact" else orc.adjoint_tf
    got = cp.backproject(torch.from_numpy(y).cuda(), th, X, X, pad=True, interpolation=interp, adjoint=mode).cpu().numpy()
    assert rel_l2(got, fn(y, th, X, X, True, IID[interp])) <= TOL


def test_fbp_matches_oracle(cp, orc):
    rng = np.random.default_rng(6)
    # ramp / None: odd-tap row filter (P = 256 with 5 images: four outputs per thread, 8-image groups; P = 184 with 40
    # images: three outputs per thread; 500 images: 16-image filter groups inside the gather's 32-image groups, last one ragged)
    for (B, A, P, xs, ys, name) in [(3, 20, 46, 30, 30, "ramp"), (9, 45, 184, 128, 128, "hann"),
                                    (2, 12, 50, 33, 21, None), (1, 7, 24, 16, 16, "shepp-logan"),
                                    (5, 9, 256, 179, 179, "ramp"), (40, 11, 184, 128, 128, "ramp"),
                                    (500, 4, 184, 128, 128, "ramp")]:   # >= 1024 gather CTAs: 32 images per gather thread
        th = _theta(A)
        sino = rng.random((B, A, P))
        filt = orc.get_fourier_filter(P, name)
        np.testing.assert_allclose(cp.get_fourier_filter(P, name), filt)
        want = orc.iradon(sino, th, xs, ys, filt)
        got = cp.iradon(torch.from_numpy(sino).cuda(), th, xs, ys, filt)
        assert got.dtype == torch.float64 and got.shape == (B, xs, ys)
        assert rel_l2(got.cpu().numpy(), want) <= TOL
    with pytest.raises(ValueError):
        cp.iradon(torch.zeros((1, 5, 16), device="cuda"), _theta(4), 8, 8, np.ones(16))


@pytest.mark.parametrize("seed", range(10))
def test_fbp_random_shapes(cp, orc, seed):
    """iradon on random detector widths (even and odd: odd-tap and dense row filters, float4 and scalar staging), batch
    sizes (8 / 16-image filter groups, ragged), image sizes and angle sets, every filter, against the float64 oracle."""
    rng = np.random.default_rng(700 + seed)
    B, A = int(rng.integers(1, 45)), int(rng.integers(1, 25))
    xs, ys = int(rng.integers(2, 120)), int(rng.integers(2, 120))
    P = int(rng.integers(2, 200)) if seed % 3 == 0 else cp.num_proj_pix(xs, ys)   # any width, or the reference's own
    name = ["ramp", None, "hann", "ramp", "shepp-logan", "ramp", "cosine", "ramp", "hamming", "ramp"][seed] if P % 2 == 0 else None
    th = rng.uniform(-4, 4, A) if seed % 2 else _theta(A)
    sino = rng.random((B, A, P))
    filt = orc.get_fourier_filter(P, name) if P % 2 == 0 else np.ones(P)   # skimage's filters need an even size
    want = orc.iradon(sino, th, xs, ys, filt)
    got = cp.iradon(torch.from_numpy(sino).cuda(), th, xs, ys, filt).cpu().numpy()
    assert rel_l2(got, want) <= TOL, (B, A, P, xs, ys, name)


@pytest.mark.parametrize("B,A,X,Y", [(3, 20, 30, 30), (9, 45, 128, 128), (2, 12, 33, 21), (20, 13, 64, 64), (17, 9, 90, 90),
                                     (5, 7, 128, 100), (33, 24, 47, 45)])
def test_fused_fbp_is_one_kernel_and_matches_the_two_kernel_path(cp, orc, B, A, X, Y):
    """iradon of images up to 128 x 128: ONE cluster kernel (row filter in shared memory, back-projection from
    distributed shared memory).  Cluster sizes 1 / 2 / 4 / 8, ragged 16-image groups, angle counts that do not fill
    the last batch.  Bit-identical to the filter + gather kernel pair, and within 1e-5 of the float64 oracle."""
    from ct_pvae_b200 import _lib, ops
    rng = np.random.default_rng(50 + B)
    P = cp.num_proj_pix(X, Y)
    th = _theta(A)
    sino = rng.random((B, A, P), dtype=np.float32)
    filt = orc.get_fourier_filter(P, "ramp")
    plan = _lib.get_fbp_plan(th, P, X, Y, filt, 0)
    x = torch.from_numpy(sino).cuda()
    try:
        assert plan.set_fused(True)
        _lib.profile_reset()
        _lib.profile_enable(True)
        fused = ops.fbp(x, plan)
        torch.cuda.synchronize()
        prof = _lib.profile_read()
        _lib.profile_enable(False)
        assert list(prof) == ["ctr_fbp_fused_kernel"] and prof["ctr_fbp_fused_kernel"][1] == 1, prof
        plan.set_fused(False)
        two = ops.fbp(x, plan)
    finally:
        _lib.profile_enable(False)
        _lib.profile_reset()
        plan.set_fused(False)                            # library default
    assert torch.equal(fused, two)
    assert rel_l2(fused.cpu().numpy(), orc.iradon(sino.astype(np.float64), th, X, Y, filt)) <= TOL


def test_large_images_take_the_two_kernel_fbp(cp, orc):
    from ct_pvae_b200 import _lib
    rng = np.random.default_rng(51)
    B, A, X = 2, 10, 200
    P = cp.num_proj_pix(X, X)
    th = _theta(A)
    sino = rng.random((B, A, P))
    filt = orc.get_fourier_filter(P, "hann")
    assert not _lib.get_fbp_plan(th, P, X, X, filt, 0).set_fused(True)      # 40000 pixels > 8 x 2048
    got = cp.iradon(torch.from_numpy(sino).cuda(), th, X, X, filt)
    assert rel_l2(got.cpu().numpy(), orc.iradon(sino, th, X, X, filt)) <= TOL


def test_fbp_reconstructs_disk(cp):
    """forward (bilinear) -> iradon(ramp) recovers a 0.8-valued disk."""
    X = 128
    yy, xx = np.meshgrid(np.arange(X) - X / 2 + 0.5, np.arange(X) - X / 2 + 0.5, indexing="ij")
    img = (0.8 * ((xx ** 2 + yy ** 2) <= 40 ** 2)).astype(np.float32)
    th = _theta(180)
    s = cp.project_tf_fast(torch.from_numpy(img).cuda()[None, ..., None], th, pad=True, dim=2, integrate_vae=True,
                           interpolation="bilinear")[..., 0]
    P = s.shape[2]
    rec = cp.iradon(s, th, X, X, cp.get_fourier_filter(P, "ramp"))[0].cpu().numpy()
    inner = (xx ** 2 + yy ** 2) <= 30 ** 2
    assert abs(rec[inner].mean() - 0.8) < 0.01


def test_pinned_host_batch_goes_through_the_chunked_pipeline(cp, orc):
    """B >= 32 pinned float32 host batch: copy-in / kernels / copy-out overlap chunk by chunk."""
    rng = np.random.default_rng(8)
    B, X, A = 70, 24, 9
    th = _theta(A)
    img = rng.random((B, X, X), dtype=np.float32)
    img_h = torch.from_numpy(img).unsqueeze(-1).pin_memory()
    s = cp.project_tf_fast(img_h, th, pad=True, dim=2, integrate_vae=True, interpolation="bilinear")
    assert s.device.type == "cpu" and s.is_pinned() and s.shape[:2] == (B, A)
    want = orc.forward(img, th, True, 1)
    assert rel_l2(s[..., 0].numpy(), want) <= TOL
    cot = rng.random(want.shape, dtype=np.float32)
    g = cp.backproject(torch.from_numpy(cot).pin_memory(), th, X, X, pad=True, interpolation="bilinear")
    assert g.device.type == "cpu" and rel_l2(g.numpy(), orc.adjoint_exact(cot, th, X, X, True, 1)) <= TOL


@pytest.fixture
def host_chunks():
    from ct_pvae_b200 import hostpipe
    yield hostpipe.set_chunk
    hostpipe.set_chunk(0, 0)


@pytest.mark.parametrize("chunk", [0, 32, 16])
def test_async_host_calls_overlap_and_match(cp, orc, chunk, host_chunks):
    """async_op=True: two host-buffer calls in flight at once (result + completion handle each); forced small
    chunks exercise the staging ring's slot reuse and the ragged last chunk."""
    host_chunks(chunk, chunk)
    rng = np.random.default_rng(18)
    B, X, A = 120, 24, 9
    th = _theta(A)
    img = rng.random((B, X, X), dtype=np.float32)
    want = orc.forward(img, th, True, 1)
    cot = rng.random(want.shape, dtype=np.float32)
    img_h, cot_h = torch.from_numpy(img).unsqueeze(-1).pin_memory(), torch.from_numpy(cot).pin_memory()
    for _ in range(3):
        s, hs = cp.project_tf_fast(img_h, th, pad=True, dim=2, integrate_vae=True, interpolation="bilinear", async_op=True)
        g, hg = cp.backproject(cot_h, th, X, X, pad=True, interpolation="bilinear", async_op=True)
        hs.wait()
        hg.wait()
        assert hs.is_completed() and hg.is_completed()
        assert s.is_pinned() and rel_l2(s[..., 0].numpy(), want) <= TOL
        assert rel_l2(g.numpy(), orc.adjoint_exact(cot, th, X, X, True, 1)) <= TOL
    with pytest.raises(ValueError):
        cp.project_tf_fast(torch.from_numpy(img[:4]).unsqueeze(-1), th, pad=True, dim=2, integrate_vae=True, async_op=True)


@pytest.mark.parametrize("chunk", [0, 16])
def test_pageable_numpy_batches_are_staged_by_the_library(cp, orc, chunk, host_chunks):
    """The drop-in case: plain (pageable) NumPy arrays, what the reference's callers pass.  The library stages them
    through the pipe's pinned ring; blocking and async_op calls, several in flight, small chunks wrapping the ring,
    and a caller-provided result buffer (out=)."""
    host_chunks(chunk, chunk)
    rng = np.random.default_rng(28)
    B, X, A = 150, 24, 9
    th = _theta(A)
    img = rng.random((B, X, X, 1), dtype=np.float32)
    want = orc.forward(img[..., 0], th, True, 1)
    cot = rng.random(want.shape, dtype=np.float32)
    gwant = orc.adjoint_exact(cot, th, X, X, True, 1)
    s = cp.project_tf_fast(img, th, pad=True, dim=2, integrate_vae=True, interpolation="bilinear")
    assert isinstance(s, np.ndarray) and s.dtype == np.float32 and rel_l2(s[..., 0], want) <= TOL
    g = cp.backproject(cot, th, X, X, pad=True, interpolation="bilinear")
    assert isinstance(g, np.ndarray) and rel_l2(g, gwant) <= TOL
    out_s = torch.empty((B, A, want.shape[2]), dtype=torch.float32).pin_memory()
    for _ in range(3):
        s2, hs = cp.project_tf_fast(img, th, pad=True, dim=2, integrate_vae=True, interpolation="bilinear", async_op=True, out=out_s)
        g2, hg = cp.backproject(cot, th, X, X, pad=True, interpolation="bilinear", async_op=True)
        img_copy = img.copy()
        img_copy[:] = 0          # a pageable input may be released / overwritten as soon as the call has returned
        hs.wait()
        hg.wait()
        assert hs.is_completed() and hg.is_completed()
        assert s2.ctypes.data == out_s.data_ptr() and rel_l2(out_s.numpy(), want) <= TOL
        assert rel_l2(g2, gwant) <= TOL
    # a pageable RESULT buffer: delivered by wait()
    out_pg = torch.empty((B, X, X), dtype=torch.float32)
    g3, hg = cp.backproject(cot, th, X, X, pad=True, interpolation="bilinear", async_op=True, out=out_pg)
    hg.wait()
    assert rel_l2(out_pg.numpy(), gwant) <= TOL
    with pytest.raises(ValueError):
        cp.backproject(cot, th, X, X, pad=True, interpolation="bilinear", out=torch.empty((B, X, X + 1)))


def test_host_handles_survive_pipe_eviction(cp, orc):
    """evicting / clearing a pipe that outstanding async handles still reference must not turn a completed
    result into an error."""
    from ct_pvae_b200 import hostpipe
    rng = np.random.default_rng(38)
    B, X, A = 40, 16, 5
    th = _theta(A)
    img = rng.random((B, X, X, 1), dtype=np.float32)
    s, h = cp.project_tf_fast(torch.from_numpy(img).pin_memory(), th, pad=True, dim=2, integrate_vae=True,
                              interpolation="bilinear", async_op=True)
    hostpipe.clear_pipes()
    assert h.is_completed()
    h.wait()
    assert rel_l2(s[..., 0].numpy(), orc.forward(img[..., 0], th, True, 1)) <= TOL


@pytest.mark.parametrize("interp", INTERPS)
@pytest.mark.parametrize("gather,X", [(False, 32), (True, 32), (False, 150)])
def test_fused_loglik_matches_oracle(cp, orc, interp, gather, X):
    """projector + mask + Normal log-prob + reduction in one pass, and its gradient."""
    rng = np.random.default_rng(9)
    # 18 images: 32-image records (X = 150: P = 216, column-windowed strips + reuse march); 6: 4-image records
    B, A_all = (18 if (gather or X > 100) else 6), 20
    th = _theta(A_all)
    angles_i = rng.permutation(A_all)[:7] if gather else None
    pnm, sreg = 1e4, float(np.finfo(np.float32).eps)
    img = rng.random((B, X, X), dtype=np.float32)
    P = orc.frame_of(X, X, True)[1]
    mask = (rng.random((B, A_all)) < 0.5).astype(np.float32) / 3.0
    clean = orc.forward(img, th, True, IID[interp])
    meas = (clean * mask[:, :, None] * (1 + 0.05 * rng.standard_normal(clean.shape))).astype(np.float32)
    logp, dproj = orc.log_prob_M_given_R(img, mask, meas, pnm, sreg, th, angles_i, True, IID[interp])
    x = torch.from_numpy(img).cuda().unsqueeze(-1).requires_grad_(True)
    ai = None if angles_i is None else torch.from_numpy(angles_i)
    per = cp.log_prob_M_given_R_sum(x, torch.from_numpy(mask).cuda(), torch.from_numpy(meas).cuda(), pnm, sreg, theta=th,
                                    angles_i=ai, pad=True, interpolation=interp, per_image=True)
    want = logp.sum(axis=(1, 2))
    # float32 terms of both signs: the error budget scales with sum |terms|, not with the (cancelling) sum
    budget = 2e-6 * np.abs(logp).sum(axis=(1, 2)) + 2e-5 * np.abs(want)
    assert (np.abs(per.detach().cpu().numpy() - want) <= budget).all()
    per.sum().backward()
    th_sub = th if angles_i is None else th[angles_i].astype(np.float32).astype(np.float64)
    gwant = orc.adjoint_exact(dproj.astype(np.float32), th_sub, X, X, True, IID[interp])
    assert rel_l2(x.grad[..., 0].cpu().numpy(), gwant) <= 2e-5
    # the drop-in full-tensor form agrees with the fused sum
    full = cp.calculate_log_prob_M_given_R(x.detach(), torch.from_numpy(mask).cuda(), torch.from_numpy(meas).cuda(), pnm, sreg,
                                           theta=th, angles_i=ai, pad=True, interpolation=interp)
    assert full.shape == (B, logp.shape[1], P, 1)
    assert (np.abs(full[..., 0].double().sum(dim=(1, 2)).cpu().numpy() - want) <= budget).all()


@pytest.mark.parametrize("B,X,A_all,n_sel", [(5, 128, 180, 20), (13, 33, 40, 7), (37, 64, 30, 30), (20, 300, 48, 9), (33, 260, 36, 5), (3, 200, 12, 1)])
def test_angle_subsets_share_one_plan(cp, orc, B, X, A_all, n_sel):
    """Training's angle minibatch (helper_functions.py:350-357): ONE plan over all the angles, the subset as a device
    index list (ctr_radon_*_sel).  Every forward shape (4 / 16 / 32-image records, whole-row and windowed strips) and
    both adjoints against the oracle run on the gathered angles; no plan is created after the first call."""
    from ct_pvae_b200 import _lib, ops
    rng = np.random.default_rng(200 + B)
    th = _theta(A_all)
    plan = _lib.get_plan(th, X, X, True, 0)
    img = rng.random((B, X, X), dtype=np.float32)
    x = torch.from_numpy(img).cuda()
    made = _lib.PLANS_CREATED
    for trial in range(2):
        idx = rng.permutation(A_all)[:n_sel]
        sel = torch.from_numpy(idx.astype(np.int32)).cuda()
        y = rng.random((B, n_sel, plan.W), dtype=np.float32)
        for interp in INTERPS:
            got = ops.radon_forward(x, plan, IID[interp], sel).cpu().numpy()
            assert got.shape == (B, n_sel, plan.W)
            assert rel_l2(got, orc.forward(img, th[idx], True, IID[interp])) <= TOL
            for mode, fn in ((0, orc.adjoint_exact), (1, orc.adjoint_tf)):
                g = ops.radon_adjoint(torch.from_numpy(y).cuda(), plan, IID[interp], mode, sel).cpu().numpy()
                assert rel_l2(g, fn(y, th[idx], X, X, True, IID[interp])) <= TOL
    assert _lib.PLANS_CREATED == made


def test_training_iterations_do_not_replan(cp):
    """vae.train_step with a fresh random angle minibatch every iteration (main_ct_vae.py:388-389): after the first
    iteration no ctr_plan is created (every iteration used to miss the plan cache)."""
    from ct_pvae_b200 import _lib, vae
    torch.manual_seed(0)
    N, X, A, b, api = 8, 32, 30, 4, 6
    theta = np.linspace(0, np.pi, A, endpoint=False)
    imgs = torch.rand((N, X, X), device="cuda")
    sino = vae.create_sinogram(imgs, theta, pad=True, interpolation="bilinear")
    masks, meas = vae.create_all_masks(sino, A, 1e4, num_sparse_angles=10, random=True)
    enc_in = vae.iradon_all(meas, masks, theta, X, X)
    model = vae.CTVAE(X, X, num_filters=1).cuda()
    g = torch.Generator().manual_seed(1)
    losses = []
    for it in range(6):
        if it == 1:
            made = _lib.PLANS_CREATED
        idx = torch.randint(0, N, (b,), generator=g).cuda()
        angles_i = torch.randperm(A, generator=g)[:api]
        loss, _, _, _ = model.train_step(meas[idx], masks[idx], enc_in[idx], 1e4, theta, angles_i=angles_i, num_samples=2)
        losses.append(float(loss))
    assert _lib.PLANS_CREATED == made, "a training iteration created a plan"
    assert all(np.isfinite(losses))


def test_dataset_helpers_on_the_projector(cp, orc, tmp_path):
    """scripts/images_to_sinograms.py on the B200 projector + the on-disk format round trip."""
    from ct_pvae_b200 import datasets

    rng = np.random.default_rng(11)
    imgs = rng.random((6, 20, 20), dtype=np.float32)
    th = _theta(9)
    s = datasets.images_to_sinograms(imgs, th, pad=True, save_path=str(tmp_path))
    assert rel_l2(s, orc.forward(imgs, th, True, 1)) <= TOL
    got, th2, P = datasets.get_sinograms(str(tmp_path))
    assert P == s.shape[2] and np.array_equal(th2, th) and np.array_equal(got, np.where(s < 0, 0, s))
    assert datasets.reconstruction_size(cp.num_proj_pix(128, 128)) == (128, 128)


def test_toy_mcmc_likelihood_differentiates_through_the_projector(cp, orc):
    """Third caller (ctvae/toy_mcmc_v2_functions.py:30-64): a 2x2 image, pad=False, dim=2, Poisson
    likelihood of the masked projection, differentiated by HMC -- here by autograd, checked against
    the analytic gradient A^T (d ll / d proj) built from the oracle's matrix."""
    theta = np.array([0, np.pi / 2])
    mask = np.array([1.0, 0.0], np.float32)
    pnm = 1e3
    O = torch.tensor([[0.1, 0.2], [0.3, 0.4]], device="cuda", requires_grad=True)
    M = torch.tensor([[0.41, 0.59], [0.0, 0.0]], device="cuda")
    proj = cp.project_tf_fast(O, theta, pad=False, dim=2, integrate_vae=False)[..., 0]       # [A, P]
    rate = proj * torch.from_numpy(mask).cuda()[:, None] * pnm
    ll = torch.distributions.Poisson(rate[0]).log_prob(M[0] * pnm).sum()                      # only the unmasked angle
    ll.backward()
    A = orc.build_matrix(theta, 2, 2, False, 0).toarray()                                      # [(A*P) x 4]
    p = A @ O.detach().cpu().numpy().reshape(-1)
    dll_dproj = np.zeros(4)
    dll_dproj[:2] = (M[0].cpu().numpy() * pnm / (p[:2] * pnm) - 1.0) * pnm
    want = (A.T @ dll_dproj).reshape(2, 2)
    np.testing.assert_allclose(O.grad.cpu().numpy(), want, rtol=1e-4)


def test_golden_fixtures(cp):
    import os

    d = os.path.join(os.path.dirname(__file__), "golden")
    z = np.load(os.path.join(d, "radon_golden.npz"))
    img, th = z["img"], z["theta"]
    for interp in INTERPS:
        got = cp.project_tf_fast(torch.from_numpy(img).cuda().unsqueeze(-1), th, pad=True, dim=2, integrate_vae=True,
                                 interpolation=interp)[..., 0].cpu().numpy()
        assert rel_l2(got, z[f"sino_{interp}"]) <= TOL
        for mode in ("exact", "tf_compat"):
            g = cp.backproject(torch.from_numpy(z["cot"]).cuda(), th, img.shape[1], img.shape[2], pad=True,
                               interpolation=interp, adjoint=mode).cpu().numpy()
            assert rel_l2(g, z[f"grad_{mode}_{interp}"]) <= TOL
    rec = cp.iradon(torch.from_numpy(z["fbp_sino"]).cuda(), th, img.shape[1], img.shape[2], z["fbp_filter"])
    assert rel_l2(rec.cpu().numpy(), z["fbp_recon"]) <= TOL


def test_errors_are_loud(cp):
    with pytest.raises(RuntimeError):
        from ct_pvae_b200 import ops, _lib

        plan = _lib.get_plan(np.array([0.0]), 4, 4, False, 0)
        ops.radon_forward(torch.zeros((1, 4, 4)), plan, 0)   # CPU tensor: no fallback
    with pytest.raises(ValueError):
        cp.project_tf_fast(torch.zeros((4, 4), device="cuda"), np.zeros((2, 2)), dim=2)  # rank-2 angles
    with pytest.raises(TypeError):
        cp.project_tf_fast(torch.zeros((4, 4), dtype=torch.int32, device="cuda"), np.zeros(2), dim=2)
