"""Pins the CPU oracle (oracle/) with everything the reference offers for this path:
the toy dataset's closed-form sinograms, theta=0 column sums, the pad_phantom size
table, plus independent implementations (numpy restatement, explicit sparse matrix,
torch grid_sample) and the committed golden vectors.  No GPU."""
import os

import numpy as np
import pytest
import torch

from conftest import rel_l2

MODES = [0, 1]  # nearest, bilinear


def test_num_proj_pix_table(orc):
    # P = 2*ceil((sqrt(X^2+Y^2)+2)/2); main_ct_vae.py:160-161 inverts it
    for X, P in [(2, 6), (128, 184), (256, 366), (512, 728)]:
        assert orc.num_proj_pix(X, X) == P == orc.lib().orc_num_proj_pix(X, X)
    for P, X in [(184, 128), (728, 512)]:
        assert int(np.floor(P / np.sqrt(2) - 2)) == X


@pytest.mark.parametrize("mode", MODES)
def test_toy_known_answers(orc, mode):
    # scripts/create_toy_images.py:36-37, scripts/images_to_sinograms.py:54-59:
    # theta=0 -> column sums; theta=pi/2 -> row sums reversed along the detector
    x0 = np.array([[1, 2], [3, 4]], np.float32) / 10
    x1 = np.array([[3, 4], [1, 2]], np.float32) / 10
    s = orc.forward(np.stack([x0, x1]), [0, np.pi / 2], False, mode)
    np.testing.assert_allclose(s[0], [[0.4, 0.6], [0.7, 0.3]], rtol=1e-6)
    np.testing.assert_allclose(s[1], [[0.4, 0.6], [0.3, 0.7]], rtol=1e-6)
    for img in (x0, x1):
        np.testing.assert_allclose(orc.forward(img[None], [0.0], False, mode)[0, 0], img.sum(axis=0), rtol=1e-6)
        np.testing.assert_allclose(orc.forward(img[None], [np.pi / 2], False, mode)[0, 0], img.sum(axis=1)[::-1], rtol=1e-6)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("pad", [False, True])
def test_theta_zero_is_column_sums(orc, mode, pad):
    rng = np.random.default_rng(0)
    img = rng.random((2, 21, 34), dtype=np.float32)
    s = orc.forward(img, [0.0], pad, mode)[:, 0]
    H, W, padx, pady = orc.frame_of(21, 34, pad)
    want = np.zeros((2, W))
    want[:, pady:pady + 34] = img.astype(np.float64).sum(axis=1)
    np.testing.assert_allclose(s, want, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("mode", MODES)
def test_c_and_numpy_restatements_agree(orc, mode):
    rng = np.random.default_rng(1)
    img = rng.random((2, 24, 30), dtype=np.float32)
    th = np.linspace(0, np.pi, 11, endpoint=False)
    for pad in (False, True):
        a = orc.forward(img, th, pad, mode)
        # numpy's float32 cos/sin can differ from libm's by an ulp -> tolerance, not equality
        assert rel_l2(orc.forward_np(img, th, pad, mode), a) <= (3e-3 if mode == 0 else 2e-6)
        assert rel_l2(orc.forward(img, th, pad, mode, dataflow=True), a) <= 1e-6
    H, W, _, _ = orc.frame_of(24, 30, True)
    assert np.abs(orc.make_transforms(th, H, W) - orc.make_transforms_np(th, H, W)).max() <= 1e-5


@pytest.mark.parametrize("mode", MODES)
def test_forward_is_the_sparse_matrix_and_adjoint_its_transpose(orc, mode):
    rng = np.random.default_rng(2)
    X, Y, A = 14, 19, 9
    th = rng.uniform(0, np.pi, A)
    img = rng.random((2, X, Y), dtype=np.float32)
    M = orc.build_matrix(th, X, Y, True, mode)
    # build_matrix uses the numpy table; rebuild it on the C table for an exact comparison
    H, W, padx, pady = orc.frame_of(X, Y, True)
    s = orc.forward(img, th, True, mode)
    s_m = (M @ img.reshape(2, -1).T.astype(np.float64)).T.reshape(s.shape)
    assert rel_l2(s, s_m) <= (5e-3 if mode == 0 else 2e-6)
    y = rng.random(s.shape, dtype=np.float32)
    g = orc.adjoint_exact(y, th, X, Y, True, mode)
    g_m = (M.T @ y.reshape(2, -1).T.astype(np.float64)).T.reshape(g.shape)
    assert rel_l2(g, g_m) <= (5e-3 if mode == 0 else 2e-6)
    # adjoint identity on the C oracle itself (same table both sides): float64-exact
    lhs = float((s.astype(np.float64) * y).sum())
    rhs = float((img.astype(np.float64) * g).sum())
    assert abs(lhs - rhs) / abs(lhs) <= 1e-6


def test_bilinear_matches_torch_grid_sample(orc):
    """Independent bilinear implementation: grid_sample(align_corners=True, zeros) on the
    same float32 coordinates reproduces the rotated stack, hence the sinogram."""
    rng = np.random.default_rng(3)
    X = Y = 20
    img = rng.random((1, X, Y), dtype=np.float32)
    th = np.linspace(0, np.pi, 7, endpoint=False)
    H, W, padx, pady = orc.frame_of(X, Y, True)
    t = orc.make_transforms(th, H, W)
    padded = np.zeros((H, W), np.float32)
    padded[padx:padx + X, pady:pady + Y] = img[0]
    ox, oy = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    out = []
    for a in range(len(th)):
        x = (t[a, 0] * ox + t[a, 1] * oy) + t[a, 2]
        y = (t[a, 3] * ox + t[a, 4] * oy) + t[a, 5]
        grid = np.stack([2 * x / (W - 1) - 1, 2 * y / (H - 1) - 1], axis=-1)[None]
        rot = torch.nn.functional.grid_sample(torch.from_numpy(padded)[None, None].double(),
                                              torch.from_numpy(grid).double(), mode="bilinear",
                                              padding_mode="zeros", align_corners=True)[0, 0]
        out.append(rot.sum(dim=0).numpy())
    assert rel_l2(orc.forward(img, th, True, 1)[0], np.stack(out)) <= 2e-6


@pytest.mark.parametrize("mode", MODES)
def test_tf_gradient_is_not_the_transpose_but_close(orc, mode):
    # SURVEY finding 3: TF's registered gradient differs from A^T by percents
    rng = np.random.default_rng(4)
    th = np.linspace(0, np.pi, 20, endpoint=False)
    W = orc.frame_of(32, 32, True)[1]
    y = rng.random((1, 20, W), dtype=np.float32)
    d = rel_l2(orc.adjoint_tf(y, th, 32, 32, True, mode), orc.adjoint_exact(y, th, 32, 32, True, mode))
    assert 1e-3 < d < 0.5


def test_inverse_transforms_are_rotations_by_plus_theta(orc):
    th = np.linspace(0.1, 3.0, 9)
    t = orc.make_transforms(th, 46, 46)
    np.testing.assert_allclose(orc.invert_transforms(t), orc.make_transforms(-th, 46, 46), atol=2e-5)


def test_iradon_restatements_and_errors(orc):
    rng = np.random.default_rng(5)
    sino = rng.random((2, 12, 40))
    th = np.linspace(0, np.pi, 12, endpoint=False)
    for name in ("ramp", "hann", None):
        f = orc.get_fourier_filter(40, name)
        a, b = orc.iradon(sino, th, 27, 22, f), orc.iradon_np(sino, th, 27, 22, f)
        assert a.dtype == np.float64 and rel_l2(a, b) <= 1e-12
    with pytest.raises(ValueError):
        orc.iradon(sino, th[:5], 27, 22, np.ones(40))
    # real(ifft(fft(s) * F)) == circular convolution with real(ifft(F)) for real s
    f = orc.get_fourier_filter(40, "ramp")
    h = np.real(np.fft.ifft(f))
    conv = np.stack([[sum(sino[b, a, k] * h[(n - k) % 40] for k in range(40)) for n in range(40)]
                     for b in range(1) for a in range(2)]).reshape(1, 2, 40)
    np.testing.assert_allclose(conv, np.real(np.fft.ifft(np.fft.fft(sino[:1, :2], axis=-1) * f, axis=-1)), atol=1e-12)


def test_iradon_reconstructs_a_disk(orc):
    # X=32 -> P=48: skimage's filter formula (restated verbatim) is only well formed when
    # P % 4 == 0, which holds for the reference's sizes (128 -> 184, 512 -> 728)
    X = 32
    yy, xx = np.meshgrid(np.arange(X) - X / 2 + 0.5, np.arange(X) - X / 2 + 0.5, indexing="ij")
    img = (0.8 * ((xx ** 2 + yy ** 2) <= 10 ** 2)).astype(np.float32)
    th = np.linspace(0, np.pi, 90, endpoint=False)
    s = orc.forward(img[None], th, True, 1)
    rec = orc.iradon(s, th, X, X, orc.get_fourier_filter(s.shape[2], "ramp"))[0]
    assert abs(rec[(xx ** 2 + yy ** 2) <= 7 ** 2].mean() - 0.8) < 0.01


def test_golden_vectors(orc):
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "radon_golden.npz"))
    img, th, cot = z["img"], z["theta"], z["cot"]
    X, Y = img.shape[1:]
    for name, iid in (("nearest", 0), ("bilinear", 1)):
        np.testing.assert_array_equal(orc.forward(img, th, True, iid), z[f"sino_{name}"])
        np.testing.assert_array_equal(orc.adjoint_exact(cot, th, X, Y, True, iid), z[f"grad_exact_{name}"])
        np.testing.assert_array_equal(orc.adjoint_tf(cot, th, X, Y, True, iid), z[f"grad_tf_compat_{name}"])
    np.testing.assert_allclose(orc.iradon(z["fbp_sino"], th, X, Y, z["fbp_filter"]), z["fbp_recon"], rtol=1e-12, atol=1e-14)


def test_edge_cases(orc):
    # empty batch, single pixel, single angle, angles outside [0, pi)
    assert orc.forward(np.zeros((0, 4, 4), np.float32), [0.3], True, 1).shape == (0, 1, orc.num_proj_pix(4, 4))
    one = orc.forward(np.ones((1, 1, 1), np.float32), [0.0, 1.0, -2.0, 7.0], True, 1)
    # a rotated unit lattice is not an exact partition of unity for the bilinear hat: mass within ~10 %
    assert one.shape == (1, 4, orc.num_proj_pix(1, 1)) and np.all(np.abs(one.sum(axis=2) - 1) < 0.1) and one[0, 0, 1] == 1
    img = np.random.default_rng(6).random((1, 9, 9), dtype=np.float32)
    np.testing.assert_allclose(orc.forward(img, [0.4], True, 1), orc.forward(img, [0.4 + 2 * np.pi], True, 1), atol=2e-5)


@pytest.mark.parametrize("mode,order,tol", [(1, 1, 2e-6), (0, 0, 2e-3)])
def test_forward_matches_scipy_affine_transform(orc, mode, order, tol):
    """Independent public implementation of the same operator: scipy.ndimage.affine_transform with the
    tfa transform as (matrix, offset), zero fill, order 1 (bilinear) / 0 (nearest), then the row sum of
    forward_functions.py:114.  float64 coordinates there, float32 here: bilinear agrees to rounding; nearest
    may flip a sample that sits within ~1e-6 of a rounding boundary (none does for this seed: 3e-8)."""
    import scipy.ndimage as ndi

    rng = np.random.default_rng(3)
    X, Y, A = 40, 33, 9
    img = rng.random((2, X, Y), dtype=np.float32)
    th = np.linspace(0, np.pi, A, endpoint=False) + 0.1
    H, W, padx, pady = orc.frame_of(X, Y, True)
    t = orc.make_transforms(th, H, W).astype(np.float64)
    want = orc.forward(img, th, True, mode)
    got = np.zeros(want.shape)
    for b in range(img.shape[0]):
        padded = np.zeros((H, W))
        padded[padx:padx + X, pady:pady + Y] = img[b]
        for a in range(A):
            c, ms, xo, s, c2, yo = t[a, :6]
            # output (row, col) -> input (row, col): y_in = s*x + c*y + y_off ; x_in = c*x - s*y + x_off
            rot = ndi.affine_transform(padded, np.array([[c2, s], [ms, c]]), offset=np.array([yo, xo]), order=order,
                                       mode="constant", cval=0.0, output_shape=(H, W))
            got[b, a] = rot.sum(axis=0)
    assert np.linalg.norm(got - want) / np.linalg.norm(want) <= tol


@pytest.mark.parametrize("mode,order,tol", [(1, 1, 2e-6), (0, 0, 2e-3)])
def test_tf_gradient_matches_scipy_affine_transform(orc, mode, order, tol):
    """TensorFlow's registered gradient of the projector graph, restated independently: the cotangent of the
    row sum is the sinogram row broadcast over the rows of the frame; ImageProjectiveTransformV3's gradient
    resamples it with the inverted transforms (same interpolation, zero fill); pad's gradient is the crop."""
    import scipy.ndimage as ndi

    rng = np.random.default_rng(4)
    X, Y, A = 40, 33, 9
    th = np.linspace(0, np.pi, A, endpoint=False) + 0.1
    H, W, padx, pady = orc.frame_of(X, Y, True)
    y = rng.random((2, A, W), dtype=np.float32)
    ti = orc.invert_transforms(orc.make_transforms(th, H, W)).astype(np.float64)
    want = orc.adjoint_tf(y, th, X, Y, True, mode)
    got = np.zeros((2, H, W))
    for b in range(2):
        for a in range(A):
            g = np.broadcast_to(y[b, a][None, :].astype(np.float64), (H, W))
            c, ms, xo, s, c2, yo = ti[a, :6]
            got[b] += ndi.affine_transform(g, np.array([[c2, s], [ms, c]]), offset=np.array([yo, xo]), order=order,
                                           mode="constant", cval=0.0, output_shape=(H, W))
    got = got[:, padx:padx + X, pady:pady + Y]
    assert np.linalg.norm(got - want) / np.linalg.norm(want) <= tol
