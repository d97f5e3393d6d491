"""Known-answer vectors of the third-party rotation the reference calls (`tfa.image.rotate`, forward_functions.py:113).

tensorflow-addons is an un-vendored dependency of the reference (requirements: tensorflow-addons 0.17.x on TF 2.8), so
its arithmetic is not under /root/reference.  What IS published is the library's own test suite: the arrays below are
the expected outputs asserted by
  * tensorflow_addons/image/tests/transform_ops_test.py::test_rotate_even / ::test_rotate_odd  (nearest, exact integers)
  * tf.contrib.image kernel_tests/image_ops_test.py::test_bilinear, kept by tfa as ::test_bilinear  (bilinear, atol 1e-3,
    "matches scipy.ndimage.rotate(image, 45, order=1, reshape=False)")
transcribed here.  They pin the conventions no restatement can derive from the reference alone: the centre of rotation,
the direction, the half-away rounding of nearest, zero fill outside the frame, the bilinear tap order.  Checked against
  1. oracle/tf_shim's `tfa.image.rotate` (the stand-in the reference's own files run under for tests/golden),
  2. the oracle's projector (column sums of the rotated image; the reference rotates by -theta, so theta = -angle),
  3. on the GPU, the CUDA projector through the drop-in API.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

EVEN_ANGLES = np.array([0.0, np.pi / 4.0, np.pi / 2.0], np.float32)
EVEN = np.array([
    [[0, 1, 2, 3, 4, 5], [6, 7, 8, 9, 10, 11], [12, 13, 14, 15, 16, 17], [18, 19, 20, 21, 22, 23], [24, 25, 26, 27, 28, 29],
     [30, 31, 32, 33, 34, 35]],
    [[0, 3, 4, 11, 17, 0], [2, 3, 9, 16, 23, 23], [1, 8, 15, 21, 22, 29], [6, 13, 20, 21, 27, 34], [12, 18, 19, 26, 33, 33],
     [0, 18, 24, 31, 32, 0]],
    [[5, 11, 17, 23, 29, 35], [4, 10, 16, 22, 28, 34], [3, 9, 15, 21, 27, 33], [2, 8, 14, 20, 26, 32], [1, 7, 13, 19, 25, 31],
     [0, 6, 12, 18, 24, 30]]], np.float32)
ODD_ANGLES = np.array([np.pi / 4.0, 1.0, -np.pi / 2.0], np.float32)
ODD = np.array([
    [[0, 3, 8, 9, 0], [1, 7, 8, 13, 19], [6, 6, 12, 18, 18], [5, 11, 16, 17, 23], [0, 15, 16, 21, 0]],
    [[0, 3, 9, 14, 0], [2, 7, 8, 13, 19], [1, 6, 12, 18, 23], [5, 11, 16, 17, 22], [0, 10, 15, 21, 0]],
    [[20, 15, 10, 5, 0], [21, 16, 11, 6, 1], [22, 17, 12, 7, 2], [23, 18, 13, 8, 3], [24, 19, 14, 9, 4]]], np.float32)
RING = np.array([[0, 0, 0, 0, 0], [0, 1, 1, 1, 0], [0, 1, 0, 1, 0], [0, 1, 1, 1, 0], [0, 0, 0, 0, 0]], np.float32)
RING_BILINEAR = np.array([[0.000, 0.000, 0.343, 0.000, 0.000], [0.000, 0.586, 0.914, 0.586, 0.000],
                          [0.343, 0.914, 0.000, 0.914, 0.343], [0.000, 0.586, 0.914, 0.586, 0.000],
                          [0.000, 0.000, 0.343, 0.000, 0.000]], np.float32)
RING_NEAREST = np.array([[0, 0, 1, 0, 0], [0, 1, 1, 1, 0], [1, 1, 0, 1, 1], [0, 1, 1, 1, 0], [0, 0, 1, 0, 0]], np.float32)

# (image, angles, expected rotated images, interpolation, atol of the upstream assertion)
CASES = [
    ("rotate_even", np.arange(36, dtype=np.float32).reshape(6, 6), EVEN_ANGLES, EVEN, "nearest", 0.0),
    ("rotate_odd", np.arange(25, dtype=np.float32).reshape(5, 5), ODD_ANGLES, ODD, "nearest", 0.0),
    ("ring_nearest", RING, np.array([np.pi / 4.0], np.float32), RING_NEAREST[None], "nearest", 0.0),
    ("ring_bilinear", RING, np.array([np.pi / 4.0], np.float32), RING_BILINEAR[None], "bilinear", 1e-3),
]
IID = {"nearest": 0, "bilinear": 1}


@pytest.fixture()
def shim_rotate():
    """`tfa.image.rotate` of oracle/tf_shim (import confined to this fixture: the shim shadows real packages by name)."""
    shim = os.path.join(ROOT, "oracle", "tf_shim")
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k.split(".")[0] in ("tensorflow", "tensorflow_addons", "tensorflow_probability")}
    sys.path.insert(0, shim)
    try:
        import tensorflow_addons as tfa
        yield tfa.image.rotate
    finally:
        sys.path.remove(shim)
        for k in list(sys.modules):
            if k.split(".")[0] in ("tensorflow", "tensorflow_addons", "tensorflow_probability"):
                del sys.modules[k]
        sys.modules.update(saved)


@pytest.mark.parametrize("name,image,angles,want,interp,atol", CASES, ids=[c[0] for c in CASES])
def test_shim_rotation_reproduces_the_librarys_known_answers(shim_rotate, name, image, angles, want, interp, atol):
    rep = np.tile(image[None, :, :, None], [len(angles), 1, 1, 1])
    got = np.asarray(shim_rotate(rep, angles, interpolation=interp))[..., 0]
    if atol == 0.0:
        np.testing.assert_array_equal(got, want)
    else:
        np.testing.assert_allclose(got, want, atol=atol)


@pytest.mark.parametrize("name,image,angles,want,interp,atol", CASES, ids=[c[0] for c in CASES])
def test_oracle_projector_reproduces_the_known_answers(orc, name, image, angles, want, interp, atol):
    # reference: tf.reduce_sum(tfa.image.rotate(imgs, -theta), 1) -> column sums of the image rotated by -theta
    theta = -angles.astype(np.float64)
    got = orc.forward(image[None], theta, False, IID[interp])[0]
    cols = want.sum(axis=1)
    if atol == 0.0:
        np.testing.assert_array_equal(got, cols)                      # sums of small integers: exact in float32
    else:
        np.testing.assert_allclose(got, cols, atol=atol * image.shape[0])
    np.testing.assert_allclose(orc.forward_np(image[None], theta, False, IID[interp])[0], cols, atol=max(atol * image.shape[0], 1e-6))


@pytest.mark.gpu
@pytest.mark.parametrize("name,image,angles,want,interp,atol", CASES, ids=[c[0] for c in CASES])
def test_cuda_projector_reproduces_the_known_answers(name, image, angles, want, interp, atol):
    import torch

    import ct_pvae_b200 as cp
    assert torch.cuda.is_available()
    theta = -angles.astype(np.float64)
    x = torch.from_numpy(image[None, :, :, None]).cuda()
    got = cp.project_tf_fast(x, theta, pad=False, dim=2, integrate_vae=True, interpolation=interp)[0, ..., 0].cpu().numpy()
    cols = want.sum(axis=1)
    if atol == 0.0:
        np.testing.assert_array_equal(got, cols)
    else:
        np.testing.assert_allclose(got, cols, atol=atol * image.shape[0])
