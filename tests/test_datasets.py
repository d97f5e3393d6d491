"""On-disk formats and metrics of SURVEY 8f-4 (host-side glue, no GPU)."""
import numpy as np

from ct_pvae_b200 import datasets


def test_dataset_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    s = rng.random((5, 9, 16)) - 0.1
    theta = np.linspace(0, np.pi, 9, endpoint=False)
    datasets.save_dataset(str(tmp_path), s, theta, 8, 8)
    got, th, P = datasets.get_sinograms(str(tmp_path))
    assert P == 16 and np.array_equal(th, theta) and got.min() >= 0 and np.array_equal(got, np.where(s < 0, 0, s))
    # the reference reads the pickled pair the same way (helper_functions.py:50-52)
    t2, p2 = np.load(tmp_path / "dataset_parameters.npy", allow_pickle=True)
    assert int(p2) == 16 and len(t2) == 9 and int(np.load(tmp_path / "x_size.npy")) == 8


def test_reconstruction_size_inverts_pad_phantom():
    assert datasets.reconstruction_size(184) == (128, 128) and datasets.reconstruction_size(728) == (512, 512)
    assert datasets.reconstruction_size(2, no_pad=True) == (2, 2)


def test_crop_is_centred():
    a = np.arange(100).reshape(10, 10)
    assert datasets.crop(a, 4, 5).shape == (4, 5) and datasets.crop(a, 4, 5)[0, 0] == a[3, 3]
    assert datasets.crop(np.stack([a, a]), 6, 6, ignore_dim_0=True).shape == (2, 6, 6)


def test_compare_metrics():
    rng = np.random.default_rng(1)
    a = rng.random((32, 32))
    mse, ssim, psnr = datasets.compare(a, a)
    assert mse == 0 and abs(ssim - 1) < 1e-12 and psnr == float("inf")
    b = a + 0.05 * rng.standard_normal(a.shape)
    mse, ssim, psnr = datasets.compare(a, b)
    dr = a.max() - a.min()
    assert abs(mse - np.mean((a - b) ** 2)) < 1e-15 and abs(psnr - 10 * np.log10(dr ** 2 / mse)) < 1e-9 and 0 < ssim < 1
    assert 0 < datasets.compare(a[:5, :5], b[:5, :5])[1] <= 1        # small images: odd window <= side
