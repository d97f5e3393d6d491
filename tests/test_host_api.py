"""Host-side mirror of the reference interface (no GPU): signatures, pad_phantom,
argument validation, and that nothing silently falls back to the CPU."""
import inspect

import numpy as np
import pytest
import torch

import ct_pvae_b200 as cp
from ct_pvae_b200 import fbp_tensorflow, forward_functions, sharding


def test_reference_signatures_are_preserved():
    # ctvae/forward_functions.py:18,49,80 and ctvae/fbp_tensorflow.py:14
    def pos(fn):
        return [(p.name, p.default) for p in inspect.signature(fn).parameters.values() if p.kind is p.POSITIONAL_OR_KEYWORD]

    E = inspect.Parameter.empty
    assert pos(forward_functions.pad_phantom) == [("phantom", E), ("dim", 3), ("integrate_vae", False)]
    assert pos(forward_functions.project_tf_low_mem) == [("phantom", E), ("theta", E), ("pad", False)]
    assert pos(forward_functions.project_tf_fast) == [("phantom", E), ("theta", E), ("pad", False), ("dim", 3), ("integrate_vae", False)]
    assert pos(fbp_tensorflow.iradon) == [("sinogram", E), ("theta", E), ("x_size", E), ("y_size", E), ("filter_1d", E)]
    kw = inspect.signature(forward_functions.project_tf_fast).parameters
    assert kw["interpolation"].default == "nearest" and kw["adjoint"].default == "exact"     # tfa default, north_star default
    assert inspect.signature(forward_functions.project_tf_low_mem).parameters["interpolation"].default == "bilinear"


def test_pad_phantom_layouts(orc):
    for X, Y in [(128, 128), (5, 8), (2, 2)]:
        P = orc.num_proj_pix(X, Y)
        a = np.arange(X * Y, dtype=np.float32).reshape(X, Y)
        p2 = cp.pad_phantom(a, dim=2)
        assert isinstance(p2, np.ndarray) and p2.shape == (P, P)
        px, py = (P - X) // 2, (P - Y) // 2
        np.testing.assert_array_equal(p2[px:px + X, py:py + Y], a)
        assert p2.sum() == a.sum()
        assert cp.pad_phantom(torch.zeros(X, Y, 3)).shape == (P, P, 3)
        assert cp.pad_phantom(torch.zeros(4, X, Y, 1), integrate_vae=True).shape == (4, P, P, 1)


def test_argument_validation_happens_before_the_device_is_needed():
    x = torch.zeros(4, 4)
    with pytest.raises(ValueError):
        cp.project_tf_fast(x, np.zeros(2), dim=3)                       # rank mismatch
    with pytest.raises(ValueError):
        cp.project_tf_fast(torch.zeros(2, 4, 4, 2), np.zeros(2), integrate_vae=True)   # channels != 1
    with pytest.raises(TypeError):
        cp.project_tf_fast(x, 0.5, dim=2)                               # len(theta) like the reference
    with pytest.raises(TypeError):
        cp.project_tf_fast(torch.zeros(4, 4, dtype=torch.int64), np.zeros(2), dim=2)
    with pytest.raises(ValueError):
        cp.iradon(np.zeros((1, 5, 16)), np.zeros(4), 8, 8, np.ones(16))  # fbp_tensorflow.py:43-45


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cp.project_tf_fast(np.ones((4, 4), np.float32), np.array([0.0]), dim=2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cp.iradon(np.zeros((1, 4, 16)), np.zeros(4), 8, 8, np.ones(16))


def test_product_code_never_touches_the_oracle():
    import os
    import re

    root = os.path.dirname(os.path.abspath(cp.__file__))
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|radon_oracle", src, flags=re.M), f


def test_fourier_filter_matches_oracle_restatement(orc):
    for name in ("ramp", "shepp-logan", "cosine", "hamming", "hann", None):
        np.testing.assert_allclose(cp.get_fourier_filter(184, name), orc.get_fourier_filter(184, name))


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 64, 720):
        for world in (1, 2, 3, 8):
            blocks = [sharding.shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(b[1] == c[0] for b, c in zip(blocks, blocks[1:]))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_tf_bridge_is_import_guarded():
    from ct_pvae_b200 import tf_bridge

    if tf_bridge.available():
        pytest.skip("TensorFlow present")
    with pytest.raises(RuntimeError, match="TensorFlow is not installed"):
        tf_bridge.project_tf_fast(np.zeros((2, 4, 4, 1), np.float32), np.zeros(3), integrate_vae=True)


def test_host_pipeline_eligibility():
    from ct_pvae_b200 import hostpipe

    x = torch.zeros(40, 8, 8)
    assert hostpipe.eligible(x)                           # pageable memory: staged through the pipe's pinned ring
    assert hostpipe.eligible(torch.from_numpy(np.zeros((40, 8, 8), np.float32)))
    assert not hostpipe.eligible(x.permute(0, 2, 1))      # the C call wants one contiguous block
    assert not hostpipe.eligible(torch.zeros(40, 8, 8, dtype=torch.float64))
    assert not hostpipe.eligible(torch.zeros(4, 8, 8))    # small batches are not worth chunking
    assert not hostpipe.eligible(torch.zeros(40, 8, 8, requires_grad=True))


def test_host_chunk_sizes(monkeypatch):
    """Chunks of the native host pipeline: four per call, whole 16-image records, enough work per launch."""
    import types

    from ct_pvae_b200 import hostpipe
    hostpipe.set_chunk(0, 0)
    c2 = types.SimpleNamespace(A=180, X=128, Y=128)
    c4 = types.SimpleNamespace(A=720, X=512, Y=512)
    assert hostpipe.chunk_for(c2, 256, "fwd") == 64
    assert hostpipe.chunk_for(c2, 1024, "adj") == 256
    assert hostpipe.chunk_for(c4, 64, "fwd") == 32         # two chunks of whole 32-image records
    assert hostpipe.chunk_for(c4, 32, "fwd") == 16
    assert hostpipe.chunk_for(c2, 40, "fwd") == 48          # one chunk: the batch is too small to split
    hostpipe.set_chunk(32, 0)
    try:
        assert hostpipe.chunk_for(c2, 256, "fwd") == 32 and hostpipe.chunk_for(c2, 256, "adj") == 64
    finally:
        hostpipe.set_chunk(0, 0)
