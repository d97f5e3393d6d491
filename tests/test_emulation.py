"""The kernels' per-thread code (ctr_core.h), replayed on the CPU with the CTA structure
of ctr_kernels.cuh (packs, strips, angle classes, pixel tiles, bin windows), against the
oracle.  Catches geometry bugs without a GPU; the GPU glue is covered by -m gpu tests."""
import ctypes

import numpy as np
import pytest

from conftest import rel_l2

f32p = ctypes.POINTER(ctypes.c_float)
f64p = ctypes.POINTER(ctypes.c_double)


def P(a):
    return a.ctypes.data_as(f32p)


CASES = [
    # B, X, Y, A, pad, R (strip rows), TW, TH, win
    (1, 2, 2, 2, False, 4, 32, 8, 44),
    (5, 16, 16, 12, True, 4, 32, 8, 44),
    (3, 16, 16, 12, False, 8, 32, 8, 44),
    (2, 33, 20, 17, True, 5, 32, 16, 44),
    (2, 20, 33, 17, False, 3, 32, 16, 44),
    (4, 64, 64, 24, True, 8, 32, 16, 44),
    (1, 96, 96, 16, True, 32, 32, 16, 44),
    (41, 12, 10, 7, True, 4, 32, 8, 44),                # two 32-image records: every register slot of the swizzled / rotated lanes
]


@pytest.mark.parametrize("B,X,Y,A,pad,R,TW,TH,win", CASES)
def test_emulated_kernels_match_oracle(emu, orc, B, X, Y, A, pad, R, TW, TH, win):
    rng = np.random.default_rng(0)
    img = rng.random((B, X, Y), dtype=np.float32)
    th = np.array([0, np.pi / 2]) if (X, Y) == (2, 2) else np.linspace(0, np.pi, A, endpoint=False)
    A = len(th)
    H, W, padx, pady = orc.frame_of(X, Y, pad)
    t = np.empty((A, 8), np.float32)
    emu.emu_make_transforms(th.ctypes.data_as(f64p), A, H, W, P(t))
    np.testing.assert_array_equal(t, orc.make_transforms(th, H, W))          # product table == oracle table, bit for bit
    ti = np.empty_like(t)
    emu.emu_invert_transforms(P(t), A, P(ti))
    np.testing.assert_array_equal(ti, orc.invert_transforms(t))
    y = rng.random((B, A, W), dtype=np.float32)
    for interp in (0, 1):
        s = np.full((B, A, W), np.nan, np.float32)
        emu.emu_forward(P(img), B, X, Y, H, W, padx, pady, P(t), A, interp, R, P(s))
        assert rel_l2(s, orc.forward(img, th, pad, interp)) <= 1e-6
        sd = np.full((B, A, W), np.nan, np.float32)       # depth-first records (16 images per record)
        emu.emu_forward_depth(P(img), B, X, Y, H, W, padx, pady, P(t), A, interp, R, P(sd))
        s32 = np.full_like(sd, np.nan)
        emu.emu_forward_rec32(P(img), B, X, Y, H, W, padx, pady, P(t), A, interp, R, P(s32))
        assert np.array_equal(s32, sd), "32-image records (swizzled 8-image lanes, reuse march) differ from 16-image records"
        s32p = np.full_like(sd, np.nan)
        emu.emu_forward_rec32_plain(P(img), B, X, Y, H, W, padx, pady, P(t), A, interp, R, P(s32p))
        assert np.array_equal(s32p, sd), "32-image records (plain march) differ from 16-image records"
        np.testing.assert_array_equal(sd, s)
        for reuse in (0, 1):
            sw = np.full_like(sd, np.nan)
            emu.emu_forward_wide(P(img), B, X, Y, H, W, padx, pady, P(t), A, interp, R, reuse, P(sw))
            assert np.array_equal(sw, sd), "32-image records (two lanes per ray, rotated 16-image loads) differ from 16-image records"
        g = np.full((B, X, Y), np.nan, np.float32)
        emu.emu_adjoint(P(y), B, X, Y, H, W, padx, pady, P(t), A, interp, 0, TW, TH, win, P(g))
        assert rel_l2(g, orc.adjoint_exact(y, th, X, Y, pad, interp)) <= 1e-6
        g2 = np.full((B, X, Y), np.nan, np.float32)
        emu.emu_adjoint(P(y), B, X, Y, H, W, padx, pady, P(ti), A, interp, 1, TW, TH, win, P(g2))
        assert rel_l2(g2, orc.adjoint_tf(y, th, X, Y, pad, interp)) <= 1e-6


def test_emulated_kernels_random_angles(emu, orc):
    rng = np.random.default_rng(1)
    th = rng.uniform(-7, 7, 9)
    img = rng.random((2, 40, 40), dtype=np.float32)
    H, W, padx, pady = orc.frame_of(40, 40, True)
    t = orc.make_transforms(th, H, W)
    for interp in (0, 1):
        s = np.full((2, 9, W), np.nan, np.float32)
        emu.emu_forward(P(img), 2, 40, 40, H, W, padx, pady, P(t), 9, interp, 7, P(s))
        assert rel_l2(s, orc.forward(img, th, True, interp)) <= 1e-6


# ---- property-based sweep: random shapes, angles (any sign / magnitude), pad, strip and tile sizes ----
from hypothesis import given, settings, strategies as st  # noqa: E402


@settings(max_examples=30, deadline=None)
@given(B=st.integers(1, 6), X=st.integers(1, 28), Y=st.integers(1, 28), pad=st.booleans(), R=st.integers(1, 9),
       th=st.lists(st.floats(-7.0, 7.0, allow_nan=False, width=32), min_size=1, max_size=6), seed=st.integers(0, 2 ** 16))
def test_emulated_kernels_property(emu, orc, B, X, Y, pad, R, th, seed):
    rng = np.random.default_rng(seed)
    img = rng.random((B, X, Y), dtype=np.float32)
    th = np.asarray(th, np.float64)
    A = len(th)
    H, W, padx, pady = orc.frame_of(X, Y, pad)
    t = orc.make_transforms(th, H, W)
    ti = orc.invert_transforms(t)
    y = rng.random((B, A, W), dtype=np.float32)
    for interp in (0, 1):
        want = orc.forward(img, th, pad, interp)
        s = np.full((B, A, W), np.nan, np.float32)
        emu.emu_forward(P(img), B, X, Y, H, W, padx, pady, P(t), A, interp, R, P(s))
        assert np.abs(s - want).max() <= 2e-5 * max(1.0, np.abs(want).max())
        sd = np.full((B, A, W), np.nan, np.float32)
        emu.emu_forward_depth(P(img), B, X, Y, H, W, padx, pady, P(t), A, interp, R, P(sd))
        s32 = np.full_like(sd, np.nan)
        emu.emu_forward_rec32(P(img), B, X, Y, H, W, padx, pady, P(t), A, interp, R, P(s32))
        assert np.array_equal(s32, sd), "32-image records (swizzled 8-image lanes, reuse march) differ from 16-image records"
        s32p = np.full_like(sd, np.nan)
        emu.emu_forward_rec32_plain(P(img), B, X, Y, H, W, padx, pady, P(t), A, interp, R, P(s32p))
        assert np.array_equal(s32p, sd), "32-image records (plain march) differ from 16-image records"
        np.testing.assert_array_equal(sd, s)
        sw = np.full_like(sd, np.nan)
        emu.emu_forward_wide(P(img), B, X, Y, H, W, padx, pady, P(t), A, interp, R, 1, P(sw))
        assert np.array_equal(sw, sd), "32-image records (two lanes per ray, rotated 16-image loads) differ from 16-image records"
        for mode, table, fn in ((0, t, orc.adjoint_exact), (1, ti, orc.adjoint_tf)):
            g = np.full((B, X, Y), np.nan, np.float32)
            emu.emu_adjoint(P(y), B, X, Y, H, W, padx, pady, P(table), A, interp, mode, 32, 8, 40, P(g))
            gw = fn(y, th, X, Y, pad, interp)
            assert np.abs(g - gw).max() <= 2e-5 * max(1.0, np.abs(gw).max())


# ---- column-windowed strips (wide detectors, 16-image records) ----
WIN_CASES = [
    # B, X, Y, A, pad, JW, NA, Rmax, budget bytes (2 strip buffers), theta kind
    (3, 64, 64, 24, True, 24, 2, 8, 60000, "even"),
    (2, 64, 64, 24, True, 16, 4, 6, 40000, "even"),
    (2, 48, 80, 20, True, 32, 2, 8, 80000, "even"),
    (2, 80, 48, 20, False, 24, 2, 5, 50000, "even"),
    (2, 64, 64, 11, True, 24, 2, 8, 80000, "random"),       # far-apart angles in one CTA: wide or whole-row windows
    (17, 40, 40, 16, True, 16, 4, 7, 30000, "even"),        # two 16-image records, second one ragged
    (37, 24, 24, 12, True, 8, 4, 5, 12000, "even"),         # two 32-image records, second one ragged
]


@pytest.mark.parametrize("B,X,Y,A,pad,JW,NA,Rmax,budget,kind", WIN_CASES)
def test_emulated_windowed_forward_matches_oracle(emu, orc, B, X, Y, A, pad, JW, NA, Rmax, budget, kind):
    rng = np.random.default_rng(A * 131 + X)
    th = np.linspace(0, np.pi, A, endpoint=False) if kind == "even" else rng.uniform(-4, 4, A)
    img = rng.random((B, X, Y), dtype=np.float32)
    H, W, padx, pady = orc.frame_of(X, Y, pad)
    t = orc.make_transforms(th, H, W)
    emu.emu_forward_window.restype = emu.emu_forward_window32.restype = emu.emu_forward_window_wide.restype = ctypes.c_int
    for interp in (0, 1):
        want = orc.forward(img, th, pad, interp)
        for fn, bud in ((emu.emu_forward_window, budget), (emu.emu_forward_window32, 2 * budget), (emu.emu_forward_window_wide, 2 * budget)):
            s = np.full((B, A, W), np.nan, np.float32)
            nwin = fn(P(img), B, X, Y, H, W, padx, pady, P(t), A, interp, JW, NA, Rmax, bud, P(s))
            assert nwin >= 0, "shape did not fit the budget / rays not covered"
            if kind == "even":
                assert nwin > 0, "no chunk was actually windowed: the case tests nothing"
            assert not np.isnan(s).any(), "a sample fell outside its strip window"
            assert rel_l2(s, want) <= 1e-6


@settings(max_examples=25, deadline=None)
@given(X=st.integers(8, 72), Y=st.integers(8, 72), pad=st.booleans(), JW=st.sampled_from([8, 16, 24, 40]),
       NA=st.sampled_from([1, 2, 4]), Rmax=st.integers(2, 12), seed=st.integers(0, 10_000),
       th=st.lists(st.floats(-7.0, 7.0, allow_nan=False), min_size=1, max_size=8))
def test_emulated_windowed_forward_property(emu, orc, X, Y, pad, JW, NA, Rmax, seed, th):
    rng = np.random.default_rng(seed)
    th = np.asarray(th, np.float64)
    A = th.size
    img = rng.random((2, X, Y), dtype=np.float32)
    H, W, padx, pady = orc.frame_of(X, Y, pad)
    t = orc.make_transforms(th, H, W)
    emu.emu_forward_window.restype = ctypes.c_int
    for interp in (0, 1):
        s = np.full((2, A, W), np.nan, np.float32)
        nwin = emu.emu_forward_window(P(img), 2, X, Y, H, W, padx, pady, P(t), A, interp, JW, NA, Rmax, 1 << 20, P(s))
        assert nwin >= 0
        assert not np.isnan(s).any()
        assert rel_l2(s, orc.forward(img, th, pad, interp)) <= 1e-6


@pytest.mark.parametrize("nranks,B,X,Y,A", [(2, 4, 12, 12, 10), (4, 8, 9, 14, 7), (8, 16, 8, 8, 24), (3, 6, 10, 7, 5)])
def test_emulated_angle_sharded_exchange(emu, orc, nranks, B, X, Y, A):
    """The fused exchange of the angle-sharded adjoint (slot / owner index routines of ctr_core.h), all ranks emulated
    in one loop: the rank-ordered sum of the angle blocks' partials equals the full adjoint."""
    rng = np.random.default_rng(5)
    th = np.linspace(0, np.pi, A, endpoint=False)
    H, W, padx, pady = orc.frame_of(X, Y, True)
    t = orc.make_transforms(th, H, W)
    y = rng.random((B, A, W), dtype=np.float32)
    emu.emu_adjoint_sharded.restype = ctypes.c_int
    for interp in (0, 1):
        g = np.full((B, X, Y), np.nan, np.float32)
        rc = emu.emu_adjoint_sharded(P(y), B, X, Y, H, W, padx, pady, P(t), A, interp, nranks, P(g))
        assert rc == 0 and not np.isnan(g).any()
        assert rel_l2(g, orc.adjoint_exact(y, th, X, Y, True, interp)) <= 1e-6
