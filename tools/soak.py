"""Soak the strip / window pipelines: many random shapes back to back, every result checked
against a second evaluation path (forward vs adjoint identity) so a race or a stale buffer shows."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ct_pvae_b200 as cp

rng = np.random.default_rng(int(os.environ.get("SOAK_SEED", "0")))
n = int(os.environ.get("SOAK_ITERS", "300"))
t0 = time.time()
worst = 0.0
for it in range(n):
    big = it % 4 == 3       # every fourth case: wide detector (column-windowed 16/32-image shapes)
    hi = 420 if big else 220
    B, X, Y, A = int(rng.integers(1, 70)), int(rng.integers(2, hi)), int(rng.integers(2, hi)), int(rng.integers(1, 48))
    pad = bool(rng.integers(0, 2))
    interp = ("nearest", "bilinear")[int(rng.integers(0, 2))]
    # random angles (sparse: wide windows / fallbacks) or an evenly spaced fan with a random offset (real windows)
    th = rng.uniform(-4, 4, A) if it % 2 else np.linspace(0, np.pi, A, endpoint=False) + rng.uniform(-1, 1)
    img = torch.rand((B, X, Y, 1), device="cuda")
    s1 = cp.project_tf_fast(img, th, pad=pad, dim=2, integrate_vae=True, interpolation=interp)
    s2 = cp.project_tf_fast(img, th, pad=pad, dim=2, integrate_vae=True, interpolation=interp)
    assert torch.equal(s1, s2), f"non-deterministic forward at it={it} {B,X,Y,A,pad,interp}"
    y = torch.rand_like(s1)
    g = cp.backproject(y, th, X, Y, pad=pad, interpolation=interp)
    lhs = float((s1.double() * y.double()).sum())
    rhs = float((img.double() * g.double()).sum())
    err = abs(lhs - rhs) / max(abs(lhs), 1e-30)
    worst = max(worst, err)
    assert err < 5e-6, f"adjoint identity broken at it={it} {B,X,Y,A,pad,interp}: {err}"
torch.cuda.synchronize()
print(f"soak ok: {n} random cases, worst adjoint-identity error {worst:.2e}, {time.time() - t0:.1f} s")
