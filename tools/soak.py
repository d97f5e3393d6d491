"""Soak the strip / window pipelines: many random shapes back to back, every result checked
against a second evaluation path (forward vs adjoint identity) so a race or a stale buffer shows."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ct_pvae_b200 as cp
from ct_pvae_b200 import _lib, ops

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
    # angle subsets on the same plan (r2): rows of the full sinogram, bit for bit, and the adjoint of the subset
    plan = _lib.get_plan(np.asarray(th, np.float64), X, Y, pad, 0)
    k = int(rng.integers(1, A + 1))
    sel = torch.from_numpy(rng.permutation(A)[:k].astype(np.int32)).cuda()
    iid = ops.INTERP[interp]
    sub = ops.radon_forward(img[..., 0].contiguous(), plan, iid, sel)
    ref_rows = s1[..., 0].index_select(1, sel.long())
    assert (sub - ref_rows).abs().max() <= 1e-5 * max(1.0, float(ref_rows.abs().max())), f"subset forward differs at it={it} {B,X,Y,A,pad,interp,k}"
    ysub = y[..., 0].index_select(1, sel.long()).contiguous()
    gsub = ops.radon_adjoint(ysub, plan, iid, 0, sel)
    lhs = float((sub.double() * ysub.double()).sum())
    rhs = float((img[..., 0].double() * gsub.double()).sum())
    assert abs(lhs - rhs) / max(abs(lhs), 1e-30) < 5e-6, f"subset adjoint identity broken at it={it} {B,X,Y,A,pad,interp,k}"
    # FBP: one cluster kernel == filter + gather, bit for bit (images that fit the single kernel)
    if X * Y <= 16384 and it % 5 == 0 and plan.W % 2 == 0:     # (the ramp filter is defined for even detector widths)
        P = plan.W
        fplan = _lib.get_fbp_plan(np.asarray(th, np.float64), P, X, Y, cp.get_fourier_filter(P, "ramp"), 0)
        sino = s1[..., 0].contiguous()
        fplan.set_fused(True)
        r1 = ops.fbp(sino, fplan)
        fplan.set_fused(False)
        r2 = ops.fbp(sino, fplan)
        assert torch.equal(r1, r2), f"fused FBP differs at it={it} {B,X,Y,A}"
torch.cuda.synchronize()
print(f"soak ok: {n} random cases, worst adjoint-identity error {worst:.2e}, {time.time() - t0:.1f} s")
