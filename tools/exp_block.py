"""Forward / adjoint time of per-rank angle blocks of C4 on one GPU (developer probe)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ct_pvae_b200 import _lib, ops, sharding
B, X, A = 64, 512, 720
theta = np.linspace(0, np.pi, A, endpoint=False)
img = torch.rand((B, X, X), device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for N in (8, 4, 2, 1):
    tot = []
    for r in range(N):
        lo, hi = sharding.cost_balanced_range(theta, r, N) if hasattr(sharding, "cost_balanced_range") and N > 1 else sharding.shard_range(A, r, N)
        plan = _lib.get_plan(theta[lo:hi], X, X, True, 0)
        for _ in range(2): ops.radon_forward(img, plan, 1)
        ts = []
        for _ in range(5):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.radon_forward(img, plan, 1); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        tot.append(float(np.median(ts)))
        if r == 0: desc = plan.describe(B)
    print(f"N={N}: fwd per rank " + " ".join(f"{t:.3f}" for t in tot) + f"  max {max(tot):.3f} sum {sum(tot):.3f}   {desc[:110]}", flush=True)
