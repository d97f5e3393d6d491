"""Cost of the C4 forward / exact adjoint per angle block on ONE GPU (developer probe behind
sharding.balanced_angle_blocks): times the kernels for each of the 2N contiguous blocks, then for the contiguous and the
balanced per-rank angle sets of an N-rank angle-sharded run."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ct_pvae_b200 import _lib, ops, sharding

B, X, A, N = 64, 512, 720, int(os.environ.get("RANKS", "8"))
theta = np.linspace(0, np.pi, A, endpoint=False)
img = torch.rand((B, X, X), device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def time_set(idx):
    plan = _lib.get_plan(theta[idx], X, X, True, 0)
    cot = torch.rand((B, len(idx), plan.W), device="cuda")
    out = []
    for fn in (lambda: ops.radon_forward(img, plan, 1), lambda: ops.radon_adjoint(cot, plan, 1, 0)):
        for _ in range(2):
            fn()
        ts = []
        for _ in range(5):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        out.append(float(np.median(ts)))
    return out


print(f"{2 * N} contiguous blocks of {A // (2 * N)} angles: fwd ms / adj ms")
for b in range(2 * N):
    lo, hi = sharding.shard_range(A, b, 2 * N)
    f, a = time_set(np.arange(lo, hi))
    print(f"  block {b:2d} [{theta[lo] * 180 / np.pi:6.1f} deg ..): {f:.3f} / {a:.3f}")
for name in ("contiguous", "balanced"):
    tot = []
    for r in range(N):
        idx = np.arange(*sharding.shard_range(A, r, N)) if name == "contiguous" else sharding.balanced_angle_blocks(theta, r, N)
        f, a = time_set(idx)
        tot.append(f + a)
        print(f"  {name} rank {r}: fwd {f:.3f} adj {a:.3f} sum {f + a:.3f}")
    print(f"{name}: max {max(tot):.3f} mean {np.mean(tot):.3f} ms  (imbalance {max(tot) / np.mean(tot) - 1:.1%})")
