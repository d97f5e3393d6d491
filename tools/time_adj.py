"""Time the back-projection kernels alone (developer tool)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ct_pvae_b200 import _lib, ops
import ct_pvae_b200 as cp
shapes = ((256, 128, 180), (64, 512, 720))
if os.environ.get("TIME_ADJ_SHAPES"):
    shapes = [tuple(int(v) for v in s.split("x")) for s in os.environ["TIME_ADJ_SHAPES"].split(",")]
for (B, X, A) in shapes:
    th = np.linspace(0, np.pi, A, endpoint=False)
    plan = _lib.get_plan(th, X, X, True, 0)
    y = torch.rand((B, A, plan.W), device="cuda")
    fplan = _lib.get_fbp_plan(th, plan.W, X, X, cp.get_fourier_filter(plan.W, "ramp"), 0)
    res = {}
    for name, fn in (("exact/bil", lambda: ops.radon_adjoint(y, plan, 1, 0)), ("exact/near", lambda: ops.radon_adjoint(y, plan, 0, 0)),
                     ("tf/bil", lambda: ops.radon_adjoint(y, plan, 1, 1)), ("fbp", lambda: ops.fbp(y, fplan))):
        for _ in range(3): fn()
        torch.cuda.synchronize(); ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        res[name] = min(ts)
    print(f"B={B} X={X} A={A}: " + "  ".join(f"{k} {v:.3f}" for k, v in res.items()), flush=True)
