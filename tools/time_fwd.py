"""Time the forward kernel alone (developer tool)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ct_pvae_b200 import _lib, ops
shapes = [(256, 128, 180), (16, 512, 720)]
if os.environ.get("TIME_FWD_SHAPES"):
    shapes = [tuple(int(v) for v in s.split("x")) for s in os.environ["TIME_FWD_SHAPES"].split(",")]
for (B, X, A) in shapes:
    th = np.linspace(0, np.pi, A, endpoint=False)
    plan = _lib.get_plan(th, X, X, True, 0)
    img = torch.rand((B, X, X), device="cuda")
    out = []
    for iid in (1, 0):
        for _ in range(3): ops.radon_forward(img, plan, iid)
        torch.cuda.synchronize(); ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.radon_forward(img, plan, iid); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        out.append(min(ts))
    print(f"B={B} X={X} A={A} P={plan.W}: bilinear {out[0]:.3f} ms  nearest {out[1]:.3f} ms   {plan.describe(B)}", flush=True)
