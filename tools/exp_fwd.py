"""Forward-kernel experiment probe (developer tool): time + checksum of the forward for the standard shapes."""
import os, sys, zlib
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ct_pvae_b200 import _lib, ops
shapes = [(64, 512, 720), (256, 128, 180), (32, 128, 180)]
if os.environ.get("EXP_SHAPES"):
    shapes = [tuple(int(v) for v in s.split("x")) for s in os.environ["EXP_SHAPES"].split(",")]
for (B, X, A) in shapes:
    th = np.linspace(0, np.pi, A, endpoint=False)
    plan = _lib.get_plan(th, X, X, True, 0)
    g = torch.Generator(device="cuda").manual_seed(1)
    img = torch.rand((B, X, X), device="cuda", generator=g)
    res = []
    for iid in (1, 0):
        for _ in range(2): o = ops.radon_forward(img, plan, iid)
        torch.cuda.synchronize(); ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); o = ops.radon_forward(img, plan, iid); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        res.append((min(ts), zlib.crc32(o.cpu().numpy().tobytes())))
    print(f"B={B} X={X} A={A}: bilinear {res[0][0]:.3f} ms crc {res[0][1]:08x}  nearest {res[1][0]:.3f} ms crc {res[1][1]:08x}  {plan.describe(B)}", flush=True)
