"""TF-compat gradient probe (developer tool): time + checksum."""
import os, sys, zlib
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ct_pvae_b200 import _lib, ops
for (B, X, A) in [(64, 512, 720), (256, 128, 180), (32, 128, 180)]:
    th = np.linspace(0, np.pi, A, endpoint=False)
    plan = _lib.get_plan(th, X, X, True, 0)
    g = torch.Generator(device="cuda").manual_seed(3)
    y = torch.rand((B, A, plan.W), device="cuda", generator=g)
    for _ in range(2): o = ops.radon_adjoint(y, plan, 1, 1)
    torch.cuda.synchronize(); ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); o = ops.radon_adjoint(y, plan, 1, 1); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"B={B} X={X} A={A}: tf_compat bilinear {min(ts):.3f} ms crc {zlib.crc32(o.cpu().numpy().tobytes()):08x}", flush=True)
