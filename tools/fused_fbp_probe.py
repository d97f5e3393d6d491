import sys, numpy as np, torch
sys.path.insert(0, ".")
import ct_pvae_b200 as cp
from ct_pvae_b200 import _lib, ops
B, X, A = 1000, 128, 180
P = cp.num_proj_pix(X, X)
th = np.linspace(0, np.pi, A, endpoint=False)
plan = _lib.get_fbp_plan(th, P, X, X, cp.get_fourier_filter(P, "ramp"), 0)
sino = torch.rand((B, A, P), device="cuda")
plan.set_fused(True)
for _ in range(3):
    ops.fbp(sino, plan)
torch.cuda.synchronize()
