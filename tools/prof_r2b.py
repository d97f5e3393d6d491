"""One launch of each kernel changed late in round 2, for `ncu --set full` (developer tool):
nearest forward at C4 (two lanes per ray x 16 images), FBP at C4 and C5 (odd-tap row filter + 32-image gather)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ct_pvae_b200 as cp
from ct_pvae_b200 import _lib, ops
th = np.linspace(0, np.pi, 720, endpoint=False)
plan = _lib.get_plan(th, 512, 512, True, 0)
img = torch.rand((64, 512, 512), device="cuda")
ops.radon_forward(img, plan, 0)
for (B, X, A) in [(64, 512, 720), (1000, 128, 180)]:
    t = np.linspace(0, np.pi, A, endpoint=False)
    P = cp.num_proj_pix(X, X)
    fplan = _lib.get_fbp_plan(t, P, X, X, cp.get_fourier_filter(P, "ramp"), 0)
    y = torch.rand((B, A, P), device="cuda")
    ops.fbp(y, fplan)
torch.cuda.synchronize()
print("ok")
