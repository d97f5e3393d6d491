"""FBP experiment probe (developer tool): per-kernel times and a checksum of iradon for the C5 / C4 shapes."""
import os, sys, zlib
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ct_pvae_b200 as cp
from ct_pvae_b200 import _lib, ops
for (B, X, A) in [(1000, 128, 180), (64, 512, 720), (32, 128, 180)]:
    th = np.linspace(0, np.pi, A, endpoint=False)
    P = cp.num_proj_pix(X, X)
    filt = cp.get_fourier_filter(P, "ramp")
    plan = _lib.get_fbp_plan(th, P, X, X, filt, 0)
    g = torch.Generator(device="cuda").manual_seed(2)
    y = torch.rand((B, A, P), device="cuda", generator=g)
    for _ in range(2): o = ops.fbp(y, plan)
    torch.cuda.synchronize()
    _lib.profile_reset(); _lib.profile_enable(True)
    for _ in range(5): o = ops.fbp(y, plan)
    torch.cuda.synchronize()
    prof = _lib.profile_read(); _lib.profile_enable(False)
    print(f"B={B} X={X} A={A} P={P}: " +
          "  ".join(f"{k} {v[0] / v[1]:.3f} ms" for k, v in prof.items()) + f"  sum {float(o.double().sum()):.9e} crc {zlib.crc32(o.cpu().numpy().tobytes()):08x}", flush=True)
