"""Host-side cost of one projector call (tiny problem, GPU time negligible)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ct_pvae_b200 import _lib, ops
import ct_pvae_b200 as cp
th = np.linspace(0, np.pi, 4, endpoint=False)
plan = _lib.get_plan(th, 16, 16, True, 0)
img = torch.rand((4, 16, 16), device="cuda")
def wall(fn, n=2000):
    for _ in range(50): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
print("ops.radon_forward (DLPack path): %.1f us/call" % wall(lambda: ops.radon_forward(img, plan, 1)))
sino = torch.empty((4, 4, plan.W), device="cuda"); ws = torch.empty(plan.forward_workspace_bytes(4), dtype=torch.uint8, device="cuda")
L = _lib.lib(); st = torch.cuda.current_stream().cuda_stream
print("raw ctr_radon_forward (pointers, preallocated): %.1f us/call" % wall(lambda: L.ctr_radon_forward(plan.handle, img.data_ptr(), sino.data_ptr(), 4, 1, ws.data_ptr(), ws.numel(), st)))
print("torch.empty x2: %.1f us" % wall(lambda: (torch.empty((4, 4, plan.W), device="cuda"), torch.empty(4096, dtype=torch.uint8, device="cuda"))))
print("DLView x3: %.1f us" % wall(lambda: (_lib.DLView(img), _lib.DLView(sino), _lib.DLView(ws))))
x4 = img.unsqueeze(-1)
print("cp.project_tf_fast (public API, CUDA tensor): %.1f us/call" % wall(lambda: cp.project_tf_fast(x4, th, pad=True, dim=2, integrate_vae=True)))
