"""Where does the host-buffer (e2e) path spend its time?  Developer probe."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ct_pvae_b200 as cp
from ct_pvae_b200 import _lib, ops, hostpipe

def wall(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3

B, X, A = 256, 128, 180
th = np.linspace(0, np.pi, A, endpoint=False)
plan = _lib.get_plan(th, X, X, True, 0)
img_h = torch.rand((B, X, X, 1)).pin_memory()
cot_h = torch.rand((B, A, plan.W)).pin_memory()
print("pinned alloc 34MB: %.3f ms" % wall(lambda: torch.empty((B, A, plan.W), pin_memory=True)))
keep = []
def alloc_keep():
    keep.append(torch.empty((B, A, plan.W), pin_memory=True))
    if len(keep) > 2: keep.pop(0)
print("pinned alloc 34MB (2 live): %.3f ms" % wall(alloc_keep))
print("fwd api e2e: %.3f ms" % wall(lambda: cp.project_tf_fast(img_h, th, pad=True, dim=2, integrate_vae=True, interpolation="bilinear")))
print("adj api e2e: %.3f ms" % wall(lambda: cp.backproject(cot_h, th, X, X, pad=True, interpolation="bilinear")))
x3 = img_h[..., 0]
for nch in (1, 2, 4, 8):
    f = lambda: hostpipe.run_chunked(lambda x: ops.radon_forward(x, plan, 1), x3, (B, A, plan.W), torch.device("cuda", 0), nchunks=nch)
    g = lambda: hostpipe.run_chunked(lambda y: ops.radon_adjoint(y, plan, 1, 0), cot_h, (B, X, X), torch.device("cuda", 0), nchunks=nch)
    print(f"nchunks={nch}: fwd {wall(f):.3f} ms  adj {wall(g):.3f} ms")
xd = x3.cuda(); yd = cot_h.cuda()
print("device-only fwd call: %.3f ms, adj call: %.3f ms" % (wall(lambda: ops.radon_forward(xd, plan, 1)), wall(lambda: ops.radon_adjoint(yd, plan, 1, 0))))
print("H2D img: %.3f ms; D2H sino: %.3f ms" % (wall(lambda: x3.to("cuda", non_blocking=True)), wall(lambda: cot_h.copy_(yd, non_blocking=True))))
print("plan lookup: %.4f ms" % wall(lambda: _lib.get_plan(ops.theta_to_host(th), X, X, True, 0), n=100))
def bench_like():
    s = cp.project_tf_fast(img_h, th, pad=True, dim=2, integrate_vae=True, interpolation="bilinear")
    g = cp.backproject(cot_h, th, X, X, pad=True, interpolation="bilinear", adjoint="exact")
    return s, g
hold = [None]
def bench_like_keep():
    hold[0] = bench_like()
print("bench-like step: %.3f ms ; keeping results: %.3f ms" % (wall(bench_like), wall(bench_like_keep)))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
print("bench-like with 256MB flush alive: %.3f ms" % wall(bench_like_keep))
