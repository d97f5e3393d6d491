"""Sweep the forward kernel's CTA shape through the CTR_FWD_* developer overrides."""
import itertools, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
from ct_pvae_b200 import _lib, ops
B, X, A = %d, %d, %d
th = np.linspace(0, np.pi, A, endpoint=False)
plan = _lib.get_plan(th, X, X, True, 0)
img = torch.rand((B, X, X), device="cuda")
for iid in (1, 0):
    for _ in range(3): ops.radon_forward(img, plan, iid)
    torch.cuda.synchronize(); ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.radon_forward(img, plan, iid); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print("%%s %%.3f" %% ("bilinear" if iid else "nearest", min(ts)), end="  ")
print()
'''
def run(tag, B, X, A, **env):
    e = dict(os.environ); e.update({k: str(v) for k, v in env.items()})
    out = subprocess.run([sys.executable, "-c", CODE % (ROOT, B, X, A)], env=e, capture_output=True, text=True)
    print(tag, env, out.stdout.strip(), out.stderr.strip()[-200:], flush=True)
if __name__ == "__main__":
    for ns, ka, smem in [(4, 1, 230000), (4, 1, 110000), (2, 1, 110000), (2, 2, 110000), (2, 1, 75000), (2, 2, 75000), (1, 2, 75000), (1, 4, 75000), (1, 2, 55000), (1, 4, 55000), (3, 1, 110000)]:
        run("C2", 256, 128, 180, CTR_FWD_NS=ns, CTR_FWD_KA=ka, CTR_FWD_SMEM=smem)
    for r in (4, 8, 16, 24):
        run("C2", 256, 128, 180, CTR_FWD_NS=2, CTR_FWD_KA=1, CTR_FWD_SMEM=110000, CTR_FWD_R=r)
    for ka, r in [(4, 12), (2, 12), (1, 12), (4, 6), (2, 6), (4, 3)]:
        run("C4", 16, 512, 720, CTR_FWD_KA=ka, CTR_FWD_R=r)
