"""Host-buffer (e2e) step: blocking vs async_op calls, and the chunk size of the native pipeline.  Developer probe."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ct_pvae_b200 as cp
from ct_pvae_b200 import _lib, hostpipe
def wall(fn, n=30, warm=4):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
B, X, A = (int(v) for v in os.environ.get("SHAPE", "256,128,180").split(","))
th = np.linspace(0, np.pi, A, endpoint=False)
plan = _lib.get_plan(th, X, X, True, 0)
img_h = torch.rand((B, X, X, 1)).pin_memory()
cot_h = torch.rand((B, A, plan.W)).pin_memory()
hold=[None]
def sync_step():
    hold[0]=(cp.project_tf_fast(img_h, th, pad=True, dim=2, integrate_vae=True, interpolation="bilinear"),
      cp.backproject(cot_h, th, X, X, pad=True, interpolation="bilinear"))
def async_step():
    s,hs=cp.project_tf_fast(img_h, th, pad=True, dim=2, integrate_vae=True, interpolation="bilinear", async_op=True)
    g,hg=cp.backproject(cot_h, th, X, X, pad=True, interpolation="bilinear", async_op=True)
    hs.wait(); hg.wait(); hold[0]=(s,g)
for ch in os.environ.get("CHUNKS", "0:0,64:64,128:64,96:96").split(","):
    cf, ca = ch.split(":")
    hostpipe.set_chunk(int(cf), int(ca))
    print("chunk fwd:adj", ch, "blocking step %.3f ms   async step %.3f ms" % (wall(sync_step), wall(async_step)), flush=True)
