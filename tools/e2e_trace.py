import os, sys, time

import numpy as np, torch
sys.path.insert(0, os.getcwd())
import ct_pvae_b200 as cp
from ct_pvae_b200 import hostpipe
hostpipe.set_trace(True)
B, X, A = 256,128,180
th = np.linspace(0, np.pi, A, endpoint=False)
img_h = torch.rand((B, X, X, 1)).pin_memory()
cot_h = torch.rand((B, A, 184)).pin_memory()
for it in range(3):
    torch.cuda.synchronize()
    t0=time.perf_counter()
    s,hs=cp.project_tf_fast(img_h, th, pad=True, dim=2, integrate_vae=True, interpolation="bilinear", async_op=True)
    t1=time.perf_counter()
    g,hg=cp.backproject(cot_h, th, X, X, pad=True, interpolation="bilinear", async_op=True)
    t2=time.perf_counter()
    print("--- iter",it, "issue fwd %.3f adj %.3f ms"%((t1-t0)*1e3,(t2-t1)*1e3), file=sys.stderr)
    hs.wait(); hg.wait()
    print("total %.3f"%((time.perf_counter()-t0)*1e3), file=sys.stderr)
