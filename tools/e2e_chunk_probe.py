"""e2e (host buffers) step time vs chunk size of the host pipeline at C4 / C2 (developer probe behind hostpipe.chunk_for)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ct_pvae_b200 as cp
from ct_pvae_b200 import hostpipe

for (B, X, A, chunks) in ((64, 512, 720, (0, 16, 32, 64)), (256, 128, 180, (0, 32, 64, 128))):
    th = np.linspace(0, np.pi, A, endpoint=False)
    P = cp.num_proj_pix(X, X)
    img_h = torch.rand((B, X, X, 1)).pin_memory()
    cot_h = torch.rand((B, A, P)).pin_memory()
    for ch in chunks:
        hostpipe.set_chunk(ch, ch)

        def step():
            s, hs = cp.project_tf_fast(img_h, th, pad=True, dim=2, integrate_vae=True, interpolation="bilinear", async_op=True)
            g, hg = cp.backproject(cot_h, th, X, X, pad=True, interpolation="bilinear", async_op=True)
            hs.wait(); hg.wait()
            return s, g

        for _ in range(3):
            keep = step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 10
        for _ in range(n):
            keep = step()
        torch.cuda.synchronize()
        print(f"B={B} X={X} A={A} chunk={ch or 'auto'}: {(time.perf_counter() - t0) / n * 1e3:.3f} ms per e2e step", flush=True)
hostpipe.set_chunk(0, 0)
