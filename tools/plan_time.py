"""How long ctr_plan_create takes (host chunk / window tables + one upload) for the BASELINE shapes (developer probe)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ct_pvae_b200 import _lib

for (X, A) in ((128, 180), (512, 720), (512, 90), (1024, 1440)):
    th = np.linspace(0, np.pi, A, endpoint=False) + 1e-9 * np.random.rand()      # a fresh key: no cache hit
    t0 = time.perf_counter()
    p = _lib.Plan(th, X, X, True, 0)
    dt = time.perf_counter() - t0
    print(f"X={X} A={A}: plan created in {dt * 1e3:.1f} ms; {p.describe(64)[:90]}")
    p.close()
