"""Quick CUDA-event timings of each kernel path (developer tool, not the bench contract)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ct_pvae_b200 as cp  # noqa: E402
from ct_pvae_b200 import _lib, ops  # noqa: E402


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))


def run(B, X, A, tag):
    th = np.linspace(0, np.pi, A, endpoint=False)
    plan = _lib.get_plan(th, X, X, True, 0)
    img = torch.rand((B, X, X), device="cuda")
    y = torch.rand((B, A, plan.W), device="cuda")
    rays = B * A * plan.W
    upd = B * A * X * X
    for name, iid in (("nearest", 0), ("bilinear", 1)):
        best, med = timeit(lambda: ops.radon_forward(img, plan, iid))
        print(f"{tag} fwd {name:8s}: {best:8.3f} ms (med {med:8.3f})  {rays / best / 1e6:8.2f} G ray-sums/s  "
              f"{B * A * (X + 1) ** 2 / best / 1e6:9.1f} G samples/s", flush=True)
        for mode, mid in (("exact", 0), ("tf_compat", 1)):
            best, med = timeit(lambda: ops.radon_adjoint(y, plan, iid, mid))
            print(f"{tag} adj {name:8s} {mode:9s}: {best:8.3f} ms (med {med:8.3f})  {upd / best / 1e6:8.2f} G updates/s", flush=True)
    filt = cp.get_fourier_filter(plan.W, "ramp")
    fplan = _lib.get_fbp_plan(th, plan.W, X, X, filt, 0)
    best, med = timeit(lambda: ops.fbp(y, fplan))
    print(f"{tag} fbp: {best:8.3f} ms (med {med:8.3f})  {upd / best / 1e6:8.2f} G updates/s", flush=True)


def pcie():
    n = 64 << 20
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    s2 = torch.cuda.Stream()
    for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
        best, med = timeit(fn, iters=5, warm=2)
        print(f"PCIe {name}: {n / best / 1e6:.1f} GB/s", flush=True)
    h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
    def both():
        s2.wait_stream(torch.cuda.current_stream())
        d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s2)
    best, med = timeit(both, iters=5, warm=2)
    print(f"PCIe duplex: {2 * n / best / 1e6:.1f} GB/s aggregate", flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    pcie()
    run(256, 128, 180, "C2")
    run(64, 512, 720, "C4")


def loglik_time(B=256, X=128, A=180):
    from ct_pvae_b200 import likelihood
    th = np.linspace(0, np.pi, A, endpoint=False)
    plan = _lib.get_plan(th, X, X, True, 0)
    img = torch.rand((B, X, X), device="cuda")
    mask = (torch.rand((B, A), device="cuda") < 0.2).float() / 20
    meas = torch.rand((B, A, plan.W), device="cuda")
    best, med = timeit(lambda: ops.radon_loglik(img, plan, mask, meas, None, 1e4, 1.19e-7, 0))
    print(f"fused loglik (nearest) fwd: {best:.3f} ms", flush=True)
    x4 = img.unsqueeze(-1)
    best, med = timeit(lambda: likelihood.calculate_log_prob_M_given_R(x4, mask, meas, 1e4, 1.19e-7, theta=th, pad=True).sum())
    print(f"unfused (project + torch elementwise + sum) fwd: {best:.3f} ms", flush=True)
    xg = x4.clone().requires_grad_(True)
    def fused_fb():
        xg.grad = None
        likelihood.log_prob_M_given_R_sum(xg, mask, meas, 1e4, 1.19e-7, theta=th, pad=True).backward()
    def unfused_fb():
        xg.grad = None
        likelihood.calculate_log_prob_M_given_R(xg, mask, meas, 1e4, 1.19e-7, theta=th, pad=True).sum().backward()
    print("fused fwd+bwd: %.3f ms ; unfused fwd+bwd: %.3f ms" % (timeit(fused_fb)[0], timeit(unfused_fb)[0]), flush=True)


if __name__ == "__main__":
    loglik_time()
