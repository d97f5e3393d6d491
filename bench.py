#!/usr/bin/env python3
"""bench.py -- Radon forward + adjoint throughput on B200 (BASELINE.json's metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c4] [--impl ours|reference]

A *step* is one pass of the hot path over one batch of synthetic foam images: the
ray-driven forward projection (bilinear) of B images at A angles, then the exact
gather adjoint of a sinogram-shaped cotangent -- what one training iteration's
projector forward + backward costs (reference: helper_functions.py:359 inside the
tape of main_ct_vae.py:471-481).  `value` is ray-sums per second over the whole job:
B*A*P ray-sums (each projected once and back-projected once) / step time.

N > 1 (torchrun, one rank per GPU): the batch is sharded -- every rank owns its own B
images (weak scaling, no data-path collective).  `--shard angle` switches to the
angle-sharded mode of SURVEY 8e (every rank holds all B images and A/N angles; the
partial back-projections are summed with one NCCL all-reduce).

`--impl reference` times the reference's CPU dataflow (oracle port; TensorFlow itself
is not installable in this image) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on
    "c2": dict(B=256, X=128, A=180, name="configs[1] projector/adjoint microbench: foam 128x128, 180 angles, batch 256"),
    # BASELINE.json configs[3]: large-scale operator sweep
    "c4": dict(B=64, X=512, A=720, name="configs[3] operator sweep: 512x512, 720 angles, batch 64"),
}
METRIC = "Radon fwd+adjoint Gray-sums/s"
UNIT = "Gray-sums/s"
INTERP = "bilinear"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def ncu_traffic(workload, kernel):
    """Per-launch DRAM bytes of `kernel` from the committed ncu capture (profiles/), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_dram_traffic.json")) as f:
            return json.load(f)[workload].get(kernel)
    except Exception:
        return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index: int):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(self.NAMES, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------- data
def synthetic_foam_torch(B, X, device, seed):
    """Unit disk with random zero-valued circular pores (stand-in for xdesign.Foam,
    scripts/create_foam_images.py:27-40; xdesign is not installable here)."""
    import torch

    g = torch.Generator(device="cpu").manual_seed(seed)
    lin = torch.linspace(-1, 1, X)
    yy, xx = torch.meshgrid(lin, lin, indexing="ij")
    disk = (xx * xx + yy * yy <= 1.0).float()
    K = 24
    r = torch.rand((B, K, 1, 1), generator=g) * 0.19 + 0.01
    ang = torch.rand((B, K, 1, 1), generator=g) * 2 * np.pi
    rad = torch.sqrt(torch.rand((B, K, 1, 1), generator=g)) * (1 - r)
    keep = (torch.rand((B, K, 1, 1), generator=g) < torch.rand((B, 1, 1, 1), generator=g)).float()
    cx, cy = rad * torch.cos(ang), rad * torch.sin(ang)
    out = torch.empty((B, X, X), dtype=torch.float32)
    for b0 in range(0, B, 32):
        sl = slice(b0, min(B, b0 + 32))
        pore = (((xx - cx[sl]) ** 2 + (yy - cy[sl]) ** 2) <= r[sl] ** 2).float() * keep[sl]
        out[sl] = disk * (1 - pore.amax(dim=1))
    return out.to(device)


# ----------------------------------------------------------------------------------------- CPU arm
def cpu_reference_pass(B, X, A, nb, seed=0):
    """One bounded sample of the reference's CPU dataflow: project_tf_fast's
    pad -> repeat -> rotate -> row-sum graph (bilinear) and TensorFlow's gradient of it,
    restated in oracle/radon_oracle.c with OpenMP over all host cores.  -> seconds."""
    from oracle import radon_oracle as orc

    rng = np.random.default_rng(seed)
    theta = np.linspace(0, np.pi, A, endpoint=False)
    img = rng.random((nb, X, X), dtype=np.float32)
    W = orc.frame_of(X, X, True)[1]
    cot = rng.random((nb, A, W), dtype=np.float32)
    t0 = time.perf_counter()
    orc.forward(img, theta, True, orc.BILINEAR, dataflow=True)
    orc.adjoint_tf(cot, theta, X, X, True, orc.BILINEAR)
    return time.perf_counter() - t0, nb * A * W


def _max_sample(wl, cores):
    """Largest image count whose per-thread rotated stack [H,W,nb] keeps the port under ~8 GB."""
    P = 2 * int(np.ceil((np.sqrt(2.0) * wl["X"] + 2) / 2))
    return int(max(1, min(wl["B"], 8e9 / (cores * P * P * 4.0))))


def cpu_baseline(wl, target_s=10.0):
    from oracle import radon_oracle as orc

    cores = int(orc.lib().orc_max_threads())
    cap = _max_sample(wl, cores)
    nb = max(1, min(cap, 2))
    cpu_reference_pass(wl["B"], wl["X"], wl["A"], 1)        # first touch: library load, page-in
    t, units = cpu_reference_pass(wl["B"], wl["X"], wl["A"], nb)
    if t < target_s / 2:
        nb = int(max(1, min(cap, nb * target_s / max(t, 1e-3))))
        t, units = cpu_reference_pass(wl["B"], wl["X"], wl["A"], nb)
    return {"value": units / t / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{nb} of {wl['B']} images, all {wl['A']} angles, fwd dataflow + TF-style gradient, {t:.2f} s"}


def run_reference(args, wl):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    from oracle import radon_oracle as orc

    # all host threads the box has (torchrun pins OMP_NUM_THREADS=1 per rank; the other ranks do no work here)
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    orc.lib().orc_set_num_threads(int(avail))
    cores = int(orc.lib().orc_max_threads())
    # size the per-step sample so warmup+steps stay within a few minutes
    cpu_reference_pass(wl["B"], wl["X"], wl["A"], 1)        # first touch: library load, page-in
    t_probe, _ = cpu_reference_pass(wl["B"], wl["X"], wl["A"], 2)
    budget = 120.0 / max(1, args.steps + args.warmup)
    nb = int(max(1, min(_max_sample(wl, cores), 2 * budget / max(t_probe, 1e-3))))
    for _ in range(args.warmup):
        cpu_reference_pass(wl["B"], wl["X"], wl["A"], nb)
    tot, units = 0.0, 0
    for _ in range(args.steps):
        t, u = cpu_reference_pass(wl["B"], wl["X"], wl["A"], nb)
        tot += t
        units += u
    value = units / tot / 1e9
    P = orc.frame_of(wl["X"], wl["X"], True)[1]
    sample = f"{nb} of {wl['B']} images per step, all {wl['A']} angles"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": tot / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (uniform random images)",
        "config": {"workload": wl["name"], "interpolation": INTERP, "adjoint": "TF gradient", "B": wl["B"],
                   "X": wl["X"], "Y": wl["X"], "A": wl["A"], "P": P, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def vae_training_leg(dev, iters=20, warm=3, seed=0, world=1):
    """End-to-end VAE training it/s on BASELINE configs[0] (README quick-start: 50 foam images
    128x128, 180 angles, -b 5 --nsa 20 --api 20 --ns 2 --pnm 1e4 --normal --random), with the
    torch restatement of the reference's networks/loss (ct_pvae_b200/vae.py) around the fused
    projector.  Random-init weights, synthetic foam; wall clock with a sync on both sides."""
    import torch

    from ct_pvae_b200 import vae

    torch.manual_seed(0)            # identical initial weights on every rank
    N, X, A, b, nsa, api, ns, pnm = 50, 128, 180, 5, 20, 20, 2, 1e4
    theta = np.linspace(0, np.pi, A, endpoint=False)
    imgs = synthetic_foam_torch(N, X, dev, seed=123 + seed)
    sino = vae.create_sinogram(imgs, theta, pad=True, interpolation="bilinear")
    masks, meas = vae.create_all_masks(sino, A, pnm, num_sparse_angles=nsa, random=True)
    enc_in = vae.iradon_all(meas, masks, theta, X, X)
    model = vae.CTVAE(X, X, num_filters=1).to(dev)
    g = torch.Generator().manual_seed(1 + seed)
    ga = torch.Generator().manual_seed(7)      # the angle minibatch is shared by all ranks

    def one():
        idx = torch.randint(0, N, (b,), generator=g).to(dev)
        angles_i = torch.randperm(A, generator=ga)[:api]
        loss, _, _, _ = model.train_step(meas[idx], masks[idx], enc_in[idx], pnm, theta, angles_i=angles_i, num_samples=ns)
        return loss

    for _ in range(warm):
        one()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(iters):
        loss = one()
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    return {"it_per_s": iters / dt, "ms_per_it": dt / iters * 1e3, "final_loss": float(loss), "global_batch": b * world,
            "config": "README quick-start: b=5 per GPU, 128x128, 180 angles, nsa=20, api=20, ns=2, pnm=1e4, --normal --random"
                      + ("; batch-sharded, gradients averaged with one NCCL all-reduce" if world > 1 else "")}


# ----------------------------------------------------------------------------------------- GPU arm
def run_ours(args, wl):
    import torch
    import torch.distributed as dist

    import ct_pvae_b200 as cp
    from ct_pvae_b200 import _lib, ops, sharding

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    json_fd = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", ""):
            os.environ["NCCL_DEBUG"] = "WARN"
        # NCCL writes its version banner to file descriptor 1 when the first communicator comes up (seen in the
        # angle-sharded run): point fd 1 at stderr for the whole run and keep the real stdout for the ONE JSON line
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)

    B, X, A = wl["B"], wl["X"], wl["A"]
    theta = np.linspace(0, np.pi, A, endpoint=False)
    angle_mode = args.shard == "angle" and world > 1
    if angle_mode:
        a_lo, a_hi = sharding.shard_range(A, rank, world)
        theta_local = theta[a_lo:a_hi]
    else:
        theta_local = theta
    plan = _lib.get_plan(theta_local, X, X, True, local)
    P, A_loc = plan.W, plan.A
    img = synthetic_foam_torch(B, X, dev, seed=(0 if angle_mode else rank))
    cot = torch.rand((B, A_loc, P), device=dev, generator=torch.Generator(device=dev).manual_seed(1 + rank))
    iid, mid = ops.INTERP[INTERP], ops.ADJOINT["exact"]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step():
        s = ops.radon_forward(img, plan, iid)
        if not angle_mode:
            return s, ops.radon_adjoint(cot, plan, iid, mid)
        # angle-sharded: sum the partial back-projections over the angle shards.  The batch is cut in
        # two so the NCCL all-reduce of one half overlaps the adjoint kernels of the other.
        half = max(32, (B // 2 + 31) // 32 * 32)
        parts, works = [], []
        for lo in range(0, B, half):
            g = ops.radon_adjoint(cot[lo:lo + half], plan, iid, mid)
            if args.angle_collective == "reduce_scatter" and g.shape[0] % world == 0:
                # result stays batch-sharded: rank r ends up with the summed images [r*n/world, (r+1)*n/world) of this half
                out = torch.empty((g.shape[0] // world,) + tuple(g.shape[1:]), dtype=g.dtype, device=g.device)
                works.append(dist.reduce_scatter_tensor(out, g, async_op=True))
                parts.append((out, g))      # keep the input alive until the collective has run
            else:
                works.append(dist.all_reduce(g, async_op=True))
                parts.append(g)
        for w in works:
            w.wait()
        return s, parts

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    warm = max(args.warmup, 3)   # timing rule: at least 3 warm-up steps
    for _ in range(warm):
        step()
    sync_all()

    # ---- timed region: K steps, per-step CUDA events, L2 flushed between steps
    sampler = ClockSampler(local) if rank == 0 else None
    _lib.profile_reset()
    _lib.profile_enable(True)
    launches0 = _lib.launch_count()
    evs = []
    sync_all()
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        evs.append((e0, e1))
    sync_all()
    launches = _lib.launch_count() - launches0
    _lib.profile_enable(False)
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    prof = _lib.profile_read()
    _lib.profile_reset()
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    units_per_step = B * A * P * (1 if angle_mode else world)   # whole-job ray-sums per step
    value = units_per_step / (ms_per_step * 1e-3) / 1e9

    # ---- side legs (reported under "extra", not part of the headline step): the reference's
    # default nearest mode, TF-compatible gradient, and the FBP evaluation pass on the same batch
    def best_ms(fn, iters=5):
        fn()
        torch.cuda.synchronize(dev)
        ts = []
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize(dev)
            ts.append(a.elapsed_time(b))
        return float(np.median(ts))

    side = {}
    if rank == 0 and not args.no_side_legs:
        nid = ops.INTERP["nearest"]
        side["fwd_nearest_ms"] = best_ms(lambda: ops.radon_forward(img, plan, nid))
        side["adj_exact_nearest_ms"] = best_ms(lambda: ops.radon_adjoint(cot, plan, nid, mid))
        side["adj_tf_compat_bilinear_ms"] = best_ms(lambda: ops.radon_adjoint(cot, plan, iid, ops.ADJOINT["tf_compat"]))
        fplan = _lib.get_fbp_plan(theta_local, P, X, X, cp.get_fourier_filter(P, "ramp"), local)
        side["fbp_ms"] = best_ms(lambda: ops.fbp(cot, fplan))
        side["fbp_gupdates_per_s"] = B * A_loc * X * X / (side["fbp_ms"] * 1e-3) / 1e9
        side["fwd_nearest_gray_sums_per_s"] = B * A_loc * P / (side["fwd_nearest_ms"] * 1e-3) / 1e9
    if not args.no_side_legs and not args.no_train_leg and (rank == 0 or world > 1):
        # world > 1: every rank trains on its own batch slice, gradients averaged over NCCL
        try:
            leg = vae_training_leg(dev, seed=rank, world=world)
            if rank == 0:
                side["vae_train"] = leg
        except Exception as exc:  # the restated VAE is a caller, never a reason to lose the headline
            side["vae_train"] = {"error": repr(exc)[:200]}
    sync_all()

    # ---- e2e: the public API with pinned HOST buffers, copies inside the timed region
    img_h = img.cpu().unsqueeze(-1).pin_memory()
    cot_h = cot.cpu().pin_memory()
    def e2e_step():
        # both calls are issued with async_op=True (pinned result + completion handle, like a torch.distributed
        # work handle) so the forward's copy-out overlaps the adjoint's copy-in; the step ends when both
        # results are in host memory
        s, hs = cp.project_tf_fast(img_h, theta_local, pad=True, dim=2, integrate_vae=True, interpolation=INTERP,
                                   async_op=True)
        g, hg = cp.backproject(cot_h, theta_local, X, X, pad=True, interpolation=INTERP, adjoint="exact", async_op=True)
        hs.wait()
        hg.wait()
        return s, g
    for _ in range(3):           # warm up holding the results like the timed loop does, so the
        s_h, g_h = e2e_step()    # pinned-host allocator has every block it will hand out
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        s_h, g_h = e2e_step()
    torch.cuda.synchronize(dev)
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = units_per_step * args.steps / float(t_e2e.item()) / 1e9
    h2d = img_h.numel() * 4 + cot_h.numel() * 4
    d2h = s_h.numel() * 4 + g_h.numel() * 4

    if rank == 0:
        peak, peak_src = measured_peaks()
        # dominant kernel of the step and its HBM roofline (compulsory bytes)
        dom = max(prof.items(), key=lambda kv: kv[1][0]) if prof else (None, (0.0, 1))
        alg_bytes = 4.0 * B * (X * X + A_loc * P)     # forward: image in + sinogram out; adjoint: the mirror
        dom_ms = dom[1][0] / max(1, dom[1][1])
        achieved = alg_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        kern = {k: {"ms_per_launch": v[0] / v[1], "launches": v[1], "share_of_step": v[0] / total_ms} for k, v in prof.items()}
        fwd_ms = kern.get("ctr_fwd_kernel", {}).get("ms_per_launch")
        adj_ms = kern.get("ctr_bp_kernel<exact>", {}).get("ms_per_launch")
        sm_hz = (clocks or {}).get("sm_mhz") or 1965.0
        smem_peak = 148 * 128 * sm_hz * 1e6 / 1e9     # GB/s of shared-memory reads at the sampled clock
        extra = {
            "kernels": kern,
            "side_legs": side,
            "fwd_gray_sums_per_s": (B * A_loc * P / (fwd_ms * 1e-3) / 1e9) if fwd_ms else None,
            "adjoint_gupdates_per_s": (B * A_loc * X * X / (adj_ms * 1e-3) / 1e9) if adj_ms else None,
            # binding on-chip limit of the forward gather: 16 B of shared memory per in-support bilinear sample
            # same for the exact adjoint: 3 candidate bins x 4 B of shared memory per pixel-angle update and image
            "adj_smem_roofline": ({"achieved_GBs": 12.0 * B * A_loc * X * X / (adj_ms * 1e-3) / 1e9, "peak_GBs": smem_peak,
                                   "frac": 12.0 * B * A_loc * X * X / (adj_ms * 1e-3) / 1e9 / smem_peak} if adj_ms else None),
            "fwd_smem_roofline": ({"achieved_GBs": 16.0 * B * A_loc * (X + 1) ** 2 / (fwd_ms * 1e-3) / 1e9,
                                   "peak_GBs": smem_peak, "frac": 16.0 * B * A_loc * (X + 1) ** 2 / (fwd_ms * 1e-3) / 1e9 / smem_peak}
                                  if fwd_ms else None),
        }
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if angle_mode else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic foam (unit disk, random circular pores), random cotangents",
            "config": {"workload": wl["name"], "interpolation": INTERP, "adjoint": "exact", "B_per_gpu": B, "X": X, "Y": X,
                       "A": A, "P": P, "sharding": ("angle/" + args.angle_collective) if angle_mode else "batch",
                       "l2": "flushed (256 MiB memset) between timed steps"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": dom[0], "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(args.workload, dom[0]) if not angle_mode else None,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "gather/issue-bound stencil: HBM fraction is small by construction (SURVEY 8d); see extra.fwd_smem_roofline"},
            "extra": extra,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(wl)
            except Exception as exc:  # the checker library is optional for the GPU numbers
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": f"failed: {exc}"}
        if json_fd is None:
            print(json.dumps(line), flush=True)
        else:
            sys.stdout.flush()
            os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--shard", default="batch", choices=["batch", "angle"])
    ap.add_argument("--angle-collective", default="all_reduce", choices=["all_reduce", "reduce_scatter"],
                    help="angle-sharded mode: how the partial back-projections are summed")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side-legs", action="store_true")
    ap.add_argument("--no-train-leg", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
