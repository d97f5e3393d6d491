#!/usr/bin/env python3
"""bench.py -- Radon forward + adjoint throughput on B200 (BASELINE.json's metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c2|c5] [--shard batch|angle]
                  [--impl ours|reference]

A *step* is one pass of the hot path over one batch of synthetic foam images: the
ray-driven forward projection (bilinear) of B images at A angles, then the exact
gather adjoint of a sinogram-shaped cotangent -- what one training iteration's
projector forward + backward costs (reference: helper_functions.py:359 inside the
tape of main_ct_vae.py:471-481).  `value` is ray-sums per second over the whole job:
B*A*P ray-sums (each projected once and back-projected once) / step time.

Default workload: BASELINE.json configs[3] (C4: 64 x 512^2 x 720 angles, P = 728) -- the
configuration north_star quotes its target on.  configs[1] (C2: 256 x 128^2 x 180) and
configs[4] (C5: FBP over 1000 x 128^2 x 180) are measured in the same run and reported
under `extra`.

N = 1: the whole batch on one GPU.
N > 1 (torchrun, one rank per GPU), default `--shard angle`: configs[3]'s angle-sharded mode --
every rank holds all B images and A/N angles; the forward writes disjoint sinogram row
blocks and the partial back-projections are summed over the ranks INSIDE the timed step
(strong scaling: the total work is fixed).  The warm-up asserts that the reduced
back-projection equals a single-rank adjoint of the same cotangent (rel-L2 <= 1e-6).
`--shard batch`: every rank owns its own B images (weak scaling, no data-path collective).

`--impl reference` times the reference's CPU dataflow (oracle port; TensorFlow itself
is not installable in this image) on the host cores, on the same workload.
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
    # BASELINE.json configs[1]: the reference's projector/adjoint microbench
    "c2": dict(B=256, X=128, A=180, name="configs[1] projector/adjoint microbench: foam 128x128, 180 angles, batch 256"),
    # BASELINE.json configs[3]: large-scale operator sweep -- the configuration north_star's target is quoted on
    "c4": dict(B=64, X=512, A=720, name="configs[3] operator sweep: 512x512, 720 angles, batch 64"),
    # BASELINE.json configs[4]: FBP evaluation pass over the full foam dataset (helper_functions.py:477-529)
    "c5": dict(B=1000, X=128, A=180, name="configs[4] FBP evaluation pass: 1000 foam sinograms 180x184 -> 128x128, ramp filter"),
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
    """Per-launch DRAM bytes (read + write) of `kernel` from the committed `ncu --set full` capture of this
    workload (profiles/dram_traffic.json, which names the .ncu-rep summary each figure comes from), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "dram_traffic.json")) as f:
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


def measured_smem_peak(sm_mhz):
    """Shared-memory read bandwidth ceiling in GB/s: bytes per clock per SM measured by tools/micro/ldsbw.cu
    (profiles/r2_ldsbw.json, conflict-free LDS.128) x 148 SMs x the clock sampled during this run."""
    per_clk, src = 128.0, "derived (128 B/clk/SM)"
    try:
        with open(os.path.join(ROOT, "profiles", "r2_ldsbw.json")) as f:
            per_clk = float(json.load(f)["lds128_bytes_per_clk_per_sm"])
            src = "measured (tools/micro/ldsbw.cu, profiles/r2_ldsbw.json)"
    except Exception:
        pass
    return 148 * per_clk * sm_mhz * 1e6 / 1e9, src


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
def synthetic_foam_cpu(B, X, seed):
    """Unit disk with random zero-valued circular pores (stand-in for xdesign.Foam,
    scripts/create_foam_images.py:27-40; xdesign is not installable here).  torch CPU tensor [B,X,X]."""
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
    return out


def synthetic_foam_torch(B, X, device, seed):
    return synthetic_foam_cpu(B, X, seed).to(device)


# ----------------------------------------------------------------------------------------- CPU arm
def cpu_reference_pass(wl, nb, seed=0):
    """One bounded sample of the reference's CPU dataflow: project_tf_fast's
    pad -> repeat -> rotate -> row-sum graph (bilinear) and TensorFlow's gradient of it,
    restated in oracle/radon_oracle.c with OpenMP over all host cores, on `nb` synthetic foam images of the
    workload (the same generator as the GPU arm).  -> (seconds, ray-sums)."""
    from oracle import radon_oracle as orc

    X, A = wl["X"], wl["A"]
    theta = np.linspace(0, np.pi, A, endpoint=False)
    img = synthetic_foam_cpu(nb, X, seed).numpy()
    W = orc.frame_of(X, X, True)[1]
    cot = np.random.default_rng(1 + seed).random((nb, A, W), dtype=np.float32)
    t0 = time.perf_counter()
    orc.forward(img, theta, True, orc.BILINEAR, dataflow=True)
    orc.adjoint_tf(cot, theta, X, X, True, orc.BILINEAR)
    return time.perf_counter() - t0, nb * A * W


def cpu_fbp_pass(wl, nb, seed=0):
    """Bounded sample of the reference's iradon (fbp_tensorflow.py:14-75) in float64 on the host cores:
    complex FFT row filter + per-angle interpolating back-projection (oracle port).  -> (seconds, updates)."""
    from oracle import radon_oracle as orc

    X, A = wl["X"], wl["A"]
    P = orc.frame_of(X, X, True)[1]
    theta = np.linspace(0, np.pi, A, endpoint=False)
    sino = np.random.default_rng(2 + seed).random((nb, A, P))
    filt = orc.get_fourier_filter(P, "ramp")
    t0 = time.perf_counter()
    orc.iradon(sino, theta, X, X, filt)
    return time.perf_counter() - t0, nb * A * X * X


def _max_sample(wl, cores):
    """Largest image count whose per-thread rotated stack [H,W,nb] keeps the port under ~8 GB."""
    P = 2 * int(np.ceil((np.sqrt(2.0) * wl["X"] + 2) / 2))
    return int(max(1, min(wl["B"], 8e9 / (cores * P * P * 4.0))))


def cpu_baseline(wl, target_s=10.0, fbp=False):
    from oracle import radon_oracle as orc

    cores = int(orc.lib().orc_max_threads())
    cap = wl["B"] if fbp else _max_sample(wl, cores)
    run = cpu_fbp_pass if fbp else cpu_reference_pass
    nb = max(1, min(cap, 2))
    run(wl, 1)                                               # first touch: library load, page-in
    t, units = run(wl, nb)
    if t < target_s / 2:
        nb = int(max(1, min(cap, nb * target_s / max(t, 1e-3))))
        t, units = run(wl, nb)
    what = ("float64 iradon: FFT row filter + interpolating back-projection" if fbp
            else "fwd dataflow + TF-style gradient")
    return {"value": units / t / 1e9, "unit": "G pixel-angle updates/s" if fbp else UNIT, "cores": cores, "kind": "port",
            "sample": f"{nb} of {wl['B']} images, all {wl['A']} angles, {what}, {t:.2f} s"}


def run_reference(args, wl):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    from oracle import radon_oracle as orc

    # all host threads the box has (torchrun pins OMP_NUM_THREADS=1 per rank; the other ranks do no work here)
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    orc.lib().orc_set_num_threads(int(avail))
    cores = int(orc.lib().orc_max_threads())
    fbp = args.workload == "c5"
    run = cpu_fbp_pass if fbp else cpu_reference_pass
    # size the per-step sample so warmup+steps stay within a few minutes
    run(wl, 1)                                               # first touch: library load, page-in
    t_probe, _ = run(wl, 2)
    budget = 120.0 / max(1, args.steps + args.warmup)
    cap = wl["B"] if fbp else _max_sample(wl, cores)
    nb = int(max(1, min(cap, 2 * budget / max(t_probe, 1e-3))))
    for _ in range(args.warmup):
        run(wl, nb)
    tot, units = 0.0, 0
    for _ in range(args.steps):
        t, u = run(wl, nb)
        tot += t
        units += u
    value = units / tot / 1e9
    P = orc.frame_of(wl["X"], wl["X"], True)[1]
    sample = f"{nb} of {wl['B']} images per step, all {wl['A']} angles"
    unit = "G pixel-angle updates/s" if fbp else UNIT
    line = {
        "impl": "reference", "metric": "FBP Gpixel-angle updates/s" if fbp else METRIC, "value": value, "unit": unit,
        "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": tot / args.steps * 1e3, "higher_is_better": True,
        # the N = 1 point belongs to the curve the N > 1 runs of the same command line continue (angle-sharded: total work fixed)
        "scaling": "strong" if args.shard == "angle" else "weak",
        "vs_baseline": None, "dtype": "f64" if fbp else "f32",
        "data": "synthetic foam (unit disk, random circular pores), random cotangents",
        "config": {"workload": wl["name"], "interpolation": INTERP,
                   "adjoint": "n/a" if fbp else "TF gradient (the reference's autodiff)", "B": wl["B"],
                   "X": wl["X"], "Y": wl["X"], "A": wl["A"], "P": P, "sample": sample},
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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

    def draw():
        return torch.randint(0, N, (b,), generator=g), torch.randperm(A, generator=ga)[:api]

    def eager(idx, angles_i):
        idx = idx.to(dev)
        return model.train_step(meas[idx], masks[idx], enc_in[idx], pnm, theta, angles_i=angles_i, num_samples=ns)[0]

    def run(step, iters, warm):
        for _ in range(warm):
            step(*draw())
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(iters):
            loss = step(*draw())
        torch.cuda.synchronize(dev)
        return time.perf_counter() - t0, float(loss)

    dt_e, loss_e = run(eager, iters, warm)
    cfg = ("README quick-start: b=5 per GPU, 128x128, 180 angles, nsa=20, api=20, ns=2, pnm=1e4, --normal --random"
           + ("; batch-sharded, gradients averaged with one NCCL all-reduce" if world > 1 else ""))
    out = {"it_per_s": iters / dt_e, "ms_per_it": dt_e / iters * 1e3, "final_loss": loss_e, "global_batch": b * world,
           "mode": "eager (one launch per op)", "config": cfg}
    if world > 1:       # the gradient all-reduce stays outside graph capture: eager only
        return out
    try:
        # the whole iteration (networks, fused projector + likelihood, backward, clipping, Adam) as ONE CUDA graph
        graphed = vae.GraphedTrainStep(model, meas, masks, enc_in, pnm, theta, batch=b, angles_per_iter=api, num_samples=ns)
        dt_g, loss_g = run(graphed, 5 * iters, warm)
        out = {"it_per_s": 5 * iters / dt_g, "ms_per_it": dt_g / (5 * iters) * 1e3, "final_loss": loss_g, "global_batch": b * world,
               "mode": "CUDA graph (vae.GraphedTrainStep): one replay per iteration", "config": cfg,
               "eager": {"it_per_s": out["it_per_s"], "ms_per_it": out["ms_per_it"]}}
    except Exception as exc:
        out["graph_error"] = repr(exc)[:300]
    return out


# ----------------------------------------------------------------------------------------- GPU arm
class Ctx:
    """Process-wide state of one bench run (device, ranks, the L2 flush buffer)."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.rank, self.world, self.local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.json_fd = None
        if self.world > 1:
            if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", ""):
                os.environ["NCCL_DEBUG"] = "WARN"
            # NCCL writes its version banner to file descriptor 1 when the first communicator comes up: point fd 1 at
            # stderr for the whole run and keep the real stdout for the ONE JSON line
            sys.stdout.flush()
            self.json_fd = os.dup(1)
            os.dup2(2, 1)
            _pin_rank_to_cores(self.local, self.world)
            dist.init_process_group("nccl", device_id=self.dev)
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=self.dev)  # > 126 MB L2

    def sync_all(self):
        self.torch.cuda.synchronize(self.dev)
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, v: float) -> float:
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def emit(self, line):
        if self.rank != 0:
            return
        if self.json_fd is None:
            print(json.dumps(line), flush=True)
        else:
            sys.stdout.flush()
            os.write(self.json_fd, (json.dumps(line) + "\n").encode())


def _pin_rank_to_cores(local, world):
    """Give every rank its own slice of the host cores (the copy engines' staging threads and the pinned-memory
    first touch then stay local to it instead of all ranks sharing core 0's neighbourhood)."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // max(1, world))
        mine = cores[local * per:(local + 1) * per] or cores
        os.sched_setaffinity(0, mine)
    except Exception:
        pass


def timed_steps(ctx, step, steps, warm):
    """W warm-up steps, then K steps timed with per-step CUDA events on the launching stream, the L2 flushed
    (256 MiB memset) before each; barrier + synchronize on both sides; MAX over ranks.
    -> (ms_per_step, per-kernel {name: (total ms, launches)}, launches, clocks)"""
    from ct_pvae_b200 import _lib
    torch = ctx.torch
    for _ in range(warm):
        step()
    ctx.sync_all()
    sampler = ClockSampler(ctx.local) if ctx.rank == 0 else None
    _lib.profile_reset()
    _lib.profile_enable(True)
    launches0 = _lib.launch_count()
    evs = []
    ctx.sync_all()
    for _ in range(steps):
        ctx.flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        evs.append((e0, e1))
    ctx.sync_all()
    launches = _lib.launch_count() - launches0
    _lib.profile_enable(False)
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    prof = _lib.profile_read()
    _lib.profile_reset()
    clocks = sampler.stop() if sampler else None
    total_ms = ctx.max_over_ranks(total_ms)
    return total_ms / steps, prof, launches, clocks, total_ms


def best_ms(ctx, fn, iters=5):
    torch = ctx.torch
    fn()
    torch.cuda.synchronize(ctx.dev)
    ts = []
    for _ in range(iters):
        ctx.flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize(ctx.dev)
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def pcie_aggregate(ctx, mbytes=128, reps=4):
    """All ranks copy pinned host memory to / from their GPU AT THE SAME TIME: the aggregate host<->device rate of the
    box, which bounds every e2e figure at N > 1 (the ranks share the host's memory system and PCIe roots)."""
    torch = ctx.torch
    h = torch.empty(mbytes << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty_like(h, device=ctx.dev)
    out = {}
    for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
        fn()
        ctx.sync_all()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize(ctx.dev)
        dt = ctx.max_over_ranks(time.perf_counter() - t0)
        out[name + "_GBs_all_ranks"] = ctx.world * reps * (mbytes << 20) / dt / 1e9
        ctx.sync_all()
    return out


def kernel_table(prof, total_ms):
    return {k: {"ms_per_launch": v[0] / v[1], "launches": v[1], "share_of_step": v[0] / total_ms} for k, v in prof.items()}


def batch_record(ctx, wl, steps, warm, seed_rank=True):
    """Device-timed fwd + exact adjoint of one workload on this rank's own batch (batch-sharded / single GPU)."""
    import ct_pvae_b200 as cp  # noqa: F401
    from ct_pvae_b200 import _lib, ops
    torch = ctx.torch
    B, X, A = wl["B"], wl["X"], wl["A"]
    theta = np.linspace(0, np.pi, A, endpoint=False)
    plan = _lib.get_plan(theta, X, X, True, ctx.local)
    P = plan.W
    img = synthetic_foam_torch(B, X, ctx.dev, seed=ctx.rank if seed_rank else 0)
    cot = torch.rand((B, A, P), device=ctx.dev, generator=torch.Generator(device=ctx.dev).manual_seed(1 + ctx.rank))
    iid, mid = ops.INTERP[INTERP], ops.ADJOINT["exact"]

    def step():
        return ops.radon_forward(img, plan, iid), ops.radon_adjoint(cot, plan, iid, mid)

    ms, prof, launches, clocks, total_ms = timed_steps(ctx, step, steps, warm)
    units = B * A * P * ctx.world
    rec = {"ms_per_step": ms, "value": units / (ms * 1e-3) / 1e9, "unit": UNIT, "units_per_step": units, "launches": launches,
           "clocks": clocks, "kernels": kernel_table(prof, total_ms), "P": P, "plan": plan.describe(B)}
    return rec, dict(plan=plan, img=img, cot=cot, theta=theta, P=P, iid=iid, mid=mid, prof=prof, total_ms=total_ms)


def smem_rooflines(rec, wl, A_loc, clocks):
    """The binding on-chip limit: shared-memory bytes of the forward gather (16 B per in-support
    bilinear sample and image) and of the exact adjoint (3 candidate bins x 4 B per pixel-angle update and image)
    against the measured LDS bandwidth at the sampled clock."""
    B, X = wl["B"], wl["X"]
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    peak, src = measured_smem_peak(sm_mhz)
    out = {"peak_GBs": peak, "peak_source": src,
           "note": "achieved = ALGORITHMIC shared-memory bytes (forward: 16 B per in-support bilinear sample-image, adjoint: 12 B per "
                   "update-image) / kernel time.  The forward's vertical-reuse march actually loads ~13 B per sample-image on "
                   "windowed shapes, so its pipe utilisation by ncu is lower than frac (C4: 77 %, profiles/r2_ncu_fwd_c4_summary.txt); "
                   "the adjoint's loads equal the algorithmic bytes (ncu 92.6 %)."}
    fwd = rec["kernels"].get("ctr_fwd_kernel", {}).get("ms_per_launch")
    adj = rec["kernels"].get("ctr_bp_kernel<exact>", {}).get("ms_per_launch")
    if fwd:
        ach = 16.0 * B * A_loc * (X + 1) ** 2 / (fwd * 1e-3) / 1e9
        out["fwd"] = {"achieved_GBs": ach, "frac": ach / peak}
    if adj:
        ach = 12.0 * B * A_loc * X * X / (adj * 1e-3) / 1e9
        out["adj"] = {"achieved_GBs": ach, "frac": ach / peak}
    return out


def roofline_of(rec, wl, A_loc, workload_key, sharded=False):
    """HBM roofline of the dominant kernel: algorithmic bytes per launch (4*B*(X*Y + A*P), the image in
    and the sinogram out, or the mirror for the adjoint) / its live event-timed duration / the measured HBM peak."""
    peak, peak_src = measured_peaks()
    kern = rec["kernels"]
    if not kern:
        return None
    name = max(kern, key=lambda k: kern[k]["ms_per_launch"] * kern[k]["launches"])
    B, X = wl["B"], wl["X"]
    # a launch may cover only part of the batch (image groups of the overlapped sharded adjoint)
    per_step = kern[name]["launches"] / max(1, rec.get("steps", 1))
    alg = 4.0 * B * (X * X + A_loc * rec["P"]) / max(1.0, per_step)
    ms = kern[name]["ms_per_launch"]
    ach = alg / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": None if sharded else ncu_traffic(workload_key, name), "peak_source": peak_src,
            "algorithmic_bytes_per_launch": alg,
            "note": "gather/issue-bound stencil: the HBM fraction is small by construction (SURVEY 8d, DESIGN 5); the "
                    "binding on-chip figure is extra.smem_roofline"}


def e2e_batch(ctx, wl, st, steps, numpy_inputs=False):
    """The public API with HOST buffers, copies inside the timed region: project_tf_fast + backproject, both issued
    with async_op=True so the forward's copy-out overlaps the adjoint's copy-in; the step ends when both results
    are in host memory.  numpy_inputs: plain (pageable) NumPy arrays, the reference callers' case."""
    import ct_pvae_b200 as cp
    torch = ctx.torch
    X = wl["X"]
    img_h = st["img"].cpu().unsqueeze(-1)
    cot_h = st["cot"].cpu()
    if numpy_inputs:
        img_h, cot_h = img_h.numpy(), cot_h.numpy()
    else:
        img_h, cot_h = img_h.pin_memory(), cot_h.pin_memory()

    def e2e_step():
        s, hs = cp.project_tf_fast(img_h, st["theta"], pad=True, dim=2, integrate_vae=True, interpolation=INTERP, async_op=True)
        g, hg = cp.backproject(cot_h, st["theta"], X, X, pad=True, interpolation=INTERP, adjoint="exact", async_op=True)
        hs.wait()
        hg.wait()
        return s, g

    for _ in range(3):
        s_h, g_h = e2e_step()
    ctx.sync_all()
    t0 = time.perf_counter()
    for _ in range(steps):
        s_h, g_h = e2e_step()
    torch.cuda.synchronize(ctx.dev)
    dt = ctx.max_over_ranks(time.perf_counter() - t0)
    n = lambda a: int(np.prod(a.shape))  # noqa: E731
    return {"value": wl["B"] * wl["A"] * st["P"] * ctx.world * steps / dt / 1e9, "unit": UNIT,
            "h2d_bytes_per_step": (n(img_h) + n(cot_h)) * 4, "d2h_bytes_per_step": (n(s_h) + n(g_h)) * 4,
            "ms_per_step": dt / steps * 1e3, "host_buffers": "pageable NumPy" if numpy_inputs else "pinned torch"}


def c5_record(ctx, steps=5):
    """BASELINE configs[4]: the FBP evaluation pass (iradon over the whole foam dataset) on one GPU, device-timed, with
    the reference's float64 iradon (oracle port) timed beside it on the host cores."""
    import ct_pvae_b200 as cp
    from ct_pvae_b200 import _lib, ops
    torch = ctx.torch
    wl = WORKLOADS["c5"]
    B, X, A = wl["B"], wl["X"], wl["A"]
    P = cp.num_proj_pix(X, X)
    theta = np.linspace(0, np.pi, A, endpoint=False)
    fplan = _lib.get_fbp_plan(theta, P, X, X, cp.get_fourier_filter(P, "ramp"), ctx.local)
    sino = torch.rand((B, A, P), device=ctx.dev, generator=torch.Generator(device=ctx.dev).manual_seed(3))
    def timed(fused):
        ran_fused = fplan.set_fused(fused)
        ops.fbp(sino, fplan)                                   # first launch (module load, attributes) outside the profile
        torch.cuda.synchronize(ctx.dev)
        _lib.profile_reset()
        _lib.profile_enable(True)
        ms = best_ms(ctx, lambda: ops.fbp(sino, fplan), iters=steps)
        _lib.profile_enable(False)
        prof = _lib.profile_read()
        _lib.profile_reset()
        return ms, {k: {"ms_per_launch": v[0] / v[1], "launches": v[1]} for k, v in prof.items()}, ran_fused

    ms1, k1, ran_fused = timed(True)       # ONE cluster kernel (opt-in: measured slower)
    ms, k2, _ = timed(False)               # the library default: row filter + gather; leaves the plan on it
    rec = {"workload": wl["name"], "ms_per_pass": ms, "value": B * A * X * X / (ms * 1e-3) / 1e9,
           "unit": "G pixel-angle updates/s", "dtype": "f32 values, f64 geometry",
           "path": "ctr_fbp_filter_kernel + ctr_bp_kernel<fbp>", "kernels": k2,
           "single_kernel_path": ({"ms_per_pass": ms1, "kernels": k1,
                                   "what": "ctr_fbp_fused_kernel: thread-block cluster, row filter in shared memory + back-projection"}
                                  if ran_fused else None),
           "tomopy_gridrec": "unavailable (tomopy is not installable in this image)"}
    try:
        rec["cpu_baseline"] = cpu_baseline(wl, target_s=8.0, fbp=True)
    except Exception as exc:
        rec["cpu_baseline"] = {"value": None, "sample": f"failed: {exc}"}
    return rec


def run_batch(args, ctx, wl):
    """N = 1, or N > 1 with --shard batch: every rank runs the workload on its own batch (weak scaling)."""
    torch = ctx.torch
    import ct_pvae_b200 as cp
    from ct_pvae_b200 import _lib, ops
    warm = max(args.warmup, 3)
    rec, st = batch_record(ctx, wl, args.steps, warm)
    rec["steps"] = args.steps
    B, X, A, P = wl["B"], wl["X"], wl["A"], st["P"]

    side = {}
    if ctx.rank == 0 and not args.no_side_legs:
        nid = ops.INTERP["nearest"]
        plan, img, cot, iid, mid = st["plan"], st["img"], st["cot"], st["iid"], st["mid"]
        side["fwd_nearest_ms"] = best_ms(ctx, lambda: ops.radon_forward(img, plan, nid))
        side["adj_exact_nearest_ms"] = best_ms(ctx, lambda: ops.radon_adjoint(cot, plan, nid, mid))
        side["adj_tf_compat_bilinear_ms"] = best_ms(ctx, lambda: ops.radon_adjoint(cot, plan, iid, ops.ADJOINT["tf_compat"]))
        fplan = _lib.get_fbp_plan(st["theta"], P, X, X, cp.get_fourier_filter(P, "ramp"), ctx.local)
        side["fbp_ms"] = best_ms(ctx, lambda: ops.fbp(cot, fplan))
        side["fbp_gupdates_per_s"] = B * A * X * X / (side["fbp_ms"] * 1e-3) / 1e9
        side["fwd_nearest_gray_sums_per_s"] = B * A * P / (side["fwd_nearest_ms"] * 1e-3) / 1e9
    if not args.no_side_legs and not args.no_train_leg and (ctx.rank == 0 or ctx.world > 1):
        try:
            leg = vae_training_leg(ctx.dev, seed=ctx.rank, world=ctx.world)
            if ctx.rank == 0:
                side["vae_train"] = leg
        except Exception as exc:  # the restated VAE is a caller, never a reason to lose the headline
            side["vae_train"] = {"error": repr(exc)[:200]}
    ctx.sync_all()

    e2e = e2e_batch(ctx, wl, st, args.steps)
    others = {}
    if not args.no_side_legs:
        try:
            others["e2e_pageable_numpy"] = e2e_batch(ctx, wl, st, max(3, args.steps // 5), numpy_inputs=True)
        except Exception as exc:
            others["e2e_pageable_numpy"] = {"error": repr(exc)[:200]}
    del st
    torch.cuda.empty_cache()
    # the other BASELINE configurations, measured in the same run
    if not args.no_side_legs:
        for key in ("c2", "c4"):
            if key != args.workload and args.workload != "c5":
                try:
                    r2, st2 = batch_record(ctx, WORKLOADS[key], max(10, args.steps // 2), 3)
                    r2["workload"] = WORKLOADS[key]["name"]
                    r2["smem_roofline"] = smem_rooflines(r2, WORKLOADS[key], WORKLOADS[key]["A"], r2["clocks"])
                    r2["e2e"] = e2e_batch(ctx, WORKLOADS[key], st2, args.steps if key == "c2" else max(5, args.steps // 5))
                    del st2
                    torch.cuda.empty_cache()
                    others[key] = r2
                except Exception as exc:
                    others[key] = {"error": repr(exc)[:300]}
        if ctx.rank == 0:
            try:
                others["c5"] = c5_record(ctx)
            except Exception as exc:
                others["c5"] = {"error": repr(exc)[:300]}
        ctx.sync_all()

    if ctx.rank == 0:
        kern = rec["kernels"]
        fwd_ms = kern.get("ctr_fwd_kernel", {}).get("ms_per_launch")
        adj_ms = kern.get("ctr_bp_kernel<exact>", {}).get("ms_per_launch")
        extra = {"kernels": kern, "plan": rec["plan"], "side_legs": side,
                 "fwd_gray_sums_per_s": (B * A * P / (fwd_ms * 1e-3) / 1e9) if fwd_ms else None,
                 "adjoint_gupdates_per_s": (B * A * X * X / (adj_ms * 1e-3) / 1e9) if adj_ms else None,
                 "smem_roofline": smem_rooflines(rec, wl, A, rec["clocks"])}
        extra.update(others)
        line = {
            "metric": METRIC, "value": rec["value"], "unit": UNIT, "n_gpus": ctx.world, "steps": args.steps, "warmup": warm,
            "ms_per_step": rec["ms_per_step"], "higher_is_better": True,
            # N = 1 is the first point of the curve the same command line continues under torchrun: angle-sharded by default
            # (total work fixed: strong), batch-sharded with --shard batch (work per GPU fixed: weak)
            "scaling": "strong" if (ctx.world == 1 and args.shard == "angle") else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic foam (unit disk, random circular pores), random cotangents",
            "config": {"workload": wl["name"], "interpolation": INTERP, "adjoint": "exact", "B_per_gpu": B, "X": X, "Y": X,
                       "A": A, "P": P, "sharding": "batch" if ctx.world > 1 else "none (one GPU)",
                       "l2": "flushed (256 MiB memset) between timed steps"},
            "clocks": rec["clocks"],
            "e2e": {k: e2e[k] for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")},
            "gpu_launches": int(rec["launches"]),
            "roofline": roofline_of(rec, wl, A, args.workload),
            "extra": extra,
        }
        line["extra"]["e2e_detail"] = e2e
        if ctx.world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(wl)
            except Exception as exc:  # the checker library is optional for the GPU numbers
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": f"failed: {exc}"}
        ctx.emit(line)


def run_angle(args, ctx, wl):
    """N > 1, --shard angle (default): BASELINE configs[3]'s angle-sharded mode, strong scaling."""
    torch, dist = ctx.torch, ctx.dist
    from ct_pvae_b200 import _lib, ops, sharding
    B, X, A = wl["B"], wl["X"], wl["A"]
    theta = np.linspace(0, np.pi, A, endpoint=False)
    op = sharding.AngleShardedRadon(theta, X, X, True, B, ctx.dev, interpolation=INTERP, adjoint="exact", algo=args.angle_algo)
    P, A_loc = op.plan.W, op.A_local
    img = synthetic_foam_torch(B, X, ctx.dev, seed=0)                      # every rank holds all B images
    # the SAME full cotangent on every rank (seeded), of which the rank uses its angle block
    gen = torch.Generator(device=ctx.dev).manual_seed(1)
    cot_full = torch.rand((B, A, P), device=ctx.dev, generator=gen)
    cot = op.local_rows(cot_full)

    # ---- parity of the exchange step (runs on whatever hardware times it): the reduced back-projection of this
    # rank's images against a single-rank adjoint of the same cotangent over ALL angles
    full_plan = _lib.get_plan(theta, X, X, True, ctx.local)
    own = torch.as_tensor(op.owned_images(), device=ctx.dev)
    want = ops.radon_adjoint(cot_full.index_select(0, own).contiguous(), full_plan, op.iid, op.mid)
    got = op.adjoint(cot)
    parity = float((got.double() - want.double()).norm() / want.double().norm())
    worst = ctx.max_over_ranks(parity)
    # bar: 1e-6 (float32 summation order differs between N angle blocks and one 720-angle loop; reported as
    # comm_nranks_ok).  Beyond north_star's 1e-5 the run is not a measurement of the same operator: stop.
    if not (worst <= 1e-5):
        raise SystemExit(f"angle-sharded adjoint differs from the single-rank adjoint: rel-L2 {worst:.3e} > 1e-5")
    # forward blocks are disjoint rows of the single-rank sinogram: check this rank's block on a few images
    s_blk = op.forward(img[:32])
    s_ref = op.local_rows(ops.radon_forward(img[:32], full_plan, op.iid))
    fwd_par = ctx.max_over_ranks(float((s_blk.double() - s_ref.double()).norm() / s_ref.double().norm()))
    if fwd_par > 1e-6:
        raise SystemExit(f"angle-sharded forward block differs from the single-rank rows: rel-L2 {fwd_par:.3e}")
    del cot_full, want, got, s_blk, s_ref
    torch.cuda.empty_cache()

    def step():
        return op.forward(img), op.adjoint(cot)

    warm = max(args.warmup, 3)
    ms, prof, launches, clocks, total_ms = timed_steps(ctx, step, args.steps, warm)
    units = B * A * P                                                     # total work is fixed: strong scaling
    rec = {"ms_per_step": ms, "kernels": kernel_table(prof, total_ms), "P": P, "steps": args.steps}

    # ---- e2e: host buffers in, host buffers out, through the sharded operator.  The host side of the job holds the
    # batch once: every rank uploads its B/N images and the full batch is assembled over NVLink (op.gather_images), so
    # each image crosses PCIe once; the cotangent rows and both results are per-rank anyway.
    own_lo = int(op.owned_images()[0])
    img_h = img[own_lo:own_lo + B // ctx.world].cpu().pin_memory()
    cot_h = cot.cpu().pin_memory()
    sino_h = torch.empty((B, A_loc, P), dtype=torch.float32).pin_memory()
    g_h = torch.empty((B // ctx.world, X, X), dtype=torch.float32).pin_memory()
    up_stream, down_stream = torch.cuda.Stream(ctx.dev), torch.cuda.Stream(ctx.dev)

    def e2e_step():
        cur = torch.cuda.current_stream(ctx.dev)
        up_stream.wait_stream(cur)
        with torch.cuda.stream(up_stream):            # cotangent upload under the image upload / forward
            cot_d = cot_h.to(ctx.dev, non_blocking=True)
        img_d = op.gather_images(img_h.to(ctx.dev, non_blocking=True))
        s = op.forward(img_d)
        down_stream.wait_stream(cur)
        with torch.cuda.stream(down_stream):          # sinogram download under the adjoint
            sino_h.copy_(s, non_blocking=True)
        s.record_stream(down_stream)
        cur.wait_stream(up_stream)
        g = op.adjoint(cot_d)
        g_h.copy_(g, non_blocking=True)
        cot_d.record_stream(cur)
        cur.synchronize()
        down_stream.synchronize()

    for _ in range(3):
        e2e_step()
    ctx.sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize(ctx.dev)
    dt = ctx.max_over_ranks(time.perf_counter() - t0)
    h2d = (img_h.numel() + cot_h.numel()) * 4 * ctx.world
    d2h = (sino_h.numel() + g_h.numel()) * 4 * ctx.world
    e2e = {"value": units * args.steps / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "ms_per_step": dt / args.steps * 1e3,
           "note": "bytes are summed over the ranks; every rank uploads B/N images (assembled over NVLink) and its cotangent rows"}

    others = {}
    if not args.no_side_legs:
        # the other exchange algorithm, and the batch-sharded (weak-scaling) figures of C2 and C4, for context
        for alt in ([a for a in ("nccl", "torch") if a != op.algo and (a == "torch" or (op.comm is not None and op.comm.has_nccl))]):
            try:
                ms2, _, _, _, _ = timed_steps(ctx, lambda: (op.forward(img), op.adjoint(cot, algo=alt)), max(10, args.steps // 2), 3)
                others[f"angle_sharded_{alt}"] = {"ms_per_step": ms2, "value": units / (ms2 * 1e-3) / 1e9}
            except Exception as exc:
                others[f"angle_sharded_{alt}"] = {"error": repr(exc)[:200]}
        op.check()
        del img, cot
        torch.cuda.empty_cache()
        for key in ("c2", "c4"):
            try:
                r2, st2 = batch_record(ctx, WORKLOADS[key], max(10, args.steps // 2), 3)
                r2["workload"] = WORKLOADS[key]["name"] + " (batch-sharded, weak scaling)"
                del st2
                torch.cuda.empty_cache()
                others["batch_sharded_" + key] = {k: r2[k] for k in ("workload", "ms_per_step", "value", "unit")}
            except Exception as exc:
                others["batch_sharded_" + key] = {"error": repr(exc)[:200]}
        try:
            others["pcie_aggregate"] = pcie_aggregate(ctx)
        except Exception as exc:
            others["pcie_aggregate"] = {"error": repr(exc)[:200]}
        if not args.no_train_leg:
            try:
                leg = vae_training_leg(ctx.dev, seed=ctx.rank, world=ctx.world)
                others["vae_train"] = leg
            except Exception as exc:
                others["vae_train"] = {"error": repr(exc)[:200]}
        ctx.sync_all()

    if ctx.rank == 0:
        extra = {"kernels": rec["kernels"], "plan": op.plan.describe(B), "exchange": op.algo,
                 "parity": {"adjoint_vs_single_rank_rel_l2": worst, "forward_block_rel_l2": fwd_par, "bar": 1e-6,
                            "comm_nranks_ok": bool(worst <= 1e-6 and fwd_par <= 1e-6)},
                 "smem_roofline": smem_rooflines(rec, wl, A_loc, clocks)}
        extra.update(others)
        line = {
            "metric": METRIC, "value": units / (ms * 1e-3) / 1e9, "unit": UNIT, "n_gpus": ctx.world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic foam (unit disk, random circular pores), random cotangents",
            "config": {"workload": wl["name"], "interpolation": INTERP, "adjoint": "exact", "B": B, "X": X, "Y": X, "A": A, "P": P,
                       "angles_per_gpu": A_loc, "angle_assignment": op.assignment,
                       "sharding": f"angle ({op.algo} reduce-scatter of the partial back-projections inside the step)",
                       "l2": "flushed (256 MiB memset) between timed steps"},
            "clocks": clocks,
            "e2e": {k: e2e[k] for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")},
            "gpu_launches": int(launches),
            "roofline": roofline_of(rec, wl, A_loc, args.workload, sharded=True),
            "extra": extra,
        }
        line["extra"]["e2e_detail"] = e2e
        ctx.emit(line)


def run_c5(args, ctx, wl):
    """--workload c5 as the headline: the FBP evaluation pass (one GPU per rank, batch-sharded)."""
    rec = c5_record(ctx, steps=max(5, args.steps))
    if ctx.rank == 0:
        line = {"metric": "FBP Gpixel-angle updates/s", "value": rec["value"] * ctx.world, "unit": rec["unit"], "n_gpus": ctx.world,
                "steps": max(5, args.steps), "warmup": 3, "ms_per_step": rec["ms_per_pass"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic (uniform random sinograms)",
                "config": {"workload": wl["name"]}, "cpu_baseline": rec.get("cpu_baseline"), "extra": rec}
        ctx.emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--shard", default="angle", choices=["batch", "angle"],
                    help="N > 1: angle-sharded (configs[3], strong scaling, default) or batch-sharded (weak scaling)")
    ap.add_argument("--angle-algo", default="auto", choices=["auto", "p2p", "nccl", "torch"],
                    help="angle-sharded mode: how the partial back-projections are summed (see sharding.AngleShardedRadon)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side-legs", action="store_true")
    ap.add_argument("--no-train-leg", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
        return
    ctx = Ctx()
    try:
        if args.workload == "c5":
            run_c5(args, ctx, wl)
        elif ctx.world > 1 and args.shard == "angle":
            run_angle(args, ctx, wl)
        else:
            run_batch(args, ctx, wl)
    finally:
        if ctx.world > 1:
            ctx.dist.barrier()
            ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
