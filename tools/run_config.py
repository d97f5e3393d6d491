"""BASELINE.json configs[0] and configs[2] driven end to end on this framework (developer runner; results kept under
profiles/).  TensorFlow, xdesign and tomopy cannot be installed here, so the reference driver itself cannot run; this
reproduces its COMMAND LINES with the restated callers (ct_pvae_b200/vae.py) around the B200 projector:

  c1  README quick-start (README.md:70-81): 50 synthetic foam images 128x128 -> 180-angle sinograms,
      main_ct_vae.py -b 5 --nsa 20 --api 20 --ns 2 --pnm 1e4 -i 1000 --normal --random
  c3  README toy run (README.md:199): 1024 toy 2x2 images (create_toy_images.py), theta = {0, pi/2}, --no_pad,
      -b 4 --pnm 1e4 --nsa 1 --ik 2 --il 5 --ks 2 --nb 3 --api 2 --se 1 --ns 10 --toy_masks --normal

Every iteration is one CUDA-graph replay (vae.GraphedTrainStep).  Prints the loss curve and, at the end, the
reconstruction quality of the posterior mean against the ground truth (MSE / SSIM / PSNR as bin/final_merit.py:97-119)
next to the FBP of the same noisy sparse-angle sinograms.

  python tools/run_config.py c1 [-i ITER]      python tools/run_config.py c3 [-i ITER]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ct_pvae_b200 import datasets, vae  # noqa: E402


def posterior_mean(model, meas, masks, enc_in, samples=8):
    """final_evaluation (main_ct_vae.py:427-461): mean of the decoder's output distribution over posterior samples."""
    eps = vae.EPS32
    with torch.no_grad():
        skips = model.encode(enc_in / 300)
        acc = 0
        for _ in range(samples):
            z = []
            for sv in skips:
                loc, log_scale = sv.chunk(2, dim=1)
                z.append(loc + (vae.positive_range(log_scale) + eps) * torch.randn_like(loc))
            alpha, beta = model.decode(z)
            acc = acc + vae.TruncatedNormal(vae.positive_range(alpha), vae.positive_range(beta), 0.0, 1e10).mean()
        return (acc / samples)[:, 0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["c1", "c3"])
    ap.add_argument("-i", "--iters", type=int, default=None)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    if args.config == "c1":
        N, X, A, b, nsa, api, ns, pnm, pad = 50, 128, 180, 5, 20, 20, 2, 1e4, True
        iters = args.iters or 1000
        theta = np.linspace(0, np.pi, A, endpoint=False)
        imgs = bench.synthetic_foam_torch(N, X, dev, seed=123)
        sino = vae.create_sinogram(imgs, theta, pad=True, interpolation="bilinear")
        masks, meas = vae.create_all_masks(sino, A, pnm, num_sparse_angles=nsa, random=True)
        model = vae.CTVAE(X, X, num_filters=1).to(dev)
        cmd = "main_ct_vae.py -b 5 --pnm 1e4 -i %d --td 50 --normal --nsa 20 --ns 2 --api 20 --random" % iters
    else:
        N, X, A, b, api, ns, pnm, pad = 1024, 2, 2, 4, 2, 10, 1e4, False
        iters = args.iters or 3000
        theta = np.array([0.0, np.pi / 2])
        x0, x1 = np.array([[1, 2], [3, 4]]) / 10, np.array([[3, 4], [1, 2]]) / 10        # create_toy_images.py:36-40
        imgs = torch.from_numpy(np.tile(np.repeat(np.stack((x0, x1)), 2, axis=0), (N // 4, 1, 1)).astype(np.float32)).to(dev)
        sino = vae.create_sinogram(imgs, theta, pad=False, interpolation="nearest")
        masks, meas = vae.create_all_masks(sino, A, pnm, toy_masks=True)
        model = vae.CTVAE(X, X, num_filters=1, num_blocks=3, kernel_size=2, stride_encode=1, intermediate_layers=5,
                          intermediate_kernel=2).to(dev)
        cmd = ("main_ct_vae.py -b 4 --pnm 10000 -i %d --td 1024 --nsa 1 --ik 2 --il 5 --ks 2 --nb 3 --api 2 --se 1 --no_pad "
               "--ns 10 --normal --toy_masks" % iters)
    enc_in = vae.iradon_all(meas, masks, theta, X, X)
    step = vae.GraphedTrainStep(model, meas, masks, enc_in, pnm, theta, batch=b, angles_per_iter=api, num_samples=ns, pad=pad)
    g = torch.Generator().manual_seed(1)
    curve, t0 = [], time.perf_counter()
    window = []
    for it in range(iters):
        loss = step(torch.randint(0, N, (b,), generator=g), torch.randperm(A, generator=g)[:api])
        window.append(loss.clone())
        if (it + 1) % max(1, iters // 10) == 0:
            curve.append(round(float(torch.stack(window).mean()), 4))
            window = []
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    rec = posterior_mean(model, meas, masks, enc_in).cpu().numpy()
    truth = imgs.cpu().numpy()
    base = enc_in[:, 0].cpu().numpy()
    mse = lambda a: float(np.mean((a - truth) ** 2))  # noqa: E731
    out = {"config": args.config, "command_reproduced": cmd, "iterations": iters, "it_per_s": iters / dt,
           "loss_curve_mean_per_tenth": curve, "loss_went_down": bool(curve[-1] < curve[0]),
           "posterior_mean_vs_truth": {"mse": mse(rec)}, "initial_reconstruction_vs_truth": {"mse": mse(base)}}
    if args.config == "c1":
        m = [datasets.compare(truth[k], rec[k]) for k in range(N)]
        f = [datasets.compare(truth[k], base[k]) for k in range(N)]
        out["posterior_mean_vs_truth"].update(ssim=float(np.mean([v[1] for v in m])), psnr=float(np.mean([v[2] for v in m])))
        out["initial_reconstruction_vs_truth"].update(ssim=float(np.mean([v[1] for v in f])), psnr=float(np.mean([v[2] for v in f])),
                                                      what="FBP (ramp) of the noisy 20-angle sinogram")
    else:
        # the ambiguous-mask examples (theta = 0 only) cannot tell x0 from x1: the posterior mean sits between them
        out["posterior_mean_first_4"] = np.round(rec[:4], 3).tolist()
        out["truth_first_4"] = truth[:4].tolist()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
