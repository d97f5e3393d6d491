import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from ct_pvae_b200 import vae, _lib
import bench
dev = torch.device("cuda", 0)
torch.manual_seed(0)
N, X, A, b, nsa, api, ns, pnm = 50, 128, 180, 5, 20, 20, 2, 1e4
theta = np.linspace(0, np.pi, A, endpoint=False)
imgs = bench.synthetic_foam_torch(N, X, dev, seed=123)
sino = vae.create_sinogram(imgs, theta, pad=True, interpolation="bilinear")
masks, meas = vae.create_all_masks(sino, A, pnm, num_sparse_angles=nsa, random=True)
enc_in = vae.iradon_all(meas, masks, theta, X, X)
model = vae.CTVAE(X, X, num_filters=1).to(dev)
g = torch.Generator().manual_seed(1); ga = torch.Generator().manual_seed(7)
def one():
    idx = torch.randint(0, N, (b,), generator=g).to(dev)
    angles_i = torch.randperm(A, generator=ga)[:api]
    loss, _, _, _ = model.train_step(meas[idx], masks[idx], enc_in[idx], pnm, theta, angles_i=angles_i, num_samples=ns)
    return loss
for _ in range(5): one()
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(20): one()
torch.cuda.synchronize(); print("ms/it", (time.perf_counter()-t0)/20*1e3)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5): one()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=60))
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=60))
