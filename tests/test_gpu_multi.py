"""Multi-GPU (NCCL) test of the sharding layer on the real kernels; needs >= 2 GPUs."""
import os
import socket

import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu
X, A, B = 40, 24, 32


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from ct_pvae_b200 import num_proj_pix, sharding

    rng = np.random.default_rng(0)
    img = torch.from_numpy(rng.random((B, X, X, 1), dtype=np.float32)).cuda()
    theta = np.linspace(0, np.pi, A, endpoint=False)
    W = num_proj_pix(X, X)
    cot = torch.from_numpy(rng.random((B, A, W), dtype=np.float32)).cuda()
    res = {}
    res["batch"] = sharding.radon_forward_sharded(img, theta, mode="batch", interpolation="bilinear", gather=True).cpu().numpy()
    res["angle"] = sharding.radon_forward_sharded(img, theta, mode="angle", interpolation="bilinear", gather=True).cpu().numpy()
    lo, hi = sharding.shard_range(A, rank, world)
    res["adj"] = sharding.radon_adjoint_angle_sharded(cot[:, lo:hi].contiguous(), theta, X, X, interpolation="bilinear").cpu().numpy()
    res["adj_scatter"] = sharding.radon_adjoint_angle_sharded(cot[:, lo:hi].contiguous(), theta, X, X,
                                                              interpolation="bilinear", scatter=True).cpu().numpy()
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), **res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_nccl_sharding_matches_oracle(tmp_path, orc):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(0)
    img = rng.random((B, X, X, 1), dtype=np.float32)[..., 0]
    theta = np.linspace(0, np.pi, A, endpoint=False)
    W = orc.frame_of(X, X, True)[1]
    cot = rng.random((B, A, W), dtype=np.float32)
    full = orc.forward(img, theta, True, 1)
    grad = orc.adjoint_exact(cot, theta, X, X, True, 1)
    per = B // world
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        assert rel_l2(z["batch"][..., 0], full) <= 1e-5
        assert rel_l2(z["angle"][..., 0], full) <= 1e-5
        assert rel_l2(z["adj"], grad) <= 1e-5
        assert rel_l2(z["adj_scatter"], grad[r * per:(r + 1) * per]) <= 1e-5
