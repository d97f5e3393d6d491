"""World-size-2 gloo tests of the batch / angle sharding layer on CPU.  The operator is
injected (the oracle plays the kernels' part here), so what is tested is the partition
and collective logic of ct_pvae_b200/sharding.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import rel_l2

X, A, B = 12, 10, 4


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ct_pvae_b200 import sharding
    from oracle import radon_oracle as orc

    rng = np.random.default_rng(0)
    img = rng.random((B, X, X), dtype=np.float32)
    theta = np.linspace(0, np.pi, A, endpoint=False)
    W = orc.frame_of(X, X, True)[1]
    cot = rng.random((B, A, W), dtype=np.float32)

    def project(im, th):
        return torch.from_numpy(orc.forward(im.numpy(), th, True, 1))

    def backproject(s, th):
        return torch.from_numpy(orc.adjoint_exact(s.numpy(), th, X, X, True, 1))

    res = {}
    res["batch"] = sharding.project_batch_sharded(project, torch.from_numpy(img), theta, gather=True).numpy()
    res["batch_local"] = sharding.project_batch_sharded(project, torch.from_numpy(img), theta).numpy()
    res["angle"] = sharding.project_angle_sharded(project, torch.from_numpy(img), theta, gather=True).numpy()
    lo, hi = sharding.shard_range(A, rank, world)
    res["adj"] = sharding.backproject_angle_sharded(backproject, torch.from_numpy(cot[:, lo:hi]), theta).numpy()
    res["adj_scatter"] = sharding.backproject_angle_sharded(backproject, torch.from_numpy(cot[:, lo:hi]), theta,
                                                            scatter=True).numpy()
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), **res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharding_matches_single_process(tmp_path, orc):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(0)
    img = rng.random((B, X, X), dtype=np.float32)
    theta = np.linspace(0, np.pi, A, endpoint=False)
    W = orc.frame_of(X, X, True)[1]
    cot = rng.random((B, A, W), dtype=np.float32)
    full = orc.forward(img, theta, True, 1)
    grad = orc.adjoint_exact(cot, theta, X, X, True, 1)
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        np.testing.assert_array_equal(z["batch"], full)               # batch shards: bit-identical, no reduction
        np.testing.assert_array_equal(z["angle"], full)               # angle shards write disjoint rows
        np.testing.assert_array_equal(z["batch_local"], full[r * 2:(r + 1) * 2])
        assert rel_l2(z["adj"], grad) <= 1e-6                         # summation order changes with world size
        assert rel_l2(z["adj_scatter"], grad[r * 2:(r + 1) * 2]) <= 1e-6
