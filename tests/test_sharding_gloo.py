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
    # host logic of the peer exchange set-up (ct_pvae_b200/comm.py): fixed-size blobs all-gathered in rank order, a
    # broadcast, and -- there is no GPU here -- a set-up failure that every rank must report TOGETHER instead of
    # leaving the others blocked in the next collective
    from ct_pvae_b200 import comm
    blobs = comm.exchange_blobs(bytes([65 + rank]) * 128)
    res["blobs_ok"] = np.array(blobs == b"".join(bytes([65 + r]) * 128 for r in range(world)))
    res["bcast_ok"] = np.array(comm.broadcast_blob(b"x" * 128 if rank == 0 else b"", 128) == b"x" * 128)
    try:
        comm.PeerComm(1 << 20, torch.device("cpu"))
        res["peer_error"] = np.array("none")
    except RuntimeError as exc:
        res["peer_error"] = np.array(str(exc)[:60])
    with pytest.raises(ValueError):
        sharding.project_angle_sharded(project, torch.from_numpy(img), theta[:1])      # fewer angles than ranks: all ranks raise
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
        assert bool(z["blobs_ok"]) and bool(z["bcast_ok"])
        assert "could not be mapped on every rank" in str(z["peer_error"])


def test_cost_balanced_angle_blocks_partition_the_angle_axis():
    from ct_pvae_b200.sharding import cost_balanced_range, shard_range

    for A in (5, 24, 180, 720, 721):
        theta = np.linspace(0, np.pi, A, endpoint=False)
        for world in (1, 2, 3, 4, 8):
            if A < world:
                continue
            for kappa in (0.0, 0.2, 0.5):
                r = [cost_balanced_range(theta, k, world, kappa) for k in range(world)]
                assert r[0][0] == 0 and r[-1][1] == A
                assert all(r[k][1] == r[k + 1][0] for k in range(world - 1)) and all(hi > lo for lo, hi in r)
                if kappa == 0.0:
                    assert r == [shard_range(A, k, world) for k in range(world)]
    # oblique directions are the expensive ones: the blocks around 45 / 135 degrees get fewer angles
    r = [cost_balanced_range(np.linspace(0, np.pi, 720, endpoint=False), k, 8) for k in range(8)]
    n = [hi - lo for lo, hi in r]
    assert n[1] < n[0] and n[2] < n[3] and sum(n) == 720
