"""Multi-GPU tests of the sharding layer on the real kernels; need >= 2 GPUs (gpurun --gpus 2).

One process per GPU (torch.multiprocessing.spawn + NCCL), like bench.py under torchrun: the batch-sharded and
angle-sharded forwards, and the angle-sharded adjoint / FBP through every exchange algorithm -- "p2p" (the
back-projection kernel's epilogue stores into the owners' buffers over NVLink peer memory,
ctr_radon_adjoint_sharded), "nccl" (ncclReduceScatter called by the library) and "torch"
(torch.distributed) -- against the CPU oracle.  A second test drives two GPUs from ONE process through
ctr_comm_create_all."""
import ctypes
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
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import ct_pvae_b200 as cp
    from ct_pvae_b200 import _lib, sharding

    rng = np.random.default_rng(0)
    img = torch.from_numpy(rng.random((B, X, X, 1), dtype=np.float32)).cuda()
    theta = np.linspace(0, np.pi, A, endpoint=False)
    W = cp.num_proj_pix(X, X)
    cot = torch.from_numpy(rng.random((B, A, W), dtype=np.float32)).cuda()
    res = {}
    res["batch"] = sharding.radon_forward_sharded(img, theta, mode="batch", interpolation="bilinear", gather=True).cpu().numpy()
    res["angle"] = sharding.radon_forward_sharded(img, theta, mode="angle", interpolation="bilinear", gather=True).cpu().numpy()
    lo, hi = sharding.shard_range(A, rank, world)
    res["adj"] = sharding.radon_adjoint_angle_sharded(cot[:, lo:hi].contiguous(), theta, X, X, interpolation="bilinear").cpu().numpy()
    res["adj_scatter"] = sharding.radon_adjoint_angle_sharded(cot[:, lo:hi].contiguous(), theta, X, X,
                                                              interpolation="bilinear", scatter=True).cpu().numpy()
    # the operator object bench.py times: every exchange algorithm, several calls back to back (both parities of
    # the exchange buffer, slot reuse), for both interpolation modes
    for interp in ("bilinear", "nearest"):
        op = sharding.AngleShardedRadon(theta, X, X, True, B, dev, interpolation=interp, algo="auto")
        assert op.algo == "p2p" and op.comm is not None and op.comm.has_nccl, op.algo
        blk = op.forward(img[..., 0].contiguous())
        res[f"fwd_block_{interp}"] = blk.cpu().numpy()
        res[f"fwd_gathered_{interp}"] = op.gather_rows(blk).cpu().numpy()
        cl = op.local_rows(cot)
        res[f"angles_{interp}"] = op.angle_indices
        for algo in ("p2p", "nccl", "torch"):
            outs = [op.adjoint(cl * (k + 1), algo=algo) for k in range(3)]       # enqueued without host syncs in between
            for k, o in enumerate(outs):
                res[f"op_{algo}_{interp}_{k}"] = o.cpu().numpy()
        res[f"op_replicate_{interp}"] = op.adjoint(cl, replicate=True).cpu().numpy()
        op.check()
    # angle-sharded FBP (scale pi / (2 A_total))
    filt = cp.get_fourier_filter(W, "ramp")
    fplan = _lib.get_fbp_plan(op.theta_local, W, X, X, filt, rank)
    for algo in ("p2p", "nccl"):
        res[f"fbp_{algo}"] = op.comm.fbp_sharded(fplan, op.local_rows(cot), A, algo=algo).cpu().numpy()
    # argument errors come back as exceptions on every rank alike (nothing was enqueued)
    with pytest.raises(ValueError):
        op.comm.adjoint_sharded(op.plan, op.local_rows(cot)[:B - 1].contiguous(), op.iid, op.mid)
    with pytest.raises(ValueError):
        sharding.AngleShardedRadon(theta[:1], X, X, True, B, dev)       # fewer angles than ranks: all ranks raise together
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), **res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(900)
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
    grad = {"bilinear": orc.adjoint_exact(cot, theta, X, X, True, 1), "nearest": orc.adjoint_exact(cot, theta, X, X, True, 0)}
    fwd = {"bilinear": full, "nearest": orc.forward(img, theta, True, 0)}
    rec = orc.iradon(cot.astype(np.float64), theta, X, X, orc.get_fourier_filter(W, "ramp"))
    per = B // world
    owned = []
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        assert rel_l2(z["batch"][..., 0], full) <= 1e-5
        assert rel_l2(z["angle"][..., 0], full) <= 1e-5
        assert rel_l2(z["adj"], grad["bilinear"]) <= 1e-5
        assert rel_l2(z["adj_scatter"], grad["bilinear"][r * per:(r + 1) * per]) <= 1e-5
        for interp in ("bilinear", "nearest"):
            mine = z[f"angles_{interp}"]
            assert rel_l2(z[f"fwd_block_{interp}"], fwd[interp][:, mine]) <= 1e-5
            assert rel_l2(z[f"fwd_gathered_{interp}"], fwd[interp]) <= 1e-5
            want = grad[interp][r * per:(r + 1) * per]
            for algo in ("p2p", "nccl", "torch"):
                for k in range(3):
                    assert rel_l2(z[f"op_{algo}_{interp}_{k}"], want * (k + 1)) <= 1e-5, (algo, interp, k)
            # the fused exchange sums in rank order: identical on repeated calls
            assert np.array_equal(z[f"op_p2p_{interp}_0"] * 2, z[f"op_p2p_{interp}_1"])
            assert rel_l2(z[f"op_replicate_{interp}"], grad[interp]) <= 1e-5
        for algo in ("p2p", "nccl"):
            assert rel_l2(z[f"fbp_{algo}"], rec[r * per:(r + 1) * per]) <= 1e-5, algo
        owned.append(z["angles_bilinear"])
    assert sorted(np.concatenate(owned).tolist()) == list(range(A))      # the ranks' angle sets partition the angle axis


@pytest.mark.timeout(600)
def test_single_process_comm_create_all(orc):
    """ctr_comm_create_all: two GPUs driven by ONE process (the ncclCommInitAll analogue).  Both ranks' calls are
    enqueued before either is waited for; each rank's kernels run on its own GPU."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    from ct_pvae_b200 import _lib, ops
    from ct_pvae_b200.sharding import shard_range

    L = _lib.lib()
    world = 2
    rng = np.random.default_rng(3)
    theta = np.linspace(0, np.pi, A, endpoint=False)
    W = orc.frame_of(X, X, True)[1]
    cot = rng.random((B, A, W), dtype=np.float32)
    comms = (ctypes.c_void_p * world)()
    devs = (ctypes.c_int * world)(0, 1)
    _lib.check(L.ctr_comm_create_all(world, devs, B * X * X * 4, comms))
    try:
        outs, keep = [], []
        for r in range(world):
            lo, hi = shard_range(A, r, world)
            with torch.cuda.device(r):
                plan = _lib.get_plan(theta[lo:hi], X, X, True, r)
                y = torch.from_numpy(np.ascontiguousarray(cot[:, lo:hi])).to(f"cuda:{r}")
                out = torch.empty((B // world, X, X), dtype=torch.float32, device=f"cuda:{r}")
                ws = torch.empty(plan.adjoint_workspace_bytes(B), dtype=torch.uint8, device=f"cuda:{r}")
                _lib.check(L.ctr_radon_adjoint_sharded(comms[r], plan.handle, y.data_ptr(), out.data_ptr(), B, 1, 0, _lib.EXCHANGE_P2P,
                                                       ws.data_ptr(), ws.numel(), ops._stream_ptr(torch.device("cuda", r))))
                outs.append(out)
                keep.append((plan, y, ws))
        want = orc.adjoint_exact(cot, theta, X, X, True, 1)
        per = B // world
        for r in range(world):
            torch.cuda.synchronize(r)
            _lib.check(L.ctr_comm_check(comms[r]))
            assert rel_l2(outs[r].cpu().numpy(), want[r * per:(r + 1) * per]) <= 1e-5
    finally:
        for r in range(world):
            L.ctr_comm_destroy(comms[r])


@pytest.mark.timeout(300)
def test_missing_peer_times_out_instead_of_hanging(orc):
    """Failure detection: a 2-rank comm whose second rank never calls.  The waiting kernel gives up after the
    timeout and ctr_comm_check names the missing rank (CTR_ECOMM)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    from ct_pvae_b200 import _lib, ops

    L = _lib.lib()
    theta = np.linspace(0, np.pi, 6, endpoint=False)
    comms = (ctypes.c_void_p * 2)()
    _lib.check(L.ctr_comm_create_all(2, (ctypes.c_int * 2)(0, 1), 8 * 16 * 16 * 4, comms))
    try:
        _lib.check(L.ctr_comm_set_timeout_ms(comms[0], 200))
        plan = _lib.get_plan(theta[:3], 16, 16, True, 0)
        y = torch.rand((8, 3, plan.W), device="cuda:0")
        out = torch.empty((4, 16, 16), device="cuda:0")
        ws = torch.empty(plan.adjoint_workspace_bytes(8), dtype=torch.uint8, device="cuda:0")
        _lib.check(L.ctr_radon_adjoint_sharded(comms[0], plan.handle, y.data_ptr(), out.data_ptr(), 8, 1, 0, _lib.EXCHANGE_P2P,
                                               ws.data_ptr(), ws.numel(), ops._stream_ptr(torch.device("cuda", 0))))
        torch.cuda.synchronize(0)
        with pytest.raises(_lib.CtrError) as ei:
            _lib.check(L.ctr_comm_check(comms[0]))
        assert ei.value.code == _lib.CTR_ECOMM and "rank 1" in str(ei.value)
    finally:
        for r in range(2):
            L.ctr_comm_destroy(comms[r])
