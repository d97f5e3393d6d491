"""Batch / angle sharding of the Radon path over the GPUs of one box.

One process per GPU, ``torch.distributed`` (NCCL) for the plumbing.

* batch-sharded (default): images are independent units, rank ``r`` owns a contiguous
  slice of the batch.  Forward, adjoint and FBP need NO collective.
* angle-sharded (large sinograms): every rank holds all images and owns a contiguous
  block of angles.  The forward writes disjoint sinogram row-blocks (an optional
  all-gather assembles them); the adjoint / FBP produce full-size partial images that
  are summed with ONE all-reduce (or reduce-scatter) -- the path's only exchange step.

The operator is passed in as a callable so the partition / collective logic can be
exercised on CPU with the gloo backend (tests/test_sharding_gloo.py); the product
entry points below bind it to the CUDA kernels.
"""
from __future__ import annotations

from typing import Callable, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) block of ``n`` items for ``rank``; the first ``n % world``
    ranks get one extra item.  Empty blocks are legal (n < world)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _check_angles_cover_ranks(num_angles: int, world: int) -> None:
    """Angle sharding needs at least one angle per rank.  The check depends only on (A, world), so EVERY rank raises
    together -- before any local work or collective -- instead of the empty ranks failing alone while the others
    block in NCCL."""
    if num_angles < world:
        raise ValueError(f"angle sharding needs at least one angle per rank: {num_angles} angles over {world} ranks")


def cost_balanced_range(theta, rank: int, world: int, kappa: float = 0.2) -> Tuple[int, int]:
    """Contiguous angle block ``[lo, hi)`` of ``rank`` with the block boundaries placed for EQUAL COST instead of equal
    angle counts.  On wide detectors (column-windowed strips) the forward projector's cost per angle grows with the
    distance of the ray direction from the image axes: measured per 45-angle block of the 64 x 512^2 x 720 sweep,
    forward + exact adjoint cost 0.78 ms on the axes and 0.96 ms next to the diagonals
    (tools/angle_block_cost.py, profiles/r2_angle_block_cost.txt), i.e. about ``1 + kappa * o`` with ``o`` in [0, 1] the
    obliqueness and kappa = 0.2.  Equal-count blocks of pi/8 then leave an 11 % spread between the fastest and slowest
    of 8 ranks, all of it spent waiting in the exchange.  (Dealing out two blocks per rank -- one cheap, one expensive
    -- was measured too: a rank's launch then mixes two direction families and runs 8 % slower on average.)
    Pure function of (theta, world, kappa): identical on all ranks.  kappa = 0 gives ``shard_range``."""
    theta = np.asarray(theta, np.float64).reshape(-1)
    A = theta.shape[0]
    if world == 1:
        return 0, A
    if kappa <= 0 or A < 2 * world:
        return shard_range(A, rank, world)
    oblique = np.abs(np.mod(theta + np.pi / 4, np.pi / 2) - np.pi / 4) / (np.pi / 4)      # 0 on the axes .. 1 on the diagonals
    w = 1.0 + kappa * oblique
    cum = np.concatenate([[0.0], np.cumsum(w)])
    cuts = [int(np.searchsorted(cum, cum[-1] * r / world, side="left")) for r in range(world + 1)]
    cuts[0], cuts[-1] = 0, A
    for r in range(1, world + 1):                      # every rank keeps at least one angle
        cuts[r] = max(cuts[r], cuts[r - 1] + 1)
    for r in range(world - 1, 0, -1):
        cuts[r] = min(cuts[r], cuts[r + 1] - 1)
    return cuts[rank], cuts[rank + 1]


def _world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def project_batch_sharded(project: Callable, images: torch.Tensor, theta, group=None, gather: bool = False):
    """Each rank projects its slice of ``images`` [B,...]; no collective unless
    ``gather`` asks for the whole sinogram batch on every rank."""
    rank, world = _world(group)
    lo, hi = shard_range(images.shape[0], rank, world)
    local = project(images[lo:hi], theta)
    if not gather or world == 1:
        return local
    return _all_gather_cat(local, images.shape[0], 0, group)


def project_angle_sharded(project: Callable, images: torch.Tensor, theta, group=None, gather: bool = False):
    """Each rank projects ALL images at its block of angles -> [B, A_local, P(,1)];
    ``gather`` all-gathers the row-blocks into the full [B, A, P(,1)]."""
    rank, world = _world(group)
    theta = np.asarray(theta)
    _check_angles_cover_ranks(theta.shape[0], world)
    lo, hi = shard_range(theta.shape[0], rank, world)
    local = project(images, theta[lo:hi])
    if not gather or world == 1:
        return local
    return _all_gather_cat(local, theta.shape[0], 1, group)


def backproject_angle_sharded(backproject: Callable, sinogram_local: torch.Tensor, theta, group=None,
                              scatter: bool = False):
    """``sinogram_local`` holds this rank's angle block [B, A_local, P].  Every rank
    back-projects its block to a full-size partial image and the partials are summed:
    all-reduce (result replicated) or reduce-scatter over the batch (``scatter``)."""
    rank, world = _world(group)
    theta = np.asarray(theta)
    _check_angles_cover_ranks(theta.shape[0], world)
    lo, hi = shard_range(theta.shape[0], rank, world)
    if sinogram_local.shape[1] != hi - lo:
        raise ValueError("sinogram_local does not match this rank's angle block")
    partial = backproject(sinogram_local, theta[lo:hi])
    if world == 1:
        return partial
    if not scatter:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
        return partial
    B = partial.shape[0]
    if B % world != 0:
        raise ValueError("reduce-scatter needs the batch to divide evenly over the ranks")
    out = torch.empty((B // world,) + tuple(partial.shape[1:]), dtype=partial.dtype, device=partial.device)
    if partial.is_cuda:
        dist.reduce_scatter_tensor(out, partial.contiguous(), op=dist.ReduceOp.SUM, group=group)
    else:  # gloo has no reduce_scatter: same result through an all-reduce
        dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
        out.copy_(partial[rank * (B // world):(rank + 1) * (B // world)])
    return out


def _all_gather_cat(local: torch.Tensor, total: int, dim: int, group=None) -> torch.Tensor:
    rank, world = _world(group)
    sizes = [shard_range(total, r, world) for r in range(world)]
    maxn = max(hi - lo for lo, hi in sizes)
    shape = list(local.shape)
    shape[dim] = maxn
    padded = torch.zeros(shape, dtype=local.dtype, device=local.device)
    padded.narrow(dim, 0, local.shape[dim]).copy_(local)
    bufs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded.contiguous(), group=group)
    return torch.cat([b.narrow(dim, 0, hi - lo) for b, (lo, hi) in zip(bufs, sizes)], dim=dim)


# ---- product bindings (CUDA kernels) ----------------------------------------------------------
def radon_forward_sharded(images, theta, pad=True, mode="batch", interpolation="nearest", gather=False, group=None):
    """images [B,X,Y,1] on this rank's GPU -> sinogram shard (see the functions above)."""
    from .forward_functions import project_tf_fast

    fn = lambda im, th: project_tf_fast(im, th, pad=pad, dim=2, integrate_vae=True, interpolation=interpolation)  # noqa: E731
    if mode == "batch":
        return project_batch_sharded(fn, images, theta, group, gather)
    if mode == "angle":
        return project_angle_sharded(fn, images, theta, group, gather)
    raise ValueError("mode must be 'batch' or 'angle'")


def radon_adjoint_angle_sharded(sinogram_local, theta, x_size, y_size, pad=True, interpolation="nearest",
                                adjoint="exact", scatter=False, group=None):
    """sinogram_local [B,A_local,P] -> summed back-projection [B,X,Y] (or its batch shard)."""
    from .forward_functions import backproject

    fn = lambda s, th: backproject(s, th, x_size, y_size, pad=pad, interpolation=interpolation, adjoint=adjoint)  # noqa: E731
    return backproject_angle_sharded(fn, sinogram_local, theta, group, scatter)


class AngleShardedRadon:
    """The angle-sharded operator pair of SURVEY 8e on this rank's GPU (BASELINE configs[3]).

    Every rank holds all ``B`` images and the contiguous angle block ``angle_indices`` (``assignment="equal"``:
    ``shard_range(A, rank, world)``; ``"cost"``, the default: boundaries placed for equal cost, see
    ``cost_balanced_range``):

    * ``forward(img [B,X,Y])`` -> this rank's sinogram rows ``[B, A_local, W]``, row ``k`` = angle ``angle_indices[k]``
      (disjoint row sets, no exchange);
    * ``adjoint(dsino_local [B, A_local, W])`` -> the back-projection summed over ALL ranks' angle blocks, left
      batch-sharded: rank ``r`` returns images ``shard_range(B, r, world)`` as ``[B/world, X, Y]``
      (``replicate=True``: the full ``[B,X,Y]`` on every rank).  This sum is the path's only exchange step.

    ``algo`` picks how the partial back-projections are summed:
      "p2p"   the adjoint kernel's epilogue stores every tile straight into its owner's exchange buffer over
              NVLink peer memory; a flag exchange and one fixed-order sum follow (``ctr_radon_adjoint_sharded``,
              ``CTR_EXCHANGE_P2P``) -- the product path;
      "nccl"  the kernel writes the partial locally and the library calls ``ncclReduceScatter`` (``CTR_EXCHANGE_NCCL``);
      "torch" the same through ``torch.distributed.reduce_scatter_tensor`` (no ctr_comm needed);
      "auto"  "p2p" when the peer buffers could be mapped, else "torch".
    """

    def __init__(self, theta, X: int, Y: int, pad: bool, B: int, device: torch.device, interpolation: str = "bilinear",
                 adjoint: str = "exact", group=None, algo: str = "auto", assignment: str = "cost"):
        from . import _lib, ops

        if algo not in ("auto", "p2p", "nccl", "torch"):
            raise ValueError("algo must be 'auto', 'p2p', 'nccl' or 'torch'")
        if assignment not in ("cost", "equal"):
            raise ValueError("assignment must be 'cost' or 'equal'")
        self.rank, self.world = _world(group)
        self.group, self.B, self.X, self.Y = group, int(B), int(X), int(Y)
        theta = np.ascontiguousarray(np.asarray(theta, np.float64).reshape(-1))
        _check_angles_cover_ranks(theta.shape[0], self.world)
        if self.B % self.world != 0:
            raise ValueError("the batch must divide evenly over the ranks (the result is left batch-sharded)")
        self.A = int(theta.shape[0])
        # the angles this rank owns (one contiguous block), in the order of its sinogram rows.  "equal": the same number
        # of angles per rank (shard_range); "cost" (default): block boundaries placed for equal cost where the
        # forward's cost depends on the direction, i.e. on detectors wide enough for column-windowed strips
        kappa = 0.0
        if assignment == "cost" and self.world > 1:
            full = _lib.get_plan(theta, self.X, self.Y, bool(pad), device.index or 0)
            kappa = 0.2 if "windowed=1" in full.describe(self.B) else 0.0
        lo, hi = cost_balanced_range(theta, self.rank, self.world, kappa)
        self._theta, self._kappa = theta, kappa
        self.angle_indices = np.arange(lo, hi)
        self.assignment = assignment if kappa > 0 else "equal"
        self.theta_local = np.ascontiguousarray(theta[self.angle_indices])
        self.device = device
        self.plan = _lib.get_plan(self.theta_local, self.X, self.Y, bool(pad), device.index or 0)
        self.iid, self.mid = ops.INTERP[interpolation], ops.ADJOINT[adjoint]
        self.comm = None
        if algo != "torch" and self.world > 1:
            try:
                from . import comm as _comm
                self.comm = _comm.PeerComm(self.B * self.X * self.Y * 4, device, group, nccl=(algo in ("nccl", "auto")))
            except Exception as exc:  # noqa: BLE001
                if algo != "auto":
                    raise
                import warnings
                warnings.warn(f"AngleShardedRadon: peer exchange unavailable ({exc}); summing the partial back-projections "
                              "with torch.distributed.reduce_scatter_tensor instead", RuntimeWarning, stacklevel=2)
        self.algo = ("p2p" if algo == "auto" else algo) if self.comm is not None else "torch"

    @property
    def A_local(self) -> int:
        return int(self.angle_indices.shape[0])

    def local_rows(self, full: torch.Tensor) -> torch.Tensor:
        """This rank's rows of a full ``[B, A, W]`` sinogram-shaped tensor, as a contiguous ``[B, A_local, W]``."""
        idx = torch.as_tensor(self.angle_indices, device=full.device)
        return full.index_select(1, idx).contiguous()

    def gather_rows(self, local: torch.Tensor) -> torch.Tensor:
        """This rank's sinogram rows ``[B, A_local, W]`` -> the full ``[B, A, W]`` on every rank (one all-gather over
        NVLink; the blocks may differ in length, so they travel padded to the longest).  Only for consumers that need
        the whole sinogram on one device -- the operator pair itself never does."""
        if self.world == 1:
            return local
        counts = [cost_balanced_range(self._theta, r, self.world, self._kappa) for r in range(self.world)]
        longest = max(hi - lo for lo, hi in counts)
        padded = torch.zeros((local.shape[0], longest, local.shape[2]), dtype=local.dtype, device=local.device)
        padded[:, :local.shape[1]].copy_(local)
        bufs = torch.empty((self.world,) + tuple(padded.shape), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(bufs, padded, group=self.group)
        return torch.cat([bufs[r, :, :hi - lo] for r, (lo, hi) in enumerate(counts)], dim=1)

    def gather_images(self, img_shard: torch.Tensor) -> torch.Tensor:
        """``[B/world, X, Y]`` (this rank's share of a host-side batch, already on the device) -> all ``B`` images on
        every rank, over NVLink (one all-gather): each image crosses PCIe once instead of ``world`` times."""
        if self.world == 1:
            return img_shard
        full = torch.empty((self.B, self.X, self.Y), dtype=img_shard.dtype, device=img_shard.device)
        dist.all_gather_into_tensor(full, img_shard.contiguous(), group=self.group)
        return full

    def forward(self, img: torch.Tensor) -> torch.Tensor:
        from . import ops
        return ops.radon_forward(img, self.plan, self.iid)

    def adjoint(self, dsino_local: torch.Tensor, replicate: bool = False, algo: str = None) -> torch.Tensor:
        from . import ops
        if dsino_local.shape[0] != self.B or dsino_local.shape[1] != self.A_local:
            raise ValueError("dsino_local must be [B, A_local, W] for this rank's angle block")
        if self.world == 1:
            return ops.radon_adjoint(dsino_local, self.plan, self.iid, self.mid)
        algo = algo or self.algo
        if algo in ("p2p", "nccl"):
            if self.comm is None:
                raise RuntimeError("this operator was built without a ctr_comm (algo='torch')")
            return self.comm.adjoint_sharded(self.plan, dsino_local, self.iid, self.mid, replicate, algo=algo)
        partial = ops.radon_adjoint(dsino_local, self.plan, self.iid, self.mid)
        if replicate:
            dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=self.group)
            return partial
        out = torch.empty((self.B // self.world, self.X, self.Y), dtype=torch.float32, device=partial.device)
        dist.reduce_scatter_tensor(out, partial, op=dist.ReduceOp.SUM, group=self.group)
        return out

    def owned_images(self):
        """Indices (into the batch) of the images this rank's ``adjoint`` result holds, in order."""
        lo, hi = shard_range(self.B, self.rank, self.world)
        return np.arange(lo, hi)

    def check(self) -> None:
        """Raises if a peer missed an exchange (died / hung) or NCCL reported an asynchronous error."""
        if self.comm is not None:
            self.comm.check()
