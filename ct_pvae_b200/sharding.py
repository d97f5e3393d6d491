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
    lo, hi = shard_range(theta.shape[0], rank, world)
    if sinogram_local.shape[1] != hi - lo:
        raise ValueError("sinogram_local does not match this rank's angle block")
    partial = backproject(sinogram_local, theta[lo:hi]) if hi > lo else None
    if partial is None:  # rank owns no angle: contributes zeros of the right shape
        raise ValueError("angle-sharded back-projection needs at least one angle per rank")
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
