"""Host-buffer calls: pinned NumPy / CPU-tensor batches through libctradon's native chunked pipeline.

A drop-in call from a CPU-side caller (``main_ct_vae.py:523-524``, ``scripts/images_to_sinograms.py:61-68`` in the
reference) moves every image to the GPU and every result back.  Run back to back those copies cost several times the
kernels, so ``ctr_hostpipe_*`` (include/ctradon.h) cuts the batch into chunks that flow through a ring of device
staging slots on three streams: copy-in of chunk k+1, kernels of chunk k, copy-out of chunk k-1 (PCIe is full duplex).
The whole pipeline is enqueued by ONE C call; torch only provides the pinned result tensor.

``async_op=True`` returns ``(result, handle)`` like a torch.distributed work handle: several calls issued back to back
share the ring, so the copy-out of one overlaps the copy-in of the next; ``handle.wait()`` makes the result readable.
"""
from __future__ import annotations

import collections
import ctypes
import os
import threading

import torch

from . import _lib


def eligible(x: torch.Tensor, min_batch: int = 32) -> bool:
    return (x.device.type == "cpu" and x.dtype == torch.float32 and x.dim() >= 1 and x.shape[0] >= min_batch
            and x.is_pinned() and x.is_contiguous() and not x.requires_grad)


class NativePipe:
    """ctr_hostpipe: device staging ring + streams for one plan and chunk size."""

    def __init__(self, plan: "_lib.Plan", chunk: int):
        self.plan, self.chunk = plan, int(chunk)      # the plan must outlive the pipe
        self.handle = ctypes.c_void_p()
        _lib.check(_lib.lib().ctr_hostpipe_create(plan.handle, self.chunk, ctypes.byref(self.handle)))

    def forward(self, x_host: torch.Tensor, out_host: torch.Tensor, interp: int) -> None:
        _lib.check(_lib.lib().ctr_hostpipe_forward(self.handle, x_host.data_ptr(), out_host.data_ptr(), int(x_host.shape[0]), interp))

    def adjoint(self, y_host: torch.Tensor, out_host: torch.Tensor, interp: int, mode: int) -> None:
        _lib.check(_lib.lib().ctr_hostpipe_adjoint(self.handle, y_host.data_ptr(), out_host.data_ptr(), int(y_host.shape[0]), interp, mode))

    def wait(self) -> None:
        _lib.check(_lib.lib().ctr_hostpipe_wait(self.handle))

    def done(self) -> bool:
        rc = _lib.lib().ctr_hostpipe_done(self.handle)
        if rc < 0:
            _lib.check(rc)
        return rc == 1

    def close(self) -> None:
        if self.handle:
            _lib.lib().ctr_hostpipe_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HostResult:
    """Completion handle of an ``async_op=True`` host-buffer call: the pinned result may be read after ``wait()``."""

    def __init__(self, pipe: NativePipe, keep):
        self._pipe, self._keep = pipe, keep          # keeps the host buffers alive while copies are in flight

    def wait(self) -> None:
        if self._pipe is not None:
            self._pipe.wait()
            self._pipe = self._keep = None

    def is_completed(self) -> bool:
        return self._pipe is None or self._pipe.done()

    def __del__(self):
        # a dropped handle must not release the pinned buffers while the copies are still in flight: the
        # pipe's streams are invisible to torch's pinned-memory allocator
        try:
            self.wait()
        except Exception:
            pass


_pipes: "collections.OrderedDict" = collections.OrderedDict()
_pipes_lock = threading.Lock()
_MAX_PIPES = 4


def chunk_for(plan, B: int, kind: str) -> int:
    """Images per chunk: four chunks per call, but never so little work that a chunk's launches leave most of the
    148 SMs idle (r1 measurement at 256 x 128^2 x 180: 1.39 ms per step with 64-image chunks, 2.0 ms with 32);
    whole 16-image pixel records."""
    env = os.environ.get("CTR_HOST_CHUNK_" + kind.upper()) or os.environ.get("CTR_HOST_CHUNK")
    if env:
        return max(1, int(env))
    work = plan.A * plan.X * plan.Y                      # pixel-angle updates per image
    by_work = -(-180_000_000 // work)                    # 64 images at 128^2 x 180, 1 at 512^2 x 720
    per = max(by_work, (B + 3) // 4)
    return min(max(16, (per + 15) // 16 * 16), (B + 15) // 16 * 16)


def get_pipe(plan: "_lib.Plan", chunk: int, kind: str) -> NativePipe:
    # forward and adjoint calls get separate pipes (own streams and staging): their kernels may then run
    # concurrently, the partial last wave of one filling SMs the other leaves idle
    key = (id(plan), chunk, kind if os.environ.get("CTR_HOST_SPLIT_PIPES") else "")
    with _pipes_lock:
        pipe = _pipes.get(key)
        if pipe is not None and pipe.plan is plan:
            _pipes.move_to_end(key)
            return pipe
        pipe = NativePipe(plan, chunk)
        _pipes[key] = pipe
        while len(_pipes) > _MAX_PIPES:
            _, old = _pipes.popitem(last=False)
            old.wait()
            old.close()
        return pipe


def clear_pipes() -> None:
    with _pipes_lock:
        for pipe in _pipes.values():
            pipe.close()
        _pipes.clear()


def _run(plan, x_host: torch.Tensor, out_shape, async_op: bool, issue, kind: str):
    pipe = get_pipe(plan, chunk_for(plan, int(x_host.shape[0]), kind), kind)
    out = torch.empty(out_shape, dtype=torch.float32, pin_memory=True)
    try:
        issue(pipe, out)
    except Exception:
        pipe.wait()        # chunks already enqueued still reference the host buffers
        raise
    if async_op:
        return out, HostResult(pipe, (x_host, out))
    pipe.wait()
    return out


def forward_host(plan, x_host: torch.Tensor, interp: int, async_op: bool = False):
    """[B,X,Y] pinned float32 -> [B,A,W] pinned float32 (project_tf_fast on a host batch)."""
    return _run(plan, x_host, (x_host.shape[0], plan.A, plan.W), async_op, lambda pipe, out: pipe.forward(x_host, out, interp), "fwd")


def adjoint_host(plan, y_host: torch.Tensor, interp: int, mode: int, async_op: bool = False):
    """[B,A,W] pinned float32 -> [B,X,Y] pinned float32."""
    return _run(plan, y_host, (y_host.shape[0], plan.X, plan.Y), async_op, lambda pipe, out: pipe.adjoint(y_host, out, interp, mode), "adj")
