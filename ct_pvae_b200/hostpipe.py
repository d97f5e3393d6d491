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
import threading

import torch

from . import _lib


def eligible(x: torch.Tensor, min_batch: int = 32) -> bool:
    """Host batches that take the chunked pipeline: contiguous float32 CPU tensors of >= min_batch items, page-locked
    (copied by DMA directly) or pageable (a NumPy array: staged through the pipe's pinned ring by the library)."""
    return (x.device.type == "cpu" and x.dtype == torch.float32 and x.dim() >= 1 and x.shape[0] >= min_batch
            and x.is_contiguous() and not x.requires_grad)


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
        if self.handle:          # a closed pipe has run dry (close() synchronises its streams): nothing left to wait for
            _lib.check(_lib.lib().ctr_hostpipe_wait(self.handle))

    def done(self) -> bool:
        if not self.handle:
            return True
        rc = _lib.lib().ctr_hostpipe_done(self.handle)
        if rc < 0:
            _lib.check(rc)
        return rc == 1

    def trace(self, on: bool = True) -> None:
        """Record every chunk's copy-in / kernel / copy-out interval; the next wait() prints them to stderr."""
        if self.handle:
            _lib.check(_lib.lib().ctr_hostpipe_trace(self.handle, int(bool(on))))

    def close(self) -> None:
        if self.handle:
            # outstanding HostResult handles may still point here: deliver their pageable results first, then the
            # destroy call synchronises all three streams before anything is freed
            _lib.lib().ctr_hostpipe_wait(self.handle)
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


_chunk_override = {"fwd": 0, "adj": 0}
_trace_all = False


def set_chunk(fwd: int = 0, adj: int = 0) -> None:
    """Force the images per chunk of forward / adjoint host calls (0 = automatic, see chunk_for).  For tests of the
    staging ring (small chunks: slot reuse, ragged last chunk) and tuning probes."""
    _chunk_override["fwd"], _chunk_override["adj"] = int(fwd), int(adj)


def set_trace(on: bool) -> None:
    """Every pipe created or used from now on records its chunks' timeline (printed by wait())."""
    global _trace_all
    _trace_all = bool(on)
    with _pipes_lock:
        for pipe in _pipes.values():
            pipe.trace(on)


def chunk_for(plan, B: int, kind: str) -> int:
    """Images per chunk: four chunks per call, but never so little work that a chunk's launches leave most of the
    148 SMs idle (r1 measurement at 256 x 128^2 x 180: 1.39 ms per step with 64-image chunks, 2.0 ms with 32);
    whole 16-image pixel records."""
    if _chunk_override[kind] > 0:
        return _chunk_override[kind]
    work = plan.A * plan.X * plan.Y                      # pixel-angle updates per image
    by_work = -(-180_000_000 // work)                    # 64 images at 128^2 x 180, 1 at 512^2 x 720
    per = max(by_work, (B + 3) // 4)
    if B >= 64:
        # whole 32-image records: the kernels' full-efficiency shapes (32-image pixel records, 32 images per adjoint
        # thread) need more than 16 images.  r2 at 64 x 512^2 x 720: 11.4 ms per step with 32-image chunks, 13.7 with 16
        per = max(per, 32)
    return min(max(16, (per + 15) // 16 * 16), (B + 15) // 16 * 16)


def get_pipe(plan: "_lib.Plan", chunk: int, kind: str) -> NativePipe:
    # forward and adjoint calls of one plan share a pipe (r1: separate pipes measured 1.49 vs 1.39 ms per C2 step)
    key = (id(plan), chunk)
    with _pipes_lock:
        pipe = _pipes.get(key)
        if pipe is not None and pipe.plan is plan:
            _pipes.move_to_end(key)
            return pipe
        pipe = NativePipe(plan, chunk)
        if _trace_all:
            pipe.trace(True)
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


def _run(plan, x_host: torch.Tensor, out_shape, async_op: bool, issue, kind: str, out=None):
    pipe = get_pipe(plan, chunk_for(plan, int(x_host.shape[0]), kind), kind)
    if out is None:
        # page-locked result (torch's caching pinned allocator: no cudaHostAlloc after warm-up); NumPy callers get a
        # view of it
        out = torch.empty(out_shape, dtype=torch.float32, pin_memory=True)
    elif tuple(out.shape) != tuple(out_shape) or out.dtype != torch.float32 or out.device.type != "cpu" or not out.is_contiguous():
        raise ValueError(f"out must be a contiguous float32 CPU tensor of shape {tuple(out_shape)}")
    issue(pipe, out)       # on failure the library has already let its three streams run dry
    if async_op:
        return out, HostResult(pipe, (x_host, out))
    pipe.wait()
    return out


def forward_host(plan, x_host: torch.Tensor, interp: int, async_op: bool = False, out=None):
    """[B,X,Y] host float32 (pinned or pageable) -> [B,A,W] pinned float32 (project_tf_fast on a host batch)."""
    return _run(plan, x_host, (x_host.shape[0], plan.A, plan.W), async_op, lambda pipe, o: pipe.forward(x_host, o, interp), "fwd", out)


def adjoint_host(plan, y_host: torch.Tensor, interp: int, mode: int, async_op: bool = False, out=None):
    """[B,A,W] host float32 (pinned or pageable) -> [B,X,Y] pinned float32."""
    return _run(plan, y_host, (y_host.shape[0], plan.X, plan.Y), async_op, lambda pipe, o: pipe.adjoint(y_host, o, interp, mode), "adj", out)
