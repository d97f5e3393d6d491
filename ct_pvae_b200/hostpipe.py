"""Chunked host<->device pipeline for calls made with HOST (pinned) buffers.

A drop-in call from a CPU-side caller moves every image to the GPU and every result
back.  Run back to back those copies cost several times the kernels, so batches are
cut into chunks and three streams overlap: copy-in of chunk k+1, kernels of chunk k,
copy-out of chunk k-1 (PCIe is full duplex).  torch provides the streams, events and
pinned allocations; the arithmetic is still only libctradon's kernels.
"""
from __future__ import annotations

from typing import Callable, Tuple

import torch

_streams = {}


def _side_streams(dev: torch.device):
    key = (dev.index or 0)
    if key not in _streams:
        _streams[key] = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))
    return _streams[key]


def eligible(x: torch.Tensor, min_batch: int = 32) -> bool:
    return (x.device.type == "cpu" and x.dtype == torch.float32 and x.dim() >= 1 and x.shape[0] >= min_batch
            and x.is_pinned() and x.is_contiguous() and not x.requires_grad)


def run_chunked(fn: Callable[[torch.Tensor], torch.Tensor], x_host: torch.Tensor, out_shape: Tuple[int, ...],
                dev: torch.device, nchunks: int = 4, align: int = 16) -> torch.Tensor:
    """y[b] = fn(x[b]) chunk by chunk over dim 0; x_host pinned float32, result pinned float32."""
    B = x_host.shape[0]
    per = max(align, ((B + nchunks - 1) // nchunks + align - 1) // align * align)
    s_in, s_out = _side_streams(dev)
    cur = torch.cuda.current_stream(dev)
    out = torch.empty(out_shape, dtype=torch.float32, pin_memory=True)
    s_in.wait_stream(cur)
    for lo in range(0, B, per):
        hi = min(B, lo + per)
        with torch.cuda.stream(s_in):
            xd = x_host[lo:hi].to(dev, non_blocking=True)
            e_in = torch.cuda.Event()
            e_in.record(s_in)
        cur.wait_event(e_in)
        xd.record_stream(cur)
        yd = fn(xd)
        e_c = torch.cuda.Event()
        e_c.record(cur)
        s_out.wait_event(e_c)
        with torch.cuda.stream(s_out):
            out[lo:hi].copy_(yd, non_blocking=True)
        yd.record_stream(s_out)
    s_out.synchronize()
    return out
