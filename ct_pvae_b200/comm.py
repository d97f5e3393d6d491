"""Angle-sharded exchange: Python face of ``ctr_comm_*`` (include/ctradon.h).

One rank per GPU.  ``PeerComm`` allocates this rank's exchange buffer inside libctradon, swaps the CUDA IPC handles
with the other ranks through ``torch.distributed`` (plumbing only: a 128-byte all-gather) and maps every peer's
buffer over NVLink.  After that the angle-sharded adjoint is ONE library call whose back-projection kernel stores
its tiles straight into the owners' buffers (``ctr_radon_adjoint_sharded`` with ``CTR_EXCHANGE_P2P``); NCCL is
only an alternative (``CTR_EXCHANGE_NCCL``, ``ncclReduceScatter`` from inside the library).
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from . import _lib


def exchange_blobs(blob: bytes, group=None) -> bytes:
    """All-gather one fixed-size byte string per rank, in rank order (works on any backend: the blob travels as a
    uint8 tensor on the backend's device)."""
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if "nccl" in str(backend) else torch.device("cpu")
    mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return b"".join(bytes(t.cpu().numpy().tobytes()) for t in out)


def broadcast_blob(blob: bytes, nbytes: int, src: int = 0, group=None) -> bytes:
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if "nccl" in str(backend) else torch.device("cpu")
    t = torch.frombuffer(bytearray(blob if blob else bytes(nbytes)), dtype=torch.uint8).to(dev)
    dist.broadcast(t, src=dist.get_global_rank(group, src) if group is not None else src, group=group)
    return bytes(t.cpu().numpy().tobytes())


class PeerComm:
    """ctr_comm of this rank: exchange buffer of ``exchange_bytes`` (>= B*X*Y*4 of the largest call), peers mapped."""

    def __init__(self, exchange_bytes: int, device: torch.device, group=None, nccl: bool = True, timeout_ms: int = 0):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerComm needs an initialised torch.distributed process group (one rank per GPU)")
        self.group, self.device = group, device
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.handle = ctypes.c_void_p()
        self.has_nccl = False
        L = _lib.lib()
        flag_dev = device if device.type == "cuda" else torch.device("cpu")

        def all_ok(ok: bool) -> bool:
            """Collective success vote: a step that failed on ONE rank (say, cudaIpcOpenMemHandle refused in a
            restricted container) must make EVERY rank give up together, or the others would block in the next
            collective."""
            t = torch.tensor([1 if ok else 0], device=flag_dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
            return bool(int(t.item()))

        err = None
        try:
            _lib.check(L.ctr_comm_create(self.world, self.rank, device.index or 0, int(exchange_bytes), ctypes.byref(self.handle)))
            blob = ctypes.create_string_buffer(_lib.COMM_HANDLE_BYTES)
            _lib.check(L.ctr_comm_export(self.handle, blob))
            mine = blob.raw
        except Exception as exc:  # noqa: BLE001
            err, mine = exc, bytes(_lib.COMM_HANDLE_BYTES)
        every = exchange_blobs(mine, group)
        if err is None:
            try:
                _lib.check(L.ctr_comm_connect(self.handle, ctypes.c_char_p(every)))
            except Exception as exc:  # noqa: BLE001
                err = exc
        if not all_ok(err is None):
            self.close()
            raise RuntimeError(f"peer exchange buffers could not be mapped on every rank (this rank: {err or 'ok'})")
        if nccl:
            idb = ctypes.create_string_buffer(_lib.NCCL_ID_BYTES)
            ok = self.rank != 0 or L.ctr_comm_nccl_unique_id(idb) == _lib.CTR_OK
            ident = broadcast_blob(idb.raw if self.rank == 0 else b"", _lib.NCCL_ID_BYTES, 0, group)
            if all_ok(ok):                      # (rank 0 could not open libnccl: everybody skips NCCL)
                rc = L.ctr_comm_nccl_init(self.handle, ctypes.c_char_p(ident))
                self.has_nccl = all_ok(rc == _lib.CTR_OK)
        if timeout_ms > 0:
            _lib.check(L.ctr_comm_set_timeout_ms(self.handle, int(timeout_ms)))

    def check(self) -> None:
        """Raises CtrError(CTR_ECOMM) if a peer missed an exchange or NCCL reported an asynchronous error."""
        _lib.check(_lib.lib().ctr_comm_check(self.handle))

    def adjoint_sharded(self, plan: "_lib.Plan", dsino_local: torch.Tensor, interp: int, mode: int, replicate: bool = False,
                        algo: str = "p2p") -> torch.Tensor:
        """dsino_local [B, A_local, W] CUDA float32 -> this rank's images of the summed back-projection
        [B/world, X, Y] (``replicate``: all-gathered to [B, X, Y])."""
        from . import ops
        y = ops._require_cuda_f32(dsino_local, "dsino_local")
        B = int(y.shape[0])
        if y.dim() != 3 or y.shape[1] != plan.A or y.shape[2] != plan.W:
            raise ValueError(f"dsino_local must be [B,{plan.A},{plan.W}], got {tuple(y.shape)}")
        if B % self.world != 0:
            raise ValueError("the batch must divide evenly over the ranks")
        out = torch.empty((B // self.world, plan.X, plan.Y), dtype=torch.float32, device=y.device)
        ws = ops._workspace(plan.adjoint_workspace_bytes(B), y.device)
        aid = _lib.EXCHANGE_NCCL if algo == "nccl" else _lib.EXCHANGE_P2P
        _lib.check(_lib.lib().ctr_radon_adjoint_sharded(self.handle, plan.handle, y.data_ptr(), out.data_ptr(), B, interp, mode, aid,
                                                        ws.data_ptr(), ws.numel(), ops._stream_ptr(y.device)))
        if not replicate:
            return out
        full = torch.empty((B, plan.X, plan.Y), dtype=torch.float32, device=y.device)
        dist.all_gather_into_tensor(full, out, group=self.group)
        return full

    def fbp_sharded(self, plan: "_lib.FbpPlan", sino_local: torch.Tensor, A_total: int, algo: str = "p2p") -> torch.Tensor:
        """Angle-sharded iradon: sino_local [B, A_local, P] -> [B/world, x_size, y_size] (scaled by pi / (2 A_total))."""
        from . import ops
        y = ops._require_cuda_f32(sino_local, "sino_local")
        B = int(y.shape[0])
        out = torch.empty((B // self.world, plan.x_size, plan.y_size), dtype=torch.float32, device=y.device)
        ws = ops._workspace(plan.workspace_bytes(B), y.device)
        aid = _lib.EXCHANGE_NCCL if algo == "nccl" else _lib.EXCHANGE_P2P
        _lib.check(_lib.lib().ctr_fbp_sharded(self.handle, plan.handle, y.data_ptr(), int(y.shape[1]), int(A_total), out.data_ptr(), B,
                                              aid, ws.data_ptr(), ws.numel(), ops._stream_ptr(y.device)))
        return out

    def close(self) -> None:
        if self.handle:
            _lib.lib().ctr_comm_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
