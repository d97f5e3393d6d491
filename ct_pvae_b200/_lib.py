"""ctypes binding of libctradon.so (the C ABI in include/ctradon.h).

The shared library is built in-tree by ``__graft_entry__.build()`` (or
``python -m ct_pvae_b200.build``).  There is no CPU compute path and no fallback:
if the library is missing, importing this module's ``lib()`` raises.
"""
from __future__ import annotations

import ctypes
import os
import threading
from collections import OrderedDict

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CTRADON_LIB", os.path.join(_HERE, "libctradon.so"))

INTERP_NEAREST, INTERP_BILINEAR = 0, 1
ADJOINT_EXACT, ADJOINT_TF_COMPAT = 0, 1
CTR_OK, CTR_EINVAL, CTR_ECUDA, CTR_EWORKSPACE, CTR_EUNSUPPORTED, CTR_ECOMM = 0, -1, -2, -3, -4, -5
EXCHANGE_P2P, EXCHANGE_NCCL = 0, 1
COMM_HANDLE_BYTES, NCCL_ID_BYTES = 128, 128

_c_int, _c_void_p, _c_size_t = ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t
_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_intp = ctypes.POINTER(ctypes.c_int)

# every symbol include/ctradon.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "ctr_version": (_c_int, []),
    "ctr_last_error": (ctypes.c_char_p, []),
    "ctr_launch_count": (ctypes.c_longlong, []),
    "ctr_profile_enable": (_c_int, [_c_int]),
    "ctr_profile_reset": (_c_int, []),
    "ctr_profile_read": (_c_int, [_c_int, _f64p, ctypes.POINTER(ctypes.c_longlong)]),
    "ctr_kernel_name": (ctypes.c_char_p, [_c_int]),
    "ctr_device_synchronize": (_c_int, [_c_int]),
    "ctr_num_proj_pix": (_c_int, [_c_int, _c_int]),
    "ctr_frame": (_c_int, [_c_int, _c_int, _c_int, _intp, _intp, _intp, _intp]),
    "ctr_make_transforms": (_c_int, [_f64p, _c_int, _c_int, _c_int, _f32p]),
    "ctr_invert_transforms": (_c_int, [_f32p, _c_int, _f32p]),
    "ctr_filter_to_spatial": (_c_int, [_f64p, _f64p, _c_int, _f64p]),
    "ctr_plan_create": (_c_int, [_f64p, _c_int, _c_int, _c_int, _c_int, _c_int, ctypes.POINTER(_c_void_p)]),
    "ctr_plan_destroy": (_c_int, [_c_void_p]),
    "ctr_plan_info": (_c_int, [_c_void_p, _intp, _intp, _intp, _intp, _intp, _intp, _intp]),
    "ctr_plan_tables": (_c_int, [_c_void_p, _f32p, _f32p]),
    "ctr_plan_describe": (_c_int, [_c_void_p, _c_int, ctypes.c_char_p, _c_size_t]),
    "ctr_forward_workspace_bytes": (_c_size_t, [_c_void_p, _c_int]),
    "ctr_adjoint_workspace_bytes": (_c_size_t, [_c_void_p, _c_int]),
    "ctr_radon_forward": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p, _c_size_t, _c_void_p]),
    "ctr_radon_adjoint": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_size_t, _c_void_p]),
    "ctr_radon_adjoint_scaled": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, ctypes.c_float, _c_void_p, _c_size_t, _c_void_p]),
    "ctr_radon_forward_sel": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p, _c_int, _c_void_p, _c_size_t, _c_void_p]),
    "ctr_radon_adjoint_sel": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, ctypes.c_float, _c_void_p, _c_int,
                                       _c_void_p, _c_size_t, _c_void_p]),
    "ctr_radon_loglik_sel": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, ctypes.c_float, ctypes.c_float,
                                      _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p, _c_size_t, _c_void_p]),
    "ctr_loglik_workspace_bytes": (_c_size_t, [_c_void_p, _c_int]),
    "ctr_radon_loglik": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, ctypes.c_float, ctypes.c_float,
                                  _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p, _c_size_t, _c_void_p]),
    "ctr_fbp_plan_create": (_c_int, [_f64p, _c_int, _c_int, _c_int, _c_int, _f64p, _f64p, _c_int, ctypes.POINTER(_c_void_p)]),
    "ctr_fbp_plan_destroy": (_c_int, [_c_void_p]),
    "ctr_fbp_plan_set_fused": (_c_int, [_c_void_p, _c_int]),
    "ctr_fbp_workspace_bytes": (_c_size_t, [_c_void_p, _c_int]),
    "ctr_fbp": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_size_t, _c_void_p]),
    "ctr_hostpipe_create": (_c_int, [_c_void_p, _c_int, ctypes.POINTER(_c_void_p)]),
    "ctr_hostpipe_destroy": (_c_int, [_c_void_p]),
    "ctr_hostpipe_forward": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int]),
    "ctr_hostpipe_adjoint": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int]),
    "ctr_hostpipe_wait": (_c_int, [_c_void_p]),
    "ctr_hostpipe_done": (_c_int, [_c_void_p]),
    "ctr_hostpipe_trace": (_c_int, [_c_void_p, _c_int]),
    "ctr_comm_create": (_c_int, [_c_int, _c_int, _c_int, _c_size_t, ctypes.POINTER(_c_void_p)]),
    "ctr_comm_export": (_c_int, [_c_void_p, _c_void_p]),
    "ctr_comm_connect": (_c_int, [_c_void_p, _c_void_p]),
    "ctr_comm_create_all": (_c_int, [_c_int, _intp, _c_size_t, ctypes.POINTER(_c_void_p)]),
    "ctr_comm_nccl_unique_id": (_c_int, [_c_void_p]),
    "ctr_comm_nccl_init": (_c_int, [_c_void_p, _c_void_p]),
    "ctr_comm_set_timeout_ms": (_c_int, [_c_void_p, _c_int]),
    "ctr_comm_info": (_c_int, [_c_void_p, _intp, _intp, _intp, _intp, ctypes.POINTER(_c_size_t)]),
    "ctr_comm_check": (_c_int, [_c_void_p]),
    "ctr_comm_destroy": (_c_int, [_c_void_p]),
    "ctr_adjoint_sharded_workspace_bytes": (_c_size_t, [_c_void_p, _c_void_p, _c_int]),
    "ctr_radon_adjoint_sharded": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p,
                                           _c_size_t, _c_void_p]),
    "ctr_fbp_sharded": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p, _c_int, _c_int, _c_void_p, _c_size_t,
                                 _c_void_p]),
    "ctr_radon_forward_dl": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_void_p]),
    "ctr_radon_adjoint_dl": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p, _c_void_p]),
    "ctr_fbp_dl": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p]),
}

_lib = None
_lock = threading.Lock()


class CtrError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libctradon error {code}: {msg}")
        self.code = code


def lib():
    """The loaded library; raises if it has not been built (no fallback exists)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} not found: build the CUDA extension first "
                        "(python -c 'import __graft_entry__ as g; g.build()'). "
                        "ct_pvae_b200 has no CPU or PyTorch fallback."
                    )
                L = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in SYMBOLS.items():
                    fn = getattr(L, name)
                    fn.restype, fn.argtypes = res, args
                _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != CTR_OK:
        msg = lib().ctr_last_error().decode("utf-8", "replace")
        if rc == CTR_EINVAL:
            raise ValueError(msg)
        raise CtrError(rc, msg)


def launch_count() -> int:
    return int(lib().ctr_launch_count())


N_KERNELS = 9


def profile_enable(on: bool) -> None:
    check(lib().ctr_profile_enable(int(bool(on))))


def profile_reset() -> None:
    check(lib().ctr_profile_reset())


def profile_read() -> dict:
    """{kernel name: (total device ms, launches)} since the last reset (synchronises)."""
    out = {}
    for k in range(N_KERNELS):
        ms, n = ctypes.c_double(), ctypes.c_longlong()
        check(lib().ctr_profile_read(k, ctypes.byref(ms), ctypes.byref(n)))
        if n.value:
            out[lib().ctr_kernel_name(k).decode()] = (ms.value, n.value)
    return out


# --------------------------------------------------------------------------- DLPack (zero copy)
_PyCapsule_GetPointer = ctypes.pythonapi.PyCapsule_GetPointer
_PyCapsule_GetPointer.restype = ctypes.c_void_p
_PyCapsule_GetPointer.argtypes = [ctypes.py_object, ctypes.c_char_p]


class DLView:
    """Borrowed DLTensor* of a torch tensor.  The capsule stays unconsumed, so its
    destructor releases the export when this object dies; the library never calls the
    deleter.  A DLManagedTensor starts with its DLTensor, so the pointers coincide."""

    __slots__ = ("capsule", "ptr")

    def __init__(self, tensor):
        from torch.utils.dlpack import to_dlpack

        self.capsule = to_dlpack(tensor)
        self.ptr = _PyCapsule_GetPointer(self.capsule, b"dltensor")


# --------------------------------------------------------------------------- plans
PLANS_CREATED = 0     # how many ctr_plan_create calls this process has made (tools/prof_train.py, tests)


class Plan:
    """Geometry of one project_tf_fast call, resident on one device (ctr_plan)."""

    def __init__(self, theta64: np.ndarray, X: int, Y: int, pad: bool, device: int):
        global PLANS_CREATED
        PLANS_CREATED += 1
        self.handle = _c_void_p()
        th = np.ascontiguousarray(theta64, np.float64)
        check(lib().ctr_plan_create(th.ctypes.data_as(_f64p), th.size, X, Y, int(bool(pad)), device, ctypes.byref(self.handle)))
        vals = [_c_int() for _ in range(7)]
        check(lib().ctr_plan_info(self.handle, *[ctypes.byref(v) for v in vals]))
        self.A, self.X, self.Y, self.H, self.W, self.padx, self.pady = (v.value for v in vals)
        self.device = device

    def tables(self):
        fwd = np.empty((self.A, 8), np.float32)
        inv = np.empty((self.A, 8), np.float32)
        check(lib().ctr_plan_tables(self.handle, fwd.ctypes.data_as(_f32p), inv.ctypes.data_as(_f32p)))
        return fwd, inv

    def describe(self, B: int) -> str:
        buf = ctypes.create_string_buffer(512)
        check(lib().ctr_plan_describe(self.handle, int(B), buf, 512))
        return buf.value.decode()

    def forward_workspace_bytes(self, B: int) -> int:
        return int(lib().ctr_forward_workspace_bytes(self.handle, B))

    def loglik_workspace_bytes(self, B: int) -> int:
        return int(lib().ctr_loglik_workspace_bytes(self.handle, B))

    def adjoint_workspace_bytes(self, B: int) -> int:
        return int(lib().ctr_adjoint_workspace_bytes(self.handle, B))

    def close(self):
        if self.handle:
            lib().ctr_plan_destroy(self.handle)
            self.handle = _c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FbpPlan:
    """Geometry + filter of one iradon call (ctr_fbp_plan)."""

    def __init__(self, theta64, P: int, x_size: int, y_size: int, filter_1d, device: int):
        self.handle = _c_void_p()
        th = np.ascontiguousarray(theta64, np.float64)
        f = np.asarray(filter_1d).reshape(-1)
        if f.size != P:
            raise ValueError("filter_1d must have num_proj_pix entries")
        fr = np.ascontiguousarray(f.real, np.float64)
        fi = np.ascontiguousarray(f.imag, np.float64) if np.iscomplexobj(f) else None
        check(lib().ctr_fbp_plan_create(
            th.ctypes.data_as(_f64p), th.size, P, x_size, y_size, fr.ctypes.data_as(_f64p),
            fi.ctypes.data_as(_f64p) if fi is not None else None, device, ctypes.byref(self.handle)))
        self.A, self.P, self.x_size, self.y_size, self.device = th.size, P, x_size, y_size, device

    def workspace_bytes(self, B: int) -> int:
        return int(lib().ctr_fbp_workspace_bytes(self.handle, B))

    def set_fused(self, on: bool) -> bool:
        """Single-kernel (cluster, filter in shared memory) vs two-kernel path; -> whether the fused path will run."""
        rc = lib().ctr_fbp_plan_set_fused(self.handle, int(bool(on)))
        if rc < 0:
            check(rc)
        return bool(on) and rc == 0

    def close(self):
        if self.handle:
            lib().ctr_fbp_plan_destroy(self.handle)
            self.handle = _c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _PlanCache:
    """Small LRU keyed by the geometry bytes; training reuses a handful of angle sets."""

    def __init__(self, capacity: int = 64):
        self.capacity = capacity
        self.items: OrderedDict = OrderedDict()
        self.lock = threading.Lock()

    def get(self, key, factory):
        with self.lock:
            hit = self.items.get(key)
            if hit is not None:
                self.items.move_to_end(key)
                return hit
        plan = factory()
        with self.lock:
            self.items[key] = plan
            while len(self.items) > self.capacity:
                self.items.popitem(last=False)   # freed by Plan.__del__ once no call or host pipe still holds it
        return plan

    def clear(self):
        with self.lock:
            for p in self.items.values():
                p.close()
            self.items.clear()


_plans = _PlanCache()
_fbp_plans = _PlanCache(16)


def get_plan(theta64: np.ndarray, X: int, Y: int, pad: bool, device: int) -> Plan:
    th = np.ascontiguousarray(theta64, np.float64)
    key = (th.tobytes(), X, Y, bool(pad), device)
    return _plans.get(key, lambda: Plan(th, X, Y, pad, device))


def get_fbp_plan(theta64, P: int, x_size: int, y_size: int, filter_1d, device: int) -> FbpPlan:
    th = np.ascontiguousarray(theta64, np.float64)
    f = np.ascontiguousarray(np.asarray(filter_1d).reshape(-1))
    key = (th.tobytes(), P, x_size, y_size, f.tobytes(), str(f.dtype), device)
    return _fbp_plans.get(key, lambda: FbpPlan(th, P, x_size, y_size, f, device))


def clear_plan_caches() -> None:
    from . import hostpipe
    hostpipe.clear_pipes()     # pipes borrow their plan: destroy them first
    _plans.clear()
    _fbp_plans.clear()
