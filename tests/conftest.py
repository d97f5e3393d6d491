import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.fixture(scope="session")
def orc():
    from oracle import radon_oracle

    radon_oracle.lib()
    return radon_oracle


@pytest.fixture(scope="session")
def emu():
    """CPU emulation of the kernels' per-thread code (tests/emu/ctr_emu.cpp)."""
    d = os.path.join(ROOT, "tests", "emu")
    so, src = os.path.join(d, "libctr_emu.so"), os.path.join(d, "ctr_emu.cpp")
    deps = [src] + [os.path.join(ROOT, "ct_pvae_b200", "csrc", h) for h in ("ctr_core.h", "ctr_host.h")]
    if not os.path.exists(so) or any(os.path.getmtime(p) > os.path.getmtime(so) for p in deps):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                               "-o", so, src])
    L = ctypes.CDLL(so)
    f32p, f64p, i = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double), ctypes.c_int
    L.emu_forward.argtypes = [f32p, i, i, i, i, i, i, i, f32p, i, i, i, f32p]
    L.emu_forward_depth.argtypes = L.emu_forward.argtypes
    L.emu_forward_rec32.argtypes = L.emu_forward.argtypes
    L.emu_forward_rec32_plain.argtypes = L.emu_forward.argtypes
    L.emu_forward_wide.argtypes = [f32p, i, i, i, i, i, i, i, f32p, i, i, i, i, f32p]
    L.emu_adjoint.argtypes = [f32p, i, i, i, i, i, i, i, f32p, i, i, i, i, i, i, f32p]
    L.emu_make_transforms.argtypes = [f64p, i, i, i, f32p]
    L.emu_invert_transforms.argtypes = [f32p, i, f32p]
    return L
