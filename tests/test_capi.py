"""The C-ABI library loads here (no GPU) and exports every symbol include/ctradon.h
declares; its host-side geometry agrees with the oracle; compute calls fail loudly."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f32p = ctypes.POINTER(ctypes.c_float)
f64p = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def L():
    import __graft_entry__ as g
    from ct_pvae_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        g.build()
    return _lib.lib()


def header_functions():
    src = open(os.path.join(ROOT, "include", "ctradon.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ctr_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(L):
    from ct_pvae_b200 import _lib

    names = header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ctradon.h but not exported"
    assert sorted(_lib.SYMBOLS) == names, "ctypes table and header disagree"
    assert L.ctr_version() == 200


def test_host_geometry_matches_oracle(L, orc):
    for X, Y in [(2, 2), (128, 128), (512, 512), (33, 20), (1, 1)]:
        assert L.ctr_num_proj_pix(X, Y) == orc.num_proj_pix(X, Y)
        v = [ctypes.c_int() for _ in range(4)]
        assert L.ctr_frame(X, Y, 1, *[ctypes.byref(x) for x in v]) == 0
        assert tuple(x.value for x in v) == orc.frame_of(X, Y, True)
    th = np.linspace(-4, 4, 37)
    t = np.empty((37, 8), np.float32)
    assert L.ctr_make_transforms(th.ctypes.data_as(f64p), 37, 184, 184, t.ctypes.data_as(f32p)) == 0
    np.testing.assert_array_equal(t, orc.make_transforms(th, 184, 184))
    ti = np.empty_like(t)
    assert L.ctr_invert_transforms(t.ctypes.data_as(f32p), 37, ti.ctypes.data_as(f32p)) == 0
    np.testing.assert_array_equal(ti, orc.invert_transforms(t))


def test_filter_to_spatial_is_real_ifft(L, orc):
    for Pn in (23, 184):
        f = orc.get_fourier_filter(Pn, "ramp") if Pn % 2 == 0 else np.random.default_rng(0).random(Pn)
        fr = np.ascontiguousarray(f, np.float64)
        fi = np.ascontiguousarray(np.random.default_rng(1).random(Pn))
        out = np.empty(Pn)
        assert L.ctr_filter_to_spatial(fr.ctypes.data_as(f64p), None, Pn, out.ctypes.data_as(f64p)) == 0
        np.testing.assert_allclose(out, np.real(np.fft.ifft(fr)), atol=1e-13)
        assert L.ctr_filter_to_spatial(fr.ctypes.data_as(f64p), fi.ctypes.data_as(f64p), Pn, out.ctypes.data_as(f64p)) == 0
        np.testing.assert_allclose(out, np.real(np.fft.ifft(fr + 1j * fi)), atol=1e-13)


def test_errors_return_codes_not_aborts(L):
    assert L.ctr_num_proj_pix(0, 4) == -1 and b"positive" in L.ctr_last_error()
    h = ctypes.c_void_p()
    assert L.ctr_plan_create(None, 3, 4, 4, 1, 0, ctypes.byref(h)) == -1          # CTR_EINVAL
    assert L.ctr_plan_info(None, *([None] * 7)) == -1
    assert L.ctr_radon_forward(None, None, None, 1, 0, None, 0, None) == -1
    assert L.ctr_plan_destroy(None) == 0
    assert L.ctr_profile_read(99, None, None) == -1
    # host pipeline and diagnostics: NULL handles are errors, never crashes
    hp = ctypes.c_void_p()
    assert L.ctr_hostpipe_create(None, 64, ctypes.byref(hp)) == -1 and not hp.value
    assert L.ctr_hostpipe_forward(None, None, None, 4, 1) == -1
    assert L.ctr_hostpipe_adjoint(None, None, None, 4, 1, 0) == -1
    assert L.ctr_hostpipe_wait(None) == -1 and L.ctr_hostpipe_done(None) == -1
    assert L.ctr_hostpipe_destroy(None) == 0
    buf = ctypes.create_string_buffer(64)
    assert L.ctr_plan_describe(None, 4, buf, 64) == -1


def test_no_gpu_means_loud_failure(L):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    th = np.zeros(2)
    h = ctypes.c_void_p()
    rc = L.ctr_plan_create(th.ctypes.data_as(f64p), 2, 4, 4, 1, 0, ctypes.byref(h))
    assert rc == -2 and not h.value, "plan creation must fail with CTR_ECUDA without a device"
    assert len(L.ctr_last_error()) > 0


def _build_c_smoke(tmp_path):
    """tests/capi/capi_smoke.c: the library driven from plain C (no Python, no torch) -- compiled against
    include/ctradon.h and linked to the in-tree libctradon.so."""
    import shutil
    import subprocess

    from ct_pvae_b200 import _lib

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    _lib.lib()                                         # raises if the library has not been built
    exe = str(tmp_path / "capi_smoke")
    cuda_lib = "/usr/local/cuda/lib64"
    cmd = [shutil.which("gcc") or "/usr/bin/gcc", "-O2", "-std=c11", "-I", os.path.join(root, "include"),
           os.path.join(root, "tests", "capi", "capi_smoke.c"), "-o", exe, "-L", os.path.join(root, "ct_pvae_b200"), "-lctradon",
           "-L", cuda_lib, "-lcudart", "-lm", "-Wl,-rpath," + os.path.join(root, "ct_pvae_b200"), "-Wl,-rpath," + cuda_lib]
    subprocess.check_call(cmd)
    return exe


def test_c_program_links_and_runs_host_checks(tmp_path):
    import subprocess

    out = subprocess.run([_build_c_smoke(tmp_path)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "capi_smoke ok (host)" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_c_program_runs_the_projector_without_python(tmp_path):
    import subprocess

    out = subprocess.run([_build_c_smoke(tmp_path), "gpu"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "capi_smoke ok (gpu)" in out.stdout, out.stdout + out.stderr
