"""Build libctradon.so in-tree with nvcc for sm_100a (no torch involved)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libctradon.so")
SOURCES = ["ctr_capi.cu"]
DEPS = ["ctr_capi.cu", "ctr_kernels.cuh", "ctr_core.h", "ctr_host.h", os.path.join("..", "..", "include", "ctradon.h"),
        os.path.join("..", "..", "include", "ctr_dlpack.h")]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def up_to_date() -> bool:
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(os.path.join(CSRC, d)) <= t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return OUT
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
           "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-o", OUT] + SOURCES
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    # the image exports CC/CXX pointing at a gcc without libgomp specs; nvcc only needs a host g++
    env = dict(os.environ)
    subprocess.check_call(cmd, cwd=CSRC, env=env)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
