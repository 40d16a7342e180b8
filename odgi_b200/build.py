"""Builds the native sm_100a library in-tree: odgi_b200/libpgsgd_b200.so (C-ABI, include/pgsgd.h)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpgsgd_b200.so")
SOURCES = ["pgsgd_kernels.cu", "pgsgd_tile2.cu", "pgsgd_lay.cu", "pgsgd_gfa.cu", "pgsgd_capi.cu"]
HEADERS = ["pgsgd_device.cuh", "pgsgd_kernels.cuh"]

NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built")
    return nvcc


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(ROOT, "include", "pgsgd.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_native(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    # experiments: PGSGD_TILE_STEPS=4096 python -m odgi_b200.build  (the default build takes the header's 2048)
    tile = ["-DPGSGD_TILE_STEPS=" + os.environ["PGSGD_TILE_STEPS"]] if os.environ.get("PGSGD_TILE_STEPS") else []
    cmd = [_nvcc()] + NVCC_FLAGS + tile + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + \
          [os.path.join(CSRC, s) for s in SOURCES] + ["-lnccl"]
    env = dict(os.environ)
    # nvcc must drive the system g++ (an alternative g++ on PATH lacks the OpenMP spec files and is untested with nvcc)
    env.pop("CXX", None)
    env.pop("CC", None)
    subprocess.run(cmd, check=True, env=env)
    return LIB


if __name__ == "__main__":
    print(build_native(force=True, verbose="-v" in __import__("sys").argv))
