"""Builds pcg_b200/lib/libpcgdo_b200.so (the C-ABI CUDA library) with nvcc for sm_100a.

In-tree on purpose: the built .so travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libpcgdo_b200.so")
SOURCES = ["do_linear.cu", "do_api.cu", "do_forest.cu", "do_affine.cu", "do_metric.cu", "do_explicit.cu", "do_peak.cu", "do_multi.cu", "do_preorder.cu", "do_tcm.cpp"]
HEADERS = ["do_device.cuh", "do_host.h", "do_ctx.h", "do_metric.cuh", "do_explicit.h", "do_pack16.cuh", os.path.join("..", "..", "include", "pcg_do_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-lpthread", "-ldl"]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    subprocess.run(cmd, cwd=CSRC, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
