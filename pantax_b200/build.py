"""Builds libpantax_gpu.so (the C-ABI library) in-tree with nvcc for sm_100a.

    python -m pantax_b200.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpantax_gpu.so")
SOURCES = ["ptx_kernels.cu", "ptx_api.cu"]
HEADERS = ["ptx_core.cuh", "ptx_fast.cuh", "ptx_internal.h", "ptx_fxorder.h", os.path.join("..", "..", "include", "pantax_gpu.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function",
    "--shared",
]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__), os.path.join(HERE, "host", "pantax_gpu_profile.cpp")]
    if not os.path.exists(os.path.join(HERE, "pantax-gpu-profile")):
        return True
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          [os.path.join(CSRC, f) for f in SOURCES] + ["-o", LIB, "-lcudart", "-ldl"]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    build_host(verbose)
    return LIB


HOST_SRC = os.path.join(HERE, "host", "pantax_gpu_profile.cpp")
HOST_BIN = os.path.join(HERE, "pantax-gpu-profile")


def build_host(verbose: bool = False) -> str:
    """The C++ host driver over the C ABI (file formats, f64 tail, TSV writers)."""
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", HOST_SRC, "-o", HOST_BIN, "-L" + HERE, "-lpantax_gpu", "-Wl,-rpath,$ORIGIN",
           "-Wl,-rpath-link," + "/usr/local/cuda/lib64", "-L/usr/local/cuda/lib64", "-lcudart", "-ldl", "-lpthread"]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return HOST_BIN


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
