"""Build libmchammer_b200.so in-tree with nvcc for sm_100a.

    python -m mchammer_b200.build [--force]

The library is one translation unit (csrc/mch_api.cu including the kernel headers); nvcc
cross-compiles it without a GPU.  The built ``.so`` is git-ignored but travels to the GPU
box with the repository snapshot.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libmchammer_b200.so")
SOURCES = ["mch_api.cu"]
HEADERS = [
    "mch_device.cuh", "mch_optimize.cuh", "mch_potential.cuh", "mch_collapse.cuh",
    os.path.join("..", "..", "include", "mchammer_b200.h"),
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-ftz=true",  # fp32 rsqrtf -> a single MUFU.RSQ; fp64 arithmetic is unaffected
    "-shared", "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build libmchammer_b200.so")
    return exe


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    built = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > built for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES]]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    return LIB


if __name__ == "__main__":
    path = build_library(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
