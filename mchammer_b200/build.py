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
# check build (asserts + cluster race detector, mu = 3 / f64 position arrays only): used by tests/test_gpu_check_build.py
CHECK_LIB = os.path.join(PKG, "libmchammer_b200_check.so")
CHECK_DEFINES = ("MCH_QUICK", "MCH_RACECHECK")
# bench-only measurement kernels (roofline denominators, sweep-only ceiling): their own library
PROBES_DIR = os.path.join(os.path.dirname(PKG), "profiles", "probes")
PROBES_SRC = os.path.join(PROBES_DIR, "mch_probes.cu")
PROBES_LIB = os.path.join(PROBES_DIR, "libmch_probes.so")
SOURCES = ["mch_api.cu", "mch_optimize_f32.cu", "mch_optimize_f64.cu", "mch_wide_f32.cu", "mch_wide_f64.cu"]
HEADERS = [
    "mch_device.cuh", "mch_optimize.cuh", "mch_wide.cuh", "mch_potential.cuh", "mch_collapse.cuh", "mch_host.h",
    os.path.join("..", "..", "include", "mchammer_b200.h"),
]

# The step kernel exists in two instantiations that must agree bit for bit (full evaluation / early rejection,
# csrc/mch_optimize.cuh).  Implicit contraction of a * b + c into an FMA is the compiler's choice per instantiation
# (it was observed to differ: last-bit differences in the carried potentials), so these translation units are
# compiled without it; every FMA of the pair sweeps is written out (fma(), packed-instruction asm) and stays.
PER_SOURCE_FLAGS = {"mch_optimize_f32.cu": ["-fmad=false"], "mch_optimize_f64.cu": ["-fmad=false"]}

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-ftz=true",  # fp32 rsqrtf -> a single MUFU.RSQ; fp64 arithmetic is unaffected
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build libmchammer_b200.so")
    return exe


def is_stale(lib: str = LIB) -> bool:
    if not os.path.exists(lib):
        return True
    built = os.path.getmtime(lib)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > built for d in deps)


def build_library(force: bool = False, verbose: bool = False, defines=(), out: str | None = None) -> str:
    """Compile every translation unit (in parallel) and link them into the shared library.

    ``defines``/``out`` build a tuning variant (``-DNAME=VALUE`` ...) into another file; the
    default build has no defines and goes to libmchammer_b200.so.
    """
    if out is None and not defines and not force and not is_stale():
        return LIB
    target = out or LIB
    tag = "_".join(d.replace("=", "-") for d in defines)
    objdir = os.path.join(PKG, "build", tag) if tag else os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    jobs = []
    for src in SOURCES:
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        cmd = [_nvcc(), *NVCC_FLAGS, *PER_SOURCE_FLAGS.get(src, []), *[f"-D{d}" for d in defines], "-c", "-o", obj,
               os.path.join(CSRC, src)]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
        jobs.append((cmd, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objects = []
    for cmd, obj, proc in jobs:
        out, _ = proc.communicate()
        if verbose:
            sys.stderr.write(out)
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + out)
        objects.append(obj)
    link = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", target, *objects]
    proc = subprocess.run(link, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(link) + "\n" + proc.stdout + proc.stderr)
    return target


def build_check_library(force: bool = False) -> str:
    """libmchammer_b200_check.so: the same sources with -DMCH_RACECHECK (csrc/mch_device.cuh) -- device asserts
    on the step descriptors and the race detector for the thread-block-cluster protocols."""
    if not force and not is_stale(CHECK_LIB):
        return CHECK_LIB
    return build_library(force=True, defines=CHECK_DEFINES, out=CHECK_LIB)


def build_probes(force: bool = False, defines=(), out: str | None = None) -> str:
    """profiles/probes/mch_probes.cu -> libmch_probes.so (bench-only; includes the product's device headers)."""
    target = out or PROBES_LIB
    deps = [PROBES_SRC, os.path.join(CSRC, "mch_device.cuh")]
    if (out is None and not defines and not force and os.path.exists(target)
            and all(os.path.getmtime(d) <= os.path.getmtime(target) for d in deps)):
        return target
    cmd = [_nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-shared", "-o", target, PROBES_SRC]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    return target


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    path = build_library(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs,
                         out=outs[0] if outs else None)
    print(path)
    if not outs:
        print(build_probes(force="--force" in sys.argv))
        print(build_check_library(force="--force" in sys.argv))
