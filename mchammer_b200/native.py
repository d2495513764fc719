"""ctypes binding of libmchammer_b200.so (declarations: include/mchammer_b200.h).

There is NO CPU fallback: if the library is missing or the call fails, the caller gets an
exception.  Device memory and streams are PyTorch's (plumbing only); the kernels see raw
pointers.
"""

from __future__ import annotations

import ctypes as C
import os
from functools import lru_cache

# MCHAMMER_B200_LIB selects another build of the same library (kernel tuning experiments).
LIB_PATH = os.environ.get("MCHAMMER_B200_LIB") or os.path.join(
    os.path.dirname(os.path.abspath(__file__)), "libmchammer_b200.so")

MCH_OK = 0
MCH_F64, MCH_F32 = 0, 1
MCH_FLAG_EARLY_REJECTION = 1  # mch_optimize_desc.flags

EXPORTS = (
    "mch_optimize_batch", "mch_optimize_smem_bytes", "mch_potential_batch",
    "mch_collapse_batch", "mch_pcg64_seed", "mch_pcg64_stream_check",
    "mch_last_error", "mch_version",
)


class NativeLibraryError(RuntimeError):
    """libmchammer_b200.so is missing, failed to load, or returned an error code."""


class MchParams(C.Structure):
    _fields_ = [
        ("step_size", C.c_double), ("target_bond_length", C.c_double),
        ("bond_epsilon", C.c_double), ("nonbond_epsilon", C.c_double),
        ("nonbond_sigma", C.c_double), ("nonbond_mu", C.c_double), ("beta", C.c_double),
    ]


class OptimizeDesc(C.Structure):
    _fields_ = [
        ("n_chains", C.c_int32), ("n_atoms", C.c_int32), ("n_subunits", C.c_int32),
        ("n_bonds", C.c_int32), ("max_subunit_atoms", C.c_int32), ("num_steps", C.c_int32),
        ("io_dtype", C.c_int32), ("compute_dtype", C.c_int32), ("inexact_undo", C.c_int32),
        ("threads_per_chain", C.c_int32), ("cluster_size", C.c_int32), ("resident_chains_hint", C.c_int32),
        ("pos_in", C.c_void_p), ("pos_in_chain_stride", C.c_int64), ("pos_out", C.c_void_p),
        ("perm", C.c_void_p), ("su_offsets", C.c_void_p), ("bonds", C.c_void_p),
        ("bond_su", C.c_void_p),
        ("params", MchParams),
        ("rng_state", C.c_void_p),
        ("system_potential", C.c_void_p), ("nonbonded_potential", C.c_void_p),
        ("max_bond_distance", C.c_void_p), ("n_accepted", C.c_void_p),
        ("last_step_info", C.c_void_p),
        ("trace_potentials", C.c_void_p), ("trace_info", C.c_void_p),
        ("trajectory", C.c_void_p), ("pair_evals", C.c_void_p),
        ("general", C.c_int32), ("n_row_lists", C.c_int32), ("max_rows", C.c_int32), ("flags", C.c_int32),
        ("bond_slots", C.c_void_p), ("row_offsets", C.c_void_p), ("row_slots", C.c_void_p),
        ("row_weights", C.c_void_p), ("run_offsets", C.c_void_p), ("runs", C.c_void_p),
    ]


class CollapseDesc(C.Structure):
    _fields_ = [
        ("n_mol", C.c_int32), ("n_atoms", C.c_int32), ("n_subunits", C.c_int32),
        ("n_bonds", C.c_int32), ("io_dtype", C.c_int32), ("compute_dtype", C.c_int32),
        ("scale_steps", C.c_int32), ("max_steps", C.c_int32), ("trace_capacity", C.c_int32),
        ("threads_per_mol", C.c_int32), ("n_moving_subunits", C.c_int32), ("n_heavy_total", C.c_int32),
        ("step_size", C.c_double), ("distance_threshold", C.c_double),
        ("pos_in", C.c_void_p), ("pos_in_mol_stride", C.c_int64), ("pos_out", C.c_void_p),
        ("perm", C.c_void_p), ("su_offsets", C.c_void_p), ("n_heavy", C.c_void_p),
        ("bonds", C.c_void_p),
        ("steps_run", C.c_void_p), ("min_distance", C.c_void_p),
        ("max_bond_distance", C.c_void_p), ("status", C.c_void_p),
        ("trace_max_bond", C.c_void_p), ("trajectory", C.c_void_p),
    ]


@lru_cache(maxsize=1)
def lib() -> C.CDLL:
    """Load the shared library (once).  Raises :class:`NativeLibraryError` if absent."""
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} not found. Build it with `python -m mchammer_b200.build` "
            "(needs nvcc; no CPU fallback exists)."
        )
    try:
        handle = C.CDLL(LIB_PATH)
    except OSError as err:  # pragma: no cover - depends on the machine
        raise NativeLibraryError(f"cannot load {LIB_PATH}: {err}") from err

    handle.mch_last_error.restype = C.c_char_p
    handle.mch_last_error.argtypes = []
    handle.mch_version.restype = C.c_int
    handle.mch_version.argtypes = []
    handle.mch_optimize_batch.restype = C.c_int
    handle.mch_optimize_batch.argtypes = [C.POINTER(OptimizeDesc), C.c_void_p]
    handle.mch_optimize_smem_bytes.restype = C.c_int
    handle.mch_optimize_smem_bytes.argtypes = [C.c_int] * 5
    handle.mch_potential_batch.restype = C.c_int
    handle.mch_potential_batch.argtypes = [
        C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
        C.POINTER(MchParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
    ]
    handle.mch_collapse_batch.restype = C.c_int
    handle.mch_collapse_batch.argtypes = [C.POINTER(CollapseDesc), C.c_void_p]
    handle.mch_pcg64_seed.restype = C.c_int
    handle.mch_pcg64_seed.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    handle.mch_pcg64_stream_check.restype = C.c_int
    handle.mch_pcg64_stream_check.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p]
    return handle


def check(code: int, what: str) -> int:
    """Raise on a negative return code, with the library's own message."""
    if code < 0:
        message = lib().mch_last_error().decode(errors="replace")
        raise NativeLibraryError(f"{what} failed ({code}): {message}")
    return code


def require_cuda():
    """Return ``torch`` after checking that a CUDA device is usable; raise otherwise."""
    import torch

    if not torch.cuda.is_available():
        raise NativeLibraryError(
            "mchammer_b200 runs its Monte-Carlo path on a CUDA device only "
            "(torch.cuda.is_available() is False; there is no CPU fallback)."
        )
    return torch
