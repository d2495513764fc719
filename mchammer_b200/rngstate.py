"""numpy PCG64 state <-> the 6-word layout the kernels consume.

The reference draws from ``np.random.default_rng(seed)`` (optimizer.py:90-93).  A chain's
stream lives on the device as ``{state_hi, state_lo, inc_hi, inc_lo, has_uint32, uinteger}``;
the kernel advances it exactly as numpy would and hands it back, so an ``Optimizer``'s
generator keeps advancing across calls like the reference's.
"""

from __future__ import annotations

import ctypes as C
from typing import Iterable

import numpy as np

from . import native

_MASK64 = (1 << 64) - 1


def words_from_bit_generator(bit_generator: np.random.PCG64) -> np.ndarray:
    state = bit_generator.state
    if state.get("bit_generator") != "PCG64":
        raise TypeError("only numpy's PCG64 bit generator is supported")
    inner = state["state"]
    return np.array(
        [
            inner["state"] >> 64, inner["state"] & _MASK64,
            inner["inc"] >> 64, inner["inc"] & _MASK64,
            int(state["has_uint32"]), int(state["uinteger"]),
        ],
        dtype=np.uint64,
    )


def words_to_bit_generator(words: np.ndarray, bit_generator: np.random.PCG64) -> None:
    w = [int(x) for x in words]
    bit_generator.state = {
        "bit_generator": "PCG64",
        "state": {"state": (w[0] << 64) | w[1], "inc": (w[2] << 64) | w[3]},
        "has_uint32": w[4],
        "uinteger": w[5],
    }


def seed_states(seeds: Iterable[int]) -> np.ndarray:
    """``[n, 6]`` uint64 state words of ``np.random.default_rng(seed)`` for each seed.

    Seeds below 2**64 go through the library's host routine ``mch_pcg64_seed`` (numpy's
    SeedSequence restated in C++, tested against numpy); larger ints use numpy itself.
    """
    seeds = [int(s) for s in seeds]
    out = np.zeros((len(seeds), 6), dtype=np.uint64)
    if not seeds:
        return out
    if any(s < 0 for s in seeds):
        raise ValueError("seeds must be non-negative")
    small = [k for k, s in enumerate(seeds) if s <= _MASK64]
    if small:
        src = np.array([seeds[k] for k in small], dtype=np.uint64)
        dst = np.zeros((len(small), 6), dtype=np.uint64)
        native.check(
            native.lib().mch_pcg64_seed(
                src.ctypes.data_as(C.c_void_p), len(small), dst.ctypes.data_as(C.c_void_p)
            ),
            "mch_pcg64_seed",
        )
        out[small] = dst
    for k, s in enumerate(seeds):
        if s > _MASK64:
            out[k] = words_from_bit_generator(np.random.PCG64(s))
    return out
