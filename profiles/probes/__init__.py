"""Measurement kernels (bench-only; not part of the product library or its ABI).

``lib()`` loads ``profiles/probes/libmch_probes.so`` (built by ``python -m mchammer_b200.build``
from ``mch_probes.cu``); ``pipe_peak`` / ``sweep_rate`` time the probes with CUDA events.
"""

from __future__ import annotations

import ctypes as C
import os
from functools import lru_cache

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MCH_PROBES_LIB") or os.path.join(HERE, "libmch_probes.so")
SOURCE = os.path.join(HERE, "mch_probes.cu")


@lru_cache(maxsize=1)
def lib() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found; build it with `python -m mchammer_b200.build`")
    handle = C.CDLL(LIB_PATH)
    handle.mch_probe_fma.restype = C.c_int
    handle.mch_probe_fma.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    handle.mch_probe_sweep.restype = C.c_int
    handle.mch_probe_sweep.argtypes = [C.c_int] * 6 + [C.c_void_p, C.c_void_p]
    return handle


def _check(code: int, what: str) -> None:
    if code != 0:
        raise RuntimeError(f"{what} failed ({code})")


def probe_fma(torch, mode: int, blocks: int, threads: int, iters: int, sink) -> float:
    """One launch of the arithmetic probe; returns milliseconds (CUDA events)."""
    stream = torch.cuda.current_stream(sink.device).cuda_stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _check(lib().mch_probe_fma(mode, blocks, threads, iters, C.c_void_p(sink.data_ptr()), C.c_void_p(stream)),
           "mch_probe_fma")
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1)


def pipe_peak(torch, mode: int, device) -> float:
    """T(FL)OP/s of a dependent-FMA stream: mode 0 = f64 FMA, 1 = f32 FMA, 3 = packed FFMA2
    (2 lanes x 2 flops), 2 = MUFU.RSQ results per second (in 1e12/s)."""
    sink = torch.zeros(1, dtype=torch.float32, device=device)
    blocks, threads, iters = 148 * 8, 256, 1 << 15
    per_op = {2: 1.0, 3: 4.0}.get(mode, 2.0)
    best = 0.0
    for _ in range(4):
        ms = probe_fma(torch, mode, blocks, threads, iters, sink)
        best = max(best, per_op * 8.0 * iters * blocks * threads / (ms * 1e-3) / 1e12)
    return best


def sweep_rate(torch, blocks: int, threads: int, iters: int, with_barrier: int, rows: int, cols: int, sink) -> float:
    """Pair terms per second of the fp32 pair sweep alone."""
    stream = torch.cuda.current_stream(sink.device).cuda_stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _check(lib().mch_probe_sweep(blocks, threads, iters, with_barrier, rows, cols, C.c_void_p(sink.data_ptr()),
                                 C.c_void_p(stream)), "mch_probe_sweep")
    e1.record()
    e1.synchronize()
    return rows * cols * iters * blocks / (e0.elapsed_time(e1) * 1e-3)
