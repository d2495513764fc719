"""Ceiling of the fp32 pair sweep alone (no Monte-Carlo control), in pair terms per second."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from mchammer_b200 import native

lib = native.lib()
sink = torch.zeros(1, dtype=torch.float32, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
configs = ((128, 8), (64, 8), (96, 10), (256, 4)) if "--all" in sys.argv else ((128, 8),)
for threads, per_sm in configs:
    for barrier in (0, 1):
        blocks, iters, rows, cols = 148 * per_sm, 2000, 50, 950
        best = 0.0
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            native.check(lib.mch_probe_sweep(blocks, threads, iters, barrier, rows, cols,
                                             C.c_void_p(sink.data_ptr()), C.c_void_p(stream)), "probe")
            e1.record(); e1.synchronize()
            best = max(best, rows * cols * iters * blocks / (e0.elapsed_time(e1) * 1e-3))
        print(f"sweep-only: threads {threads} x {per_sm} CTAs/SM barrier={barrier}: {best:.3e} pair terms/s")
