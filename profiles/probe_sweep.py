"""Ceiling of the fp32 pair sweep alone (no Monte-Carlo control), in pair terms per second."""
import sys

import torch

sys.path.insert(0, ".")
from profiles import probes

sink = torch.zeros(1, dtype=torch.float32, device="cuda")
configs = ((128, 8), (64, 8), (96, 10), (256, 4)) if "--all" in sys.argv else ((128, 8),)
for threads, per_sm in configs:
    for barrier in (0, 1):
        blocks, iters, rows, cols = 148 * per_sm, 2000, 50, 950
        best = 0.0
        for _ in range(3):
            best = max(best, probes.sweep_rate(torch, blocks, threads, iters, barrier, rows, cols, sink))
        print(f"sweep-only: threads {threads} x {per_sm} CTAs/SM barrier={barrier}: {best:.3e} pair terms/s")
