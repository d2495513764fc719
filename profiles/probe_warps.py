"""Sweep-only ceiling (no column-sum stores) against the number of resident warps per SM."""
import sys

import torch

sys.path.insert(0, ".")
from profiles import probes

sink = torch.zeros(1, dtype=torch.float32, device="cuda")
for threads in (32, 64):
    for per_sm in (8, 10, 12, 14, 16, 20, 24, 32):
        if threads * per_sm > 1024:
            continue
        blocks, iters, rows, cols = 148 * per_sm, 1000, 50, 950
        best = max(probes.sweep_rate(torch, blocks, threads, iters, 2, rows, cols, sink) for _ in range(3))
        print(f"sweep-only no-store: threads {threads} x {per_sm} CTAs/SM ({threads*per_sm//32} warps/SM): {best:.3e} pair terms/s")
