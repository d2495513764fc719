"""Issue-slot cost of the packed fp32 instructions next to MUFU (mch_probe_fma modes 1..6).

Prints warp instructions per clock per SM sub-partition (one scheduler issues at most 1.0).
"""
import sys

import torch

sys.path.insert(0, ".")
from profiles import probes

sink = torch.zeros(1, dtype=torch.float32, device="cuda")
props = torch.cuda.get_device_properties(0)
sms, clk = props.multi_processor_count, 1.965e9
modes = {1: ("8 FFMA", 8), 2: ("8 MUFU.RSQ", 8), 3: ("8 FFMA2", 8), 4: ("6 FFMA2 + 2 MUFU.RSQ", 8),
         5: ("6 FFMA + 2 MUFU.RSQ", 8), 6: ("12 FFMA + 2 MUFU.RSQ", 14),
         7: ("6 FFMA2 + 2 MUFU + LDS.128", 9)}
for threads, per_sm in ((256, 8), (128, 8), (64, 8)):
    for mode, (name, n_instr) in modes.items():
        blocks, iters = sms * per_sm, 1 << 14
        best = 1e30
        for _ in range(3):
            best = min(best, probes.probe_fma(torch, mode, blocks, threads, iters, sink) * 1e-3)
        warp_instr = n_instr * iters * blocks * threads / 32
        cycles_per_round = best * clk * sms * 4 / (iters * blocks * threads / 32)
        print(f"{threads:4d} thr x {per_sm}/SM  {name:26s} {warp_instr / best / clk / sms / 4:.3f} warp-instr/clk/SMSP"
              f"  ({cycles_per_round:.2f} SMSP-cycles per round of {n_instr})")
