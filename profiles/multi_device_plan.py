"""One process, one host thread, several GPUs: BatchPlan over devices=[0..G-1] (no torchrun).

Checks that the sharded run returns exactly what one device returns and reports the
host-to-host throughput.      python profiles/multi_device_plan.py [n_devices]
"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import mchammer_b200 as mch
from mchammer_b200 import rngstate, synthetic

g = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
mol, lb, su = synthetic.cube_cage()
opt = mch.Optimizer(0.25, 1.2, 500, precision="fp32")
small = list(range(1000, 1000 + 64))
one = opt.get_results_batch(mol, lb, su, seeds=small, devices=[0])
many = opt.get_results_batch(mol, lb, su, seeds=small, devices=list(range(g)))
assert np.array_equal(one.positions, many.positions) and np.array_equal(one.system_potential, many.system_potential)
print(f"{g} devices: 64 chains sharded == 64 chains on device 0 (bitwise)")
for devices in ([0], list(range(g))):
    chains = 4096 * len(devices)
    plan = opt.prepare_batch(mol, lb, su, chains, devices=devices, per_chain_positions=True)
    words = rngstate.seed_states(list(range(1000, 1000 + chains)))
    for _ in range(2):
        plan.run(rng_words=words, positions=plan.h_pos_in)
    ts = []
    for _ in range(4):
        t = time.perf_counter(); plan.run(rng_words=words, positions=plan.h_pos_in); ts.append(time.perf_counter() - t)
    print(f"BatchPlan on devices {devices}: {chains} chains x 499 moves in {1e3 * min(ts):.1f} ms = {chains * 499 / min(ts):.3e} chain-steps/s host to host")
