"""Batched get_trajectory: frames [C, num_steps, N, 3] streamed off the device slice by slice.

    python profiles/trajectory_stream.py [n_chains]
"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import mchammer_b200 as mch
from mchammer_b200 import synthetic

n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 512
mol, bonds, subunits = synthetic.cube_cage()
opt = mch.Optimizer(step_size=0.25, target_bond_length=1.2, num_steps=500, precision="fp32")
seeds = list(range(1000, 1000 + n_chains))
for io_dtype in (np.float32, np.float64):
    plan = opt.prepare_batch(mol, bonds, subunits, n_chains, io_dtype=io_dtype)
    plan.run(seeds=seeds)
    t0 = time.perf_counter(); plan.run(seeds=seeds); t_last = time.perf_counter() - t0
    for slice_bytes in (1 << 28, 1 << 30):
        plan.run_trajectories(seeds=seeds, slice_bytes=slice_bytes)
        t0 = time.perf_counter()
        traj = plan.run_trajectories(seeds=seeds, slice_bytes=slice_bytes)
        t = time.perf_counter() - t0
        print(f"{n_chains} chains x 500 steps x 1000 atoms, frames {np.dtype(io_dtype).name}: {traj.d2h_bytes / 1e9:.2f} GB "
              f"in {1e3 * t:.0f} ms = {traj.d2h_bytes / t / 1e9:.1f} GB/s to the host "
              f"(slices of {slice_bytes >> 20} MiB; last-step-only run of the same chains: {1e3 * t_last:.1f} ms)")
