import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import mchammer_b200 as mch
from mchammer_b200 import synthetic, rngstate
for name, (mol, lb, su), chains, steps in (("m12l24", synthetic.cuboctahedron_cage(), 1024, 100), ("cube1000", synthetic.cube_cage(), 4096, 500)):
    opt = mch.Optimizer(0.25, 1.2, steps, precision="fp32")
    words = rngstate.seed_states(list(range(1000, 1000 + chains)))
    for ns in (1, 2, 4, 8, 16):
        plan = opt.prepare_batch(mol, lb, su, chains, per_chain_positions=True, n_slices=ns)
        pos = plan.h_pos_in
        for _ in range(2): plan.run(rng_words=words, positions=pos)
        ts = []
        for _ in range(4):
            torch.cuda.synchronize(); t = time.perf_counter(); plan.run(rng_words=words, positions=pos); ts.append(time.perf_counter() - t)
        print(f"{name} n_slices {ns:2d}: {1e3*min(ts):.2f} ms  -> {chains*steps/min(ts):.3e} chain-steps/s")
