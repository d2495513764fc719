"""Kernel time vs num_steps on the bench workload: separates step-0 cost from per-step cost."""
import sys
import time

import torch

sys.path.insert(0, ".")
import mchammer_b200 as mch
from mchammer_b200 import synthetic

mol, lb, su = synthetic.cube_cage()
chains = 4096
for prec in ("fp32", "fp64"):
    out = {}
    for steps in (1, 2, 101, 501):
        opt = mch.Optimizer(0.25, 1.2, steps, precision=prec)
        plan = opt.prepare_batch(mol, lb, su, chains)
        plan.stage_inputs(seeds=list(range(chains)))
        plan.upload()
        for _ in range(2):
            plan.launch()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            plan.launch()
        e1.record(); e1.synchronize()
        out[steps] = e0.elapsed_time(e1) / 3
    per_step = (out[501] - out[101]) / 400
    print(prec, {k: round(v, 3) for k, v in out.items()}, "ms; per MC step %.4f ms; step 0 = %.1f steps" % (per_step, out[1] / per_step))
