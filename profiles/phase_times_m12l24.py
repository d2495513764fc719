import sys, torch
sys.path.insert(0, ".")
import mchammer_b200 as mch
from mchammer_b200 import synthetic
mol, lb, su = synthetic.cuboctahedron_cage()
chains = 296
for prec in ("fp32",):
    out = {}
    for steps in (1, 21, 101):
        opt = mch.Optimizer(0.25, 1.2, steps, precision=prec)
        plan = opt.prepare_batch(mol, lb, su, chains)
        plan.stage_inputs(seeds=list(range(chains))); plan.upload()
        for _ in range(2): plan.launch()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): plan.launch()
        e1.record(); e1.synchronize()
        out[steps] = e0.elapsed_time(e1) / 3
    per = (out[101] - out[21]) / 80
    print(prec, out, "per step %.4f ms; step0 = %.1f steps" % (per, out[1] / per))
