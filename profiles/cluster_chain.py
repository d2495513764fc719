"""One chain on a thread-block cluster (2 / 4 / 8 CTAs = SMs): same results, lower latency.

    python profiles/cluster_chain.py
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import mchammer_b200 as mch
from mchammer_b200 import engine, rngstate, synthetic
from mchammer_b200.lowering import lower_topology


def run(name, mol, bonds, su, steps, precision, chains=1):
    opt = mch.Optimizer(0.25, 1.2, steps, precision=precision)
    topo = lower_topology(mol.get_num_atoms(), bonds, su)
    dtopo = engine.upload_topology(topo, 0)
    pos = torch.from_numpy(mol.get_position_matrix()).cuda()
    words = rngstate.seed_states(list(range(1000, 1000 + chains)))
    base = None
    for cluster in (1, 2, 4, 8, 16):
        best, out = 1e9, None
        try:
            for _ in range(4):
                rng = engine.rng_words_to_device(words, torch.device("cuda", 0))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out = engine.optimize_on_device(dtopo, pos, chains, rng, opt._params(), steps, precision,
                                                cluster_size=cluster)
                e1.record(); e1.synchronize()
                best = min(best, e0.elapsed_time(e1))
        except Exception as err:  # e.g. not enough shared memory for the cluster layout
            print(f"{name:9s} {precision} cluster {cluster}: {str(err)[:110]}")
            continue
        u = out.system_potential.cpu().numpy()
        acc = out.n_accepted.cpu().numpy()
        p = out.positions.cpu().numpy()
        if base is None:
            base = (u, acc, p)
        du = float(np.max(np.abs(u - base[0]) / np.abs(base[0])))
        dp = float(np.max(np.abs(p - base[2])))
        print(f"{name:9s} {precision} cluster {cluster}: {best:8.3f} ms = {1e3 * best / steps:7.2f} us/step   "
              f"accepted {acc.tolist()[:4]}  max rel dU vs 1 CTA {du:.1e}  max |dpos| {dp:.1e}")


cube = synthetic.cube_cage()
m12 = synthetic.cuboctahedron_cage()
for prec in ("fp64", "fp32"):
    run("cube1000", *cube, 200, prec)
for prec in ("fp32", "fp64"):
    run("m12l24", *m12, 200, prec)
run("m12l24x4", *m12, 100, "fp32", chains=4)
