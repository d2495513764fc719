"""Early rejection (MCH_FLAG_EARLY_REJECTION) on / off: chain-steps/s and pair terms evaluated for batches of the
1000-atom cube cage, both precision modes, with a bit-for-bit comparison of the two runs' results.

    python profiles/early_reject_ab.py [--steps 500] [--chains 4096,2048,1024,512] [--workload cube1000|m12l24]
"""
import argparse
import sys

import torch

sys.path.insert(0, ".")
import mchammer_b200 as mch
from mchammer_b200 import engine, rngstate, synthetic
from mchammer_b200.lowering import lower_topology

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=500)
ap.add_argument("--chains", default="4096,2048,1024,512")
ap.add_argument("--precisions", default="fp32,fp64")
ap.add_argument("--workload", default="cube1000")
args = ap.parse_args()

mol, bonds, su = synthetic.cube_cage() if args.workload == "cube1000" else synthetic.cuboctahedron_cage()
topo = lower_topology(mol.get_num_atoms(), bonds, su)
dtopo = engine.upload_topology(topo, 0)
dev = torch.device("cuda", 0)
pos = torch.from_numpy(mol.get_position_matrix()).to(dev)
for precision in args.precisions.split(","):
    opt = mch.Optimizer(0.25, 1.2, args.steps, precision=precision)
    for chains in (int(c) for c in args.chains.split(",")):
        words = rngstate.seed_states(list(range(1000, 1000 + chains)))
        runs = {}
        for early in (False, True):
            best, out = None, None
            for _ in range(3):
                rng = engine.rng_words_to_device(words, dev)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out = engine.optimize_on_device(dtopo, pos, chains, rng, opt._params(), args.steps, precision,
                                                early_reject=early)
                e1.record(); e1.synchronize()
                ms = e0.elapsed_time(e1)
                best = ms if best is None else min(best, ms)
            runs[early] = (best, out)
        (t0, a), (t1, b) = runs[False], runs[True]
        same = all(torch.equal(getattr(a, k), getattr(b, k)) for k in
                   ("positions", "system_potential", "nonbonded_potential", "max_bond_distance", "n_accepted",
                    "last_step_info", "rng_state"))
        rate0, rate1 = chains * (args.steps - 1) / t0 * 1e3, chains * (args.steps - 1) / t1 * 1e3
        pairs0, pairs1 = int(a.pair_evals.sum()), int(b.pair_evals.sum())
        print(f"{args.workload} {precision} chains {chains:5d} | full {t0:7.2f} ms {rate0:.3e} | early {t1:7.2f} ms "
              f"{rate1:.3e} ({rate1 / rate0:.3f}) | pair terms {pairs1 / pairs0:.3f} | same bits: {same}", flush=True)
