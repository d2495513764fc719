"""Throughput of the resident Metropolis kernel against the batch size (the strong-scaling question):
chain-steps/s of C chains of the 1000-atom cube cage for every (threads per chain, cluster) choice,
next to what the library picks on its own (threads 0, cluster 0).

    python profiles/chain_sweep.py [--precision fp32|fp64] [--steps 500] [--chains 148,256,...]
"""
import argparse
import sys

import torch

sys.path.insert(0, ".")
import mchammer_b200 as mch
from mchammer_b200 import engine, rngstate, synthetic
from mchammer_b200.lowering import lower_topology

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="fp32")
ap.add_argument("--steps", type=int, default=500)
ap.add_argument("--chains", default="148,256,512,1024,2048,4096")
ap.add_argument("--workload", default="cube1000")
ap.add_argument("--configs", default="0x0,64x1,128x1,256x1,512x1,64x2,128x2,256x2,128x4")
args = ap.parse_args()

mol, bonds, su = synthetic.cube_cage() if args.workload == "cube1000" else synthetic.cuboctahedron_cage()
opt = mch.Optimizer(0.25, 1.2, args.steps, precision=args.precision)
topo = lower_topology(mol.get_num_atoms(), bonds, su)
dtopo = engine.upload_topology(topo, 0)
dev = torch.device("cuda", 0)
pos = torch.from_numpy(mol.get_position_matrix()).to(dev)
configs = [tuple(int(v) for v in c.split("x")) for c in args.configs.split(",")]
full_rate = None
rows = []
for chains in sorted((int(c) for c in args.chains.split(",")), reverse=True):
    words = rngstate.seed_states(list(range(1000, 1000 + chains)))
    line = []
    for threads, cluster in configs:
        best = None
        try:
            for _ in range(3):
                rng = engine.rng_words_to_device(words, dev)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                engine.optimize_on_device(dtopo, pos, chains, rng, opt._params(), args.steps, args.precision,
                                          threads_per_chain=threads, cluster_size=cluster)
                e1.record(); e1.synchronize()
                ms = e0.elapsed_time(e1)
                best = ms if best is None else min(best, ms)
        except Exception as err:  # a combination the library refuses
            line.append(f"{threads}x{cluster}: refused")
            continue
        rate = chains * (args.steps - 1) / best * 1e3
        if chains == 4096 and (threads, cluster) == (0, 0):
            full_rate = rate
        line.append(f"{threads}x{cluster}: {best:7.2f} ms {rate:.3e}" + (f" ({rate / full_rate:.2f})" if full_rate else ""))
    print(f"{args.workload} {args.precision} chains {chains:5d} | " + " | ".join(line), flush=True)
