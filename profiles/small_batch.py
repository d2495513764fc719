"""Batched throughput on a small molecule (poc.mol, N=112): chain-steps/s."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import mchammer_b200 as mch
from tests.conftest import golden_bonds, golden_subunits, load_golden
g = load_golden("poc_graph_seed1000")
mol = mch.Molecule(atoms=tuple(mch.Atom(id=i, element_string=str(e)) for i, e in enumerate(g["elements"])),
                   bonds=tuple(mch.Bond(id=i, atom_ids=tuple(int(x) for x in b)) for i, b in enumerate(g["bonds"])),
                   position_matrix=g["pos0"])
chains, steps = 16384, 500
for prec in ("fp32", "fp64"):
    for threads in (32, 64):
        opt = mch.Optimizer(0.25, 1.2, steps, precision=prec)
        plan = opt.prepare_batch(mol, golden_bonds(g), golden_subunits(g), chains, threads_per_chain=threads)
        plan.stage_inputs(seeds=list(range(chains))); plan.upload()
        for _ in range(2): plan.launch()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): plan.launch()
        e1.record(); e1.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"poc {prec} threads {threads}: {ms:.2f} ms per {chains} chains x {steps} steps = {chains * (steps - 1) / ms * 1e3:.3e} chain-steps/s")
