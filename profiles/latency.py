"""Single-molecule latency of the drop-in API (what a reference user sees) vs the CPU oracle port."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import mchammer_b200 as mch
from mchammer_b200 import synthetic
from oracle import mch_oracle as orc
from tests.conftest import golden_bonds, golden_subunits, load_golden


def mol_of(g):
    return mch.Molecule(
        atoms=tuple(mch.Atom(id=i, element_string=str(e)) for i, e in enumerate(g["elements"])),
        bonds=tuple(mch.Bond(id=i, atom_ids=tuple(int(x) for x in b)) for i, b in enumerate(g["bonds"])),
        position_matrix=g["pos0"])


def best(fn, n=5):
    ts = []
    for _ in range(n):
        t = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
    return min(ts)


cases = []
g = load_golden("benzene_seed1000"); cases.append(("benzene N=12, 100 steps", mol_of(g), golden_bonds(g), golden_subunits(g), 100))
g = load_golden("poc_graph_seed1000"); cases.append(("poc N=112, 500 steps", mol_of(g), golden_bonds(g), golden_subunits(g), 500))
m, lb, su = synthetic.cube_cage(); cases.append(("cube cage N=1000, 500 steps", m, lb, su, 500))
for name, mol, lb, su, steps in cases:
    for prec in ("fp64", "fp32"):
        opt = mch.Optimizer(0.25, 1.2, steps, precision=prec)
        opt.get_result(mol, lb, su)
        t_res = best(lambda: opt.get_result(mol, lb, su))
        t_traj = best(lambda: opt.get_trajectory(mol, lb, su), 3)
        print(f"{name:30s} {prec}: get_result {1e3 * t_res:8.2f} ms   get_trajectory {1e3 * t_traj:8.2f} ms")
    p = orc.OptimizerParams(0.25, 1.2, min(steps, 100))
    t = time.perf_counter(); orc.optimizer_run(mol.get_position_matrix(), lb, su, p, np.random.default_rng(1)); t = time.perf_counter() - t
    print(f"{name:30s} oracle port: {1e3 * t * steps / min(steps, 100):8.2f} ms (scaled from {min(steps, 100)} steps)")

m12, lb12, su12 = synthetic.cuboctahedron_cage()
for prec in ("fp64", "fp32"):
    opt = mch.Optimizer(0.25, 1.2, 500, precision=prec)
    opt.get_result(m12, lb12, su12)
    print("%-30s %s: get_result %8.2f ms  (one chain on a thread-block cluster, chosen automatically)" % (
        "M12L24 cage N=5004, 500 steps", prec, 1e3 * best(lambda: opt.get_result(m12, lb12, su12), 3)))
g = load_golden("collapser_poc")
mol = mol_of(g)
coll = mch.Collapser(0.05, 1.2, True)
coll.get_result(mol, golden_bonds(g), golden_subunits(g))
print("Collapser poc (13 steps): get_result %.2f ms, get_trajectory %.2f ms" % (
    1e3 * best(lambda: coll.get_result(mol, golden_bonds(g), golden_subunits(g))),
    1e3 * best(lambda: coll.get_trajectory(mol, golden_bonds(g), golden_subunits(g)), 3)))
m, lb, su = synthetic.cube_cage(h_fraction=0.3)
coll = mch.Collapser(0.02, 1.5, True)
_, last = coll.get_result(m, lb, su)
print("Collapser cube cage N=1000: %d steps, get_result %.2f ms" % (last.step, 1e3 * best(lambda: coll.get_result(m, lb, su))))
pos = np.stack([m.get_position_matrix() * (1.0 + 0.002 * k) for k in range(1024)])
res = coll.get_results_batch(m, lb, su, pos)
t = best(lambda: coll.get_results_batch(m, lb, su, pos), 3)
print("Collapser batch 1024 cages: %.1f ms, %d total steps, %.3e molecule-steps/s" % (1e3 * t, res.steps_run.sum(), res.steps_run.sum() / t))
