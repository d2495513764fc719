"""Small case for compute-sanitizer: every kernel once (both precisions, trace on, Collapser)."""
import sys

import numpy as np

sys.path.insert(0, ".")
import mchammer_b200 as mch
from mchammer_b200 import synthetic

mol, lb, su = synthetic.cube_cage(atoms_per_subunit=7, seed=3, spread=5.0, h_fraction=0.2)
for prec in ("fp64", "fp32"):
    opt = mch.Optimizer(0.25, 1.2, 30, precision=prec)
    res = opt.get_results_batch(mol, lb, su, seeds=[1, 2, 3])
    m2, traj = opt.get_trajectory(mol, lb, su)
    u, unb = opt.compute_potential(mol, lb)
    print(prec, res.system_potential, traj.get_step_count(), u)
coll = mch.Collapser(0.05, 1.3, True)
m3, r = coll.get_trajectory(mol, lb, su)
print("collapser steps", r.get_step_count())
big, lb2, su2 = synthetic.cube_cage()
res = mch.Optimizer(0.25, 1.2, 12, precision="fp32").get_results_batch(big, lb2, su2, seeds=[5, 6])
print("cube", res.n_accepted)
