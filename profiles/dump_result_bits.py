"""Final geometries, potentials and accept counts of 300 fp64 chains of the bench cage as one .npy -- to compare two
builds bit for bit (MCHAMMER_B200_LIB=variants/libmch_A.so python profiles/dump_result_bits.py /tmp/a.npy; same for B;
np.array_equal).  Used for the changes that claim "the same bits" (warp_sum3, lazy molecule sums)."""
import sys, numpy as np
sys.path.insert(0, ".")
import mchammer_b200 as mch
from mchammer_b200 import synthetic
mol, bonds, su = synthetic.cube_cage()
opt = mch.Optimizer(0.25, 1.2, 300, precision="fp64")
r = opt.get_results_batch(mol, bonds, su, seeds=list(range(500, 500 + 300)))
np.save(sys.argv[1], np.concatenate([r.positions.reshape(300, -1), r.system_potential[:, None], r.nonbonded_potential[:, None], r.n_accepted[:, None].astype(float)], axis=1))
