"""Cluster cases for the check build (libmchammer_b200_check.so: -DMCH_RACECHECK = device asserts + the
race detector for the thread-block-cluster protocols, csrc/mch_device.cuh).  Prints one JSON line with the
results; a detected race or a failed assert traps the kernel (non-zero exit, message on stdout).

    MCHAMMER_B200_LIB=mchammer_b200/libmchammer_b200_check.so python profiles/check_build_case.py
"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import mchammer_b200 as mch
from mchammer_b200 import engine, rngstate, synthetic
from mchammer_b200.lowering import lower_topology

dev = torch.device("cuda", 0)
out = {}


def run(tag, pos, bonds, subunits, steps, precision, cluster, chains, split, threads=0, early=False):
    topo = lower_topology(pos.shape[0], bonds, subunits, allow_overlap=split)
    dtopo = engine.upload_topology(topo, 0)
    rng = engine.rng_words_to_device(rngstate.seed_states(list(range(40, 40 + chains))), dev)
    opt = mch.Optimizer(0.25, 1.2, steps, precision=precision)
    res = engine.optimize_on_device(dtopo, torch.from_numpy(np.ascontiguousarray(pos, dtype=np.float64)).to(dev),
                                    chains, rng, opt._params(), steps, precision, cluster_size=cluster,
                                    threads_per_chain=threads, force_split_state=split, early_reject=early)
    torch.cuda.synchronize()
    out[tag] = {"n_accepted": res.n_accepted.cpu().tolist(), "U": [float(u).hex() for u in res.system_potential.cpu().tolist()]}


mol, bonds, subunits = synthetic.cube_cage()
pos = mol.get_position_matrix()
for precision in ("fp64", "fp32"):
    for g in (2, 4, 8, 16):  # the replicated-state cluster of the one-CTA kernel (csrc/mch_optimize.cuh, CL = true)
        run(f"replicated {precision} G={g}", pos, bonds, subunits, 80, precision, g, 3, False)
    for g in (1, 2, 4, 8, 16):  # the split-state kernel (csrc/mch_wide.cuh)
        if g == 16 and precision == "fp64":
            continue
        run(f"split {precision} G={g}", pos, bonds, subunits, 30, precision, g, 2, True)
# overlapping subunits (the split-state kernel's own routing), small molecule spread over more CTAs than it has subunits
small, sb, ss = synthetic.cube_cage(atoms_per_subunit=6, seed=3, spread=5.0)
ids = sorted(ss)
ss[ids[0]] = list(ss[ids[0]]) + list(ss[ids[1]])[:2]  # two atoms listed by two subunits
for g in (1, 2, 4, 8):
    run(f"overlap fp64 G={g}", small.get_position_matrix(), sb, ss, 60, "fp64", g, 2, True, threads=64)
# early rejection (csrc/mch_optimize.cuh, EARLY = true): the control warp's loop over decided proposals under the
# descriptor asserts -- small subunits (one warp regroups) and the 208-atom linkers (all warps re-sum the row)
big, bb, bs = synthetic.cuboctahedron_cage()
for precision in ("fp64", "fp32"):
    run(f"early {precision} cube", pos, bonds, subunits, 120, precision, 1, 160, False, threads=128, early=True)
    run(f"early {precision} m12l24", big.get_position_matrix(), bb, bs, 40, precision, 1, 80, False, threads=256, early=True)
print(json.dumps(out))
