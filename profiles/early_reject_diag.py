"""Where do a full and an early-rejection run of the same chains part?  Lists differing output fields per chain and
bisects the first num_steps at which a chain differs (how the implicit-FMA difference between the two instantiations
was found; with -fmad=false on the step kernel TUs nothing differs).

    python profiles/early_reject_diag.py
"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import mchammer_b200 as mch
from mchammer_b200 import engine, rngstate, synthetic
from mchammer_b200.lowering import lower_topology
from mchammer_b200.batch import unpack_info

mol, bonds, su = synthetic.cube_cage(atoms_per_subunit=12, seed=5, spread=6.5)
topo = lower_topology(mol.get_num_atoms(), bonds, su)
dtopo = engine.upload_topology(topo, 0)
dev = torch.device("cuda", 0)
pos = torch.from_numpy(mol.get_position_matrix()).to(dev)
precision = "fp64"
KEYS = ("positions", "system_potential", "nonbonded_potential", "max_bond_distance", "n_accepted", "last_step_info", "rng_state", "pair_evals")

def run(seeds, steps, early, trace=False):
    opt = mch.Optimizer(0.3, 1.2, steps, precision=precision)
    words = rngstate.seed_states(seeds)
    rng = engine.rng_words_to_device(words, dev)
    out = engine.optimize_on_device(dtopo, pos, len(seeds), rng, opt._params(), steps, precision, threads_per_chain=128,
                                    early_reject=early, want_trace=trace)
    torch.cuda.synchronize()
    return out

seeds = list(range(40, 40 + 96))
a, a2, b = run(seeds, 400, False), run(seeds, 400, False), run(seeds, 400, True)
for name, x, y in (("full vs full", a, a2), ("full vs early", a, b)):
    for k in KEYS[:-1]:
        xa, ya = getattr(x, k), getattr(y, k)
        d = (xa != ya)
        while d.dim() > 1: d = d.any(-1)
        print(name, k, "differing chains:", d.nonzero().flatten().tolist()[:20])
d = (a.system_potential != b.system_potential).nonzero().flatten().tolist()
if d:
    c = d[0]
    print("chain", c, "U full", repr(float(a.system_potential[c])), "early", repr(float(b.system_potential[c])),
          "Unb", repr(float(a.nonbonded_potential[c])), repr(float(b.nonbonded_potential[c])))
    lo, hi = 1, 400
    while hi - lo > 1:
        mid = (lo + hi) // 2
        x, y = run([seeds[c]], mid, False), run([seeds[c]], mid, True)
        if float(x.system_potential[0]) != float(y.system_potential[0]) or float(x.nonbonded_potential[0]) != float(y.nonbonded_potential[0]): hi = mid
        else: lo = mid
    print("first num_steps with a difference:", hi)
    t = run([seeds[c]], hi + 2, False, trace=True)
    bond, passed, kind, moved = unpack_info(t.trace_info[0].cpu().numpy())
    for s in range(max(0, hi - 6), hi + 2):
        print(s, "bond", bond[s], "passed", passed[s], "kind", kind[s], "ms", moved[s], "U", repr(float(t.trace_potentials[0, s, 0])), "Unb", repr(float(t.trace_potentials[0, s, 1])))
    for n in range(hi - 3, hi + 2):
        x, y = run([seeds[c]], n, False), run([seeds[c]], n, True)
        print("num_steps", n, "full U", repr(float(x.system_potential[0])), "early U", repr(float(y.system_potential[0])), "pairs", int(x.pair_evals[0]), int(y.pair_evals[0]), "nacc", int(x.n_accepted[0]), int(y.n_accepted[0]))
