"""Where the host time of one drop-in call goes (small molecule: the kernel is a fraction of the call).

    python profiles/host_overhead.py
"""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, ".")
import mchammer_b200 as mch
from tests.conftest import golden_bonds, golden_subunits, load_golden


def mol_of(g):
    return mch.Molecule(
        atoms=tuple(mch.Atom(id=i, element_string=str(e)) for i, e in enumerate(g["elements"])),
        bonds=tuple(mch.Bond(id=i, atom_ids=tuple(int(x) for x in b)) for i, b in enumerate(g["bonds"])),
        position_matrix=g["pos0"])


for case, steps in (("benzene_seed1000", 100), ("poc_graph_seed1000", 500)):
    g = load_golden(case)
    mol, lb, su = mol_of(g), golden_bonds(g), golden_subunits(g)
    opt = mch.Optimizer(0.25, 1.2, steps)
    for _ in range(5):
        opt.get_result(mol, lb, su)
    torch.cuda.synchronize()
    n = 200
    t0 = time.perf_counter()
    for _ in range(n):
        opt.get_result(mol, lb, su)
    torch.cuda.synchronize()
    print(f"{case}: get_result {1e3 * (time.perf_counter() - t0) / n:.3f} ms per call")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(n):
        opt.get_result(mol, lb, su)
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats("cumulative").print_stats(18)
