"""One chain of a small molecule (poc.mol, 112 atoms), for ncu: the step time is the control section's latency."""
import sys

import torch

sys.path.insert(0, ".")
import mchammer_b200 as mch
from tests.conftest import golden_bonds, golden_subunits, load_golden

prec = sys.argv[1] if len(sys.argv) > 1 else "fp64"
g = load_golden("poc_graph_seed1000")
mol = mch.Molecule(
    atoms=tuple(mch.Atom(id=i, element_string=str(e)) for i, e in enumerate(g["elements"])),
    bonds=tuple(mch.Bond(id=i, atom_ids=tuple(int(x) for x in b)) for i, b in enumerate(g["bonds"])),
    position_matrix=g["pos0"])
opt = mch.Optimizer(0.25, 1.2, 500, precision=prec)
for _ in range(3):
    opt.get_result(mol, golden_bonds(g), golden_subunits(g))
torch.cuda.synchronize()
