import sys, numpy as np, torch
sys.path.insert(0, ".")
import mchammer_b200 as mch
from mchammer_b200 import engine, rngstate, synthetic
from mchammer_b200.lowering import lower_topology
PREC = sys.argv[1] if len(sys.argv) > 1 else "fp32"
mol, bonds, su = synthetic.cuboctahedron_cage()
opt = mch.Optimizer(0.25, 1.2, 200, precision=PREC)
topo = lower_topology(mol.get_num_atoms(), bonds, su)
dtopo = engine.upload_topology(topo, 0)
pos = torch.from_numpy(mol.get_position_matrix()).cuda()
for _ in range(2):
    rng = engine.rng_words_to_device(rngstate.seed_states([1000]), torch.device("cuda", 0))
    out = engine.optimize_on_device(dtopo, pos, 1, rng, opt._params(), 200, PREC, cluster_size=8)
torch.cuda.synchronize()
print(out.system_potential.item())
