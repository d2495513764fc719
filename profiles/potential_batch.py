"""compute_potential for a batch of molecules (mch_potential_batch), device resident:
all N(N-1)/2 pair terms + bond terms per molecule; pair rate and FP-pipe fraction (13 flops/pair).

    python profiles/potential_batch.py [n_mol]
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import mchammer_b200 as mch
from mchammer_b200 import engine, synthetic

n_mol = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda", 0)
opt = mch.Optimizer(step_size=0.25, target_bond_length=1.2, num_steps=2)
rng = np.random.default_rng(3)
for name, (mol, bonds, subunits) in (("cube1000", synthetic.cube_cage()), ("m12l24", synthetic.cuboctahedron_cage())):
    n = mol.get_num_atoms()
    count = n_mol if n <= 1000 else max(n_mol // 16, 1)
    base = mol.get_position_matrix()
    pos = torch.from_numpy(np.stack([base * (1.0 + 0.01 * rng.random()) for _ in range(count)])).to(dev)
    bonds_dev = torch.tensor(np.asarray(bonds, dtype=np.int32).reshape(-1), dtype=torch.int32, device=dev)
    pairs = count * (n * (n - 1) // 2)
    for precision in ("fp64", "fp32"):
        engine.potential_on_device(pos, bonds_dev, len(bonds), opt._params(), precision)
        best = 1e30
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            engine.potential_on_device(pos, bonds_dev, len(bonds), opt._params(), precision)
            e1.record(); e1.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e-3)
        print(f"compute_potential {name} {precision}: {count} molecules x {n} atoms in {1e3 * best:.2f} ms = "
              f"{count / best:.3e} molecules/s, {pairs / best:.3e} pair terms/s, {13 * pairs / best / 1e12:.1f} TFLOP/s (13 flops/pair)")
