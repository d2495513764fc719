"""Collapser kernel alone, device resident: molecule-steps/s, contact-test pair rate and the
fraction of the FP pipe peak (9 flops per pair: 3 sub + mul + 2 fma + compare-select).

    python profiles/collapse_batch.py [n_mol [fp64|fp32]]
"""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from mchammer_b200 import engine, native, synthetic
from mchammer_b200.lowering import lower_topology

n_mol = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
FLOPS_PER_PAIR = 9.0
dev = torch.device("cuda", 0)
mol, bonds, subunits = synthetic.cube_cage(h_fraction=0.3)
elements = [a.get_element_string() for a in mol.get_atoms()]
topo = lower_topology(mol.get_num_atoms(), bonds, subunits, elements=elements, require_bond_subunits=False)
dtopo = engine.upload_topology(topo, 0)
nh = topo.n_heavy.astype(np.int64)
pairs_per_test = int((nh.sum() ** 2 - (nh ** 2).sum()) // 2)
rng = np.random.default_rng(5)
base = mol.get_position_matrix()
pos = torch.from_numpy(np.stack([base * (1.0 + 0.004 * rng.random()) for _ in range(n_mol)])).to(dev)


def pipe_peak(code):
    from profiles import probes
    return probes.pipe_peak(torch, code, dev) * 1e12


for precision in ((sys.argv[2],) if len(sys.argv) > 2 else ("fp64", "fp32")):
    peak = pipe_peak(native.MCH_F64) if precision == "fp64" else max(pipe_peak(native.MCH_F32), pipe_peak(3))
    for threads in (0, 128, 256):
        out = engine.collapse_on_device(dtopo, pos, n_mol, 0.02, 1.5, True, precision, threads_per_mol=threads)
        steps = int(out.steps_run.sum().item())
        tests = steps + n_mol  # one contact test before the first step and one after every step
        best = 1e30
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            engine.collapse_on_device(dtopo, pos, n_mol, 0.02, 1.5, True, precision, threads_per_mol=threads)
            e1.record(); e1.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e-3)
        rate = tests * pairs_per_test / best
        print(f"collapse {precision} threads {threads or 'auto':>4}: {n_mol} cages x {steps / n_mol:.1f} steps in {1e3 * best:.2f} ms"
              f" = {steps / best:.3e} molecule-steps/s, {rate:.3e} contact pairs/s,"
              f" {FLOPS_PER_PAIR * rate / 1e12:.2f} TFLOP/s = {FLOPS_PER_PAIR * rate / peak:.3f} of the {precision} pipe peak ({peak / 1e12:.1f})")
