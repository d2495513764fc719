"""Kernel time of ONE chain (latency case) for several block sizes."""
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import mchammer_b200 as mch
from mchammer_b200 import engine, rngstate, synthetic
from mchammer_b200.lowering import lower_topology
from tests.conftest import golden_bonds, golden_subunits, load_golden
def run(name, mol, bonds, su, steps, precision, threads=0):
    opt = mch.Optimizer(0.25, 1.2, steps, precision=precision)
    topo = lower_topology(mol.get_num_atoms(), bonds, su)
    dtopo = engine.upload_topology(topo, 0)
    pos = torch.from_numpy(mol.get_position_matrix()).cuda()
    words = rngstate.seed_states([1000])
    rng = engine.rng_words_to_device(words, torch.device("cuda", 0))
    best = 1e9
    for _ in range(5):
        r = rng.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        engine.optimize_on_device(dtopo, pos, 1, r, opt._params(), steps, precision, threads_per_chain=threads)
        e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{name:10s} {precision} threads {threads or 'auto':>4}: kernel {best:.3f} ms = {1e3*best/steps:.2f} us/step")


def mol_of(g):
    return mch.Molecule(
        atoms=tuple(mch.Atom(id=i, element_string=str(e)) for i, e in enumerate(g["elements"])),
        bonds=tuple(mch.Bond(id=i, atom_ids=tuple(int(x) for x in b)) for i, b in enumerate(g["bonds"])),
        position_matrix=g["pos0"])

for case, steps in (("benzene_seed1000", 100), ("poc_graph_seed1000", 500)):
    g = load_golden(case)
    for prec in ("fp64", "fp32"):
        for th in (0, 64, 128):
            run(case[:8], mol_of(g), golden_bonds(g), golden_subunits(g), steps, prec, th)
m, lb, su = synthetic.cube_cage()
for prec in ("fp64", "fp32"):
    for th in (0, 128, 256, 512):
        run("cube1000", m, lb, su, 500, prec, th)
m, lb, su = synthetic.cuboctahedron_cage()
for prec in ("fp64", "fp32"):
    for th in (0, 256, 512):
        run("m12l24", m, lb, su, 100, prec, th)
