"""Where a swept step's cycles go, per warp (tuning build with -DMCH_TIMING: the kernel overwrites the first 12
output coordinates of every chain with clock64 sums): deferred work after barrier A, the warp's own share of the
sweep, its wait at barrier B, the control section between B and the next A.

    bash profiles/quick_variant.sh timing f32 -DMCH_TIMING
    MCHAMMER_B200_LIB=variants/libmch_timing.so python profiles/phase_clocks.py [--precision fp32] [--chains 4096]
"""
import argparse
import sys

import torch

sys.path.insert(0, ".")
import mchammer_b200 as mch
from mchammer_b200 import engine, rngstate, synthetic
from mchammer_b200.lowering import lower_topology

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="fp32")
ap.add_argument("--chains", type=int, default=4096)
ap.add_argument("--steps", type=int, default=500)
args = ap.parse_args()
mol, bonds, su = synthetic.cube_cage()
topo = lower_topology(mol.get_num_atoms(), bonds, su)
dtopo = engine.upload_topology(topo, 0)
dev = torch.device("cuda", 0)
pos = torch.from_numpy(mol.get_position_matrix()).to(dev)
opt = mch.Optimizer(0.25, 1.2, args.steps, precision=args.precision)
words = rngstate.seed_states(list(range(1000, 1000 + args.chains)))
for early in (False, True):
    for _ in range(2):
        rng = engine.rng_words_to_device(words, dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = engine.optimize_on_device(dtopo, pos, args.chains, rng, opt._params(), args.steps, args.precision,
                                        early_reject=early)
        e1.record(); e1.synchronize()
    t = out.positions.reshape(args.chains, -1)[:, :12].cpu().numpy()
    n = t[:, 4].mean()
    print(f"{args.precision} early={early}: {e0.elapsed_time(e1):.2f} ms, swept steps per chain {n:.1f}; cycles per swept step:")
    for w in (0, 1):
        d, s, wt, c = (t[:, 6 * w + k].sum() / t[:, 6 * w + 4].sum() for k in range(4))
        print(f"  warp {w}: after A -> sweep start {d:8.0f} | own sweep {s:8.0f} | wait at B {wt:8.0f} | B -> next A {c:8.0f}")
