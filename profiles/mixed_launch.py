"""Feasibility of balancing a batch that does not divide over the SMs: most chains as one-CTA chains, the
remainder as 2-CTA clusters in a concurrent launch, so that every SM carries the same number of CTAs.

    python profiles/mixed_launch.py [--chains 512] [--precision fp32]
"""
import argparse
import sys

import torch

sys.path.insert(0, ".")
import mchammer_b200 as mch
from mchammer_b200 import engine, rngstate, synthetic
from mchammer_b200.lowering import lower_topology

ap = argparse.ArgumentParser()
ap.add_argument("--chains", type=int, default=512)
ap.add_argument("--precision", default="fp32")
ap.add_argument("--steps", type=int, default=500)
args = ap.parse_args()
mol, bonds, su = synthetic.cube_cage()
opt = mch.Optimizer(0.25, 1.2, args.steps, precision=args.precision)
topo = lower_topology(mol.get_num_atoms(), bonds, su)
dtopo = engine.upload_topology(topo, 0)
dev = torch.device("cuda", 0)
pos = torch.from_numpy(mol.get_position_matrix()).to(dev)
C = args.chains
words = rngstate.seed_states(list(range(1000, 1000 + C)))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    best = None
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record(); e1.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return best


def single(threads):
    rng = engine.rng_words_to_device(words, dev)
    return lambda: engine.optimize_on_device(dtopo, pos, C, rng, opt._params(), args.steps, args.precision,
                                             threads_per_chain=threads, cluster_size=1)


def mixed(k, threads_single, threads_cluster, cluster_first):
    rng_a = engine.rng_words_to_device(words[: C - k], dev)
    rng_b = engine.rng_words_to_device(words[C - k:], dev)

    def go():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur); s2.wait_stream(cur)

        def a():
            with torch.cuda.stream(s1):
                engine.optimize_on_device(dtopo, pos, C - k, rng_a, opt._params(), args.steps, args.precision,
                                          threads_per_chain=threads_single, cluster_size=1, resident_chains_hint=C)

        def b():
            with torch.cuda.stream(s2):
                engine.optimize_on_device(dtopo, pos, k, rng_b, opt._params(), args.steps, args.precision,
                                          threads_per_chain=threads_cluster, cluster_size=2)
        (b(), a()) if cluster_first else (a(), b())
        cur.wait_stream(s1); cur.wait_stream(s2)
    return go


base = timed(single(128))
print(f"{C} chains, one launch, 128 threads: {base:.3f} ms")
sms = 148
q = -(-C // sms)
k_full = q * sms - C
for k in sorted({k_full // 2, k_full, min(C, k_full + sms // 2)}):
    if k <= 0:
        continue
    for tc in (128, 256):
        for first in (True, False):
            ms = timed(mixed(k, 128, tc, first))
            print(f"  {C - k} one-CTA chains + {k} chains on 2-CTA clusters of {tc} threads, "
                  f"{'clusters' if first else 'singles'} launched first: {ms:.3f} ms ({base / ms:.3f}x)")
