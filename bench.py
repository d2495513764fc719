#!/usr/bin/env python
"""bench.py -- MC chain-steps/s of the resident Metropolis kernel on the BASELINE workload.

    python bench.py --gpus N --steps K --warmup W            (own arm, CUDA)
    python bench.py --impl reference --gpus N --steps K ...  (CPU arm: the oracle port)

Workload (BASELINE.json configs[3]): 4096 independent chains PER GPU of the synthetic
1000-atom cube cage (20 rigid subunits x 50 atoms, 24 long bonds), Optimizer(step_size=0.25,
target_bond_length=1.2, num_steps=500) with reference defaults otherwise, distinct seeds.
One bench "step" = one launch of the whole batch = chains x (num_steps - 1) Monte-Carlo
moves.  `value` = chain-steps/s with inputs resident in HBM; `e2e` = the same through
Optimizer.prepare_batch(...).run(...) with host buffers (pinned H2D of per-chain positions +
RNG states, D2H of final positions + per-chain scalars inside the timed region).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = "mc_chain_steps_per_s"
UNIT = "chain-steps/s"
FLOPS_PER_PAIR = 13.0  # 10 FP-pipe instructions (FMA = 2 flops) + 1 MUFU per pair, mu = 3
WORKLOADS = {
    # name: (builder, atoms, pairs per move)
    "cube1000": dict(chains=4096, num_steps=500),
    "m12l24": dict(chains=1024, num_steps=100),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="cube1000", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp64"])
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (default: workload's)")
    ap.add_argument("--num-steps", type=int, default=0, help="Optimizer num_steps (default: workload's)")
    ap.add_argument("--threads", type=int, default=0, help="threads per chain (0 = library default)")
    ap.add_argument("--early-reject", default="auto", choices=["auto", "on", "off"],
                    help="early rejection in the step kernel (auto = the library's default for the precision mode)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the fp64 / M12L24 / strong-scaling lines")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline budget per core")
    ap.add_argument("--cpu-worker", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--cpu-moves", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--cpu-chains", type=int, default=0, help=argparse.SUPPRESS)
    return ap.parse_args()


def build_workload(name: str):
    from mchammer_b200 import synthetic

    if name == "cube1000":
        return synthetic.cube_cage()
    return synthetic.cuboctahedron_cage()


# =====================================================================================
# CPU arm: the UNMODIFIED reference (baseline/_ref, `pip install --target` of /root/reference,
# see DESIGN.md section 7) when it is on this box, else the oracle port (oracle/mch_oracle.py,
# labelled "port") -- one chain per process on all host cores either way
# =====================================================================================
REF_DIR = os.path.join(REPO, "baseline", "_ref")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "mchammer", "__init__.py"))


def _cpu_chain(job):
    """One chain of `moves` Monte-Carlo moves on the CPU; returns seconds."""
    import numpy as np

    kind, elements, bonds, pos, long_bonds, subunits, moves, seed = job
    if kind == "reference":
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        import mchammer as ref  # the unmodified reference package

        mol = ref.Molecule(
            atoms=tuple(ref.Atom(id=i, element_string=e) for i, e in enumerate(elements)),
            bonds=tuple(ref.Bond(id=i, atom_ids=b) for i, b in enumerate(bonds)),
            position_matrix=pos,
        )
        opt = ref.Optimizer(step_size=0.25, target_bond_length=1.2, num_steps=moves + 1, random_seed=seed)
        t0 = time.perf_counter()
        opt.get_result(mol, long_bonds, subunits)
        return time.perf_counter() - t0
    from oracle import mch_oracle as orc

    p = orc.OptimizerParams(step_size=0.25, target_bond_length=1.2, num_steps=moves + 1)
    t0 = time.perf_counter()
    orc.optimizer_run(pos, long_bonds, subunits, p, np.random.default_rng(seed))
    return time.perf_counter() - t0


def cpu_sample(workload: str, moves: int, chains: int, cores: int, kind: str):
    """Run `chains` CPU chains of `moves` moves over `cores` processes; wall seconds."""
    import multiprocessing as mp

    mol, long_bonds, subunits = build_workload(workload)
    pos = mol.get_position_matrix()
    elements = [a.get_element_string() for a in mol.get_atoms()]
    bonds = [(b.get_atom1_id(), b.get_atom2_id()) for b in mol.get_bonds()]
    jobs = [(kind, elements, bonds, pos, long_bonds, subunits, moves, 1000 + c) for c in range(chains)]
    ctx = mp.get_context("fork")
    with ctx.Pool(processes=cores) as pool:
        warm = [(kind, elements, bonds, pos, long_bonds, subunits, 2, 1) for _ in range(cores)]
        pool.map(_cpu_chain, warm)  # warm the workers (imports, first-call costs)
        t0 = time.perf_counter()
        per_chain = pool.map(_cpu_chain, jobs, chunksize=1)
        wall = time.perf_counter() - t0
    return wall, sum(per_chain)


def cpu_worker_main(args):
    """Child process entry (never imports torch, so fork() is safe)."""
    cores = os.cpu_count() or 1
    kind = "reference" if reference_available() else "port"
    wall, busy = cpu_sample(args.workload, args.cpu_moves, args.cpu_chains, cores, kind)
    print(json.dumps({"wall": wall, "busy": busy, "cores": cores, "kind": kind}))


def run_cpu_subprocess(workload: str, moves: int, chains: int):
    cmd = [sys.executable, os.path.abspath(__file__), "--cpu-worker", "--workload", workload,
           "--cpu-moves", str(moves), "--cpu-chains", str(chains)]
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=1500)
    if out.returncode != 0:
        raise RuntimeError("cpu worker failed: " + out.stderr[-2000:])
    return json.loads(out.stdout.strip().splitlines()[-1])


def _cpu_what(kind: str) -> str:
    if kind == "reference":
        return ("the unmodified reference (mchammer.Optimizer.get_result imported from baseline/_ref, "
                "pure Python + numpy/scipy)")
    return "oracle/mch_oracle.py (numpy/scipy restatement of the reference; baseline/_ref absent on this box)"


def cpu_baseline(workload: str, budget_s: float):
    """The CPU path on all host cores, bounded to about `budget_s` seconds per core."""
    cores = os.cpu_count() or 1
    n_atoms = 1000 if workload == "cube1000" else 5004
    est_move_s = 1.0e-2 * (n_atoms / 1000.0) ** 2  # ~100 moves/s/core at N = 1000
    moves = max(5, int(budget_s / est_move_s / 2))
    chains = 2 * cores
    r = run_cpu_subprocess(workload, moves, chains)
    value = chains * moves / r["wall"]
    return {
        "value": value, "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
        "sample": f"{chains} chains x {moves} moves of {workload} through {_cpu_what(r['kind'])}, "
                  f"one chain per process on {r['cores']} cores, "
                  f"{r['wall']:.1f} s wall; per-core rate {chains * moves / r['busy']:.1f} chain-steps/s",
    }


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    spec = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    n_atoms = 1000 if args.workload == "cube1000" else 5004
    est_move_s = 1.0e-2 * (n_atoms / 1000.0) ** 2
    moves = max(3, int(1.0 / est_move_s))  # ~1 s of work per chain per bench step
    chains = cores
    total = args.steps + args.warmup
    times, kind = [], "port"
    for k in range(total):
        r = run_cpu_subprocess(args.workload, moves, chains)
        kind = r["kind"]
        if k >= args.warmup:
            times.append(r["wall"])
    seconds = sum(times)
    value = args.steps * chains * moves / seconds
    sample = (f"each step = {chains} chains x {moves} moves of {args.workload} through {_cpu_what(kind)}, "
              f"one chain per process on {cores} cores")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * seconds / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"{args.workload}: {spec['chains']} chains/GPU x {spec['num_steps']} num_steps "
                               "(CPU arm runs a bounded sample of it)", "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# =====================================================================================
# clocks during the timed region
# =====================================================================================
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.thread = [], None, None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows]
        for row in inside:
            f = [x.strip() for x in row.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax = float(f[1]); power.append(float(f[2]))
            except ValueError:
                continue
            for name, flag in zip(names, f[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(power) if power else None}


# =====================================================================================
# own arm
# =====================================================================================
def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def profiled_traffic(precision: str):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` summary."""
    for rnd in ("r02", "r01"):  # latest committed capture first
        path = os.path.join(REPO, "profiles", f"{rnd}_ncu_full_{precision}.json")
        if os.path.exists(path):
            with open(path) as fh:
                return json.load(fh).get("dram_bytes_per_launch")
    return None


def probe_pipe_peak(torch, dtype_code: int, device) -> float:
    """TFLOP/s of a pure dependent-FMA kernel (8 chains/thread) -- the arithmetic-pipe roofline.
    The probe kernels live in the bench-only library profiles/probes (not in the product ABI)."""
    from profiles import probes

    return probes.pipe_peak(torch, dtype_code, device)


def time_resident(torch, dist, device, launch, n_timed: int, flush) -> float:
    """Seconds of `n_timed` launches (CUDA events around each, L2 flushed in between, one
    untimed warm-up launch first), MAX over ranks."""
    launch()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    events = []
    for _ in range(n_timed):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launch()
        e1.record()
        events.append((e0, e1))
    torch.cuda.synchronize()
    seconds = sum(a.elapsed_time(b) for a, b in events) * 1e-3
    if dist is not None:
        t = torch.tensor([seconds], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        seconds = float(t[0])
    return seconds


def secondary_lines(torch, dist, device, world: int, rank: int, flush, main_value: float, main_chains: int):
    """The other numbers BASELINE.json's metric names, measured in this same run, device resident,
    CUDA events, max over ranks (every rank runs them; rank 0 reports):
      fp64         parity mode on the headline workload (cube1000, 4096 chains/GPU, num_steps 500)
      m12l24_fp32  BASELINE configs[4]: 65,536 chains of the 5004-atom cage in all, 65,536 / world per GPU (strong)
      strong_4096  the FIXED 4096-chain cube1000 batch split over the ranks (strong scaling)
      collapser_fp64 / collapser_fp32, potential_fp64 / potential_fp32   the Collapser loop and compute_potential
                   on 4096 molecules of 1000 atoms per GPU (rank 0's numbers)
    """
    import mchammer_b200 as mch
    from mchammer_b200 import engine, native, rngstate, synthetic
    from mchammer_b200.lowering import lower_topology

    def resident(workload, chains, num_steps, precision, n_timed, early=False):
        mol, long_bonds, subunits = build_workload(workload)
        opt = mch.Optimizer(step_size=0.25, target_bond_length=1.2, num_steps=num_steps, precision=precision,
                            device=device.index)
        topo = lower_topology(mol.get_num_atoms(), long_bonds, subunits)
        dtopo = engine.upload_topology(topo, device.index)
        pos = torch.from_numpy(mol.get_position_matrix()).to(device)
        words = rngstate.seed_states([1000 + rank * chains + c for c in range(chains)])
        rng0 = engine.rng_words_to_device(words, device)
        rng = rng0.clone()
        state = {"out": None}

        def launch():
            rng.copy_(rng0)
            state["out"] = engine.optimize_on_device(dtopo, pos, chains, rng, opt._params(), num_steps, precision,
                                                     out=state["out"], early_reject=early)

        seconds = time_resident(torch, dist, device, launch, n_timed, flush)
        pairs = int(state["out"].pair_evals.sum().item())
        return seconds / n_timed, pairs, mol.get_num_atoms()

    out = {}
    # --- fp64 parity mode on the headline workload
    chains, num_steps = WORKLOADS["cube1000"]["chains"], WORKLOADS["cube1000"]["num_steps"]
    sec, pairs, n_atoms = resident("cube1000", chains, num_steps, "fp64", 2)
    peak64 = probe_pipe_peak(torch, native.MCH_F64, device) if rank == 0 else None
    ach = FLOPS_PER_PAIR * pairs / sec / 1e12
    out["fp64"] = {
        "workload": f"cube1000: {chains} chains/GPU x {n_atoms} atoms, num_steps={num_steps}, parity mode",
        "value": chains * world * (num_steps - 1) / sec, "unit": UNIT, "ms_per_launch": 1e3 * sec,
        "pair_evals_per_s": pairs * world / sec,
        "roofline": {"bound": "fp64_pipe", "achieved": ach, "peak": peak64, "unit": "TFLOP/s",
                     "frac": ach / peak64 if peak64 else None},
    }
    # --- early rejection (same results bit for bit, fewer pair terms; the library's default in parity mode): the
    # roofline lines above / below time the FULL evaluation, these two the same launches with the flag set
    full_pairs = pairs
    sec_e, pairs_e, _ = resident("cube1000", chains, num_steps, "fp64", 2, early=True)
    out["fp64_early_rejection"] = {
        "workload": out["fp64"]["workload"] + ", MCH_FLAG_EARLY_REJECTION (library default in this mode)",
        "value": chains * world * (num_steps - 1) / sec_e, "unit": UNIT, "ms_per_launch": 1e3 * sec_e,
        "pair_terms_vs_full": pairs_e / full_pairs, "speedup_vs_full": sec / sec_e,
        # what a kernel evaluating EVERY pair term would have to sustain for this rate -- the pair terms of the
        # reference algorithm per second, most of which this launch never executes; not a pipe utilisation
        "reference_equivalent_pipe_frac": (FLOPS_PER_PAIR * full_pairs / sec_e / 1e12 / peak64) if peak64 else None,
    }
    sec_e, pairs_e, _ = resident("cube1000", chains, num_steps, "fp32", 3, early=True)
    out["fp32_early_rejection"] = {
        "workload": f"cube1000: {chains} chains/GPU x {n_atoms} atoms, num_steps={num_steps}, fast mode, "
                    "MCH_FLAG_EARLY_REJECTION (opt-in in this mode)",
        "value": chains * world * (num_steps - 1) / sec_e, "unit": UNIT, "ms_per_launch": 1e3 * sec_e,
        "pair_terms_vs_full": pairs_e / full_pairs,
        "speedup_vs_full": (chains * (num_steps - 1) / sec_e) / main_value if main_chains == chains else None,
    }
    # --- M12L24 (BASELINE configs[4]): the FIXED 65,536-chain batch split over the ranks
    # (65,536 on one GPU ... 8192 per GPU on 8; final geometries of one GPU's share: 7.9 GB at N=1)
    m_chains = 65536 // world
    m_steps = WORKLOADS["m12l24"]["num_steps"]
    sec, pairs, n_atoms = resident("m12l24", m_chains, m_steps, "fp32", 2)
    peak32 = None
    if rank == 0:
        peak32 = max(probe_pipe_peak(torch, native.MCH_F32, device), probe_pipe_peak(torch, 3, device))
    ach = FLOPS_PER_PAIR * pairs / sec / 1e12
    out["m12l24_fp32"] = {
        "workload": f"m12l24: {m_chains} chains/GPU x {n_atoms} atoms ({m_chains * world} chains in all), "
                    f"num_steps={m_steps} (reduced from 500, BASELINE configs[4] allows it), fast mode",
        "value": m_chains * world * (m_steps - 1) / sec, "unit": UNIT, "ms_per_launch": 1e3 * sec,
        "pair_evals_per_s": pairs * world / sec,
        "scaling": "strong",
        "roofline": {"bound": "fp32_pipe", "achieved": ach, "peak": peak32, "unit": "TFLOP/s",
                     "frac": ach / peak32 if peak32 else None},
    }
    # --- the sibling kernels of the path (rank 0's GPU only; every rank runs them so that the ranks stay in step)
    import numpy as np

    def best_of(fn, reps=3):
        fn()
        best = None
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); e1.synchronize()
            t = e0.elapsed_time(e1) * 1e-3
            best = t if best is None else min(best, t)
        return best

    # Collapser loop (reference collapser.py:200-340): 4096 cages of 1000 atoms (30 % hydrogens), ~14 steps each
    cmol, cbonds, csub = synthetic.cube_cage(h_fraction=0.3)
    ctopo = lower_topology(cmol.get_num_atoms(), cbonds, csub, elements=[a.get_element_string() for a in cmol.get_atoms()],
                           require_bond_subunits=False)
    cdtopo = engine.upload_topology(ctopo, device.index)
    nh = ctopo.n_heavy.astype(np.int64)
    pairs_per_test = int((nh.sum() ** 2 - (nh ** 2).sum()) // 2)
    crng = np.random.default_rng(5)
    cbase = cmol.get_position_matrix()
    n_cages = 4096
    cpos = torch.from_numpy(np.stack([cbase * (1.0 + 0.004 * crng.random()) for _ in range(n_cages)])).to(device)
    for precision, peak in (("fp64", peak64), ("fp32", peak32)):
        res = engine.collapse_on_device(cdtopo, cpos, n_cages, 0.02, 1.5, True, precision)
        steps_run = int(res.steps_run.sum().item())
        sec = best_of(lambda: engine.collapse_on_device(cdtopo, cpos, n_cages, 0.02, 1.5, True, precision))
        rate = (steps_run + n_cages) * pairs_per_test / sec  # one contact test before the first step, one after each
        out["collapser_" + precision] = {
            "workload": f"Collapser(step_size=0.02, distance_threshold=1.5, scale_steps=True): {n_cages} cages x 1000 atoms "
                        f"({int(nh.sum())} non-H), {steps_run / n_cages:.1f} steps each, device resident",
            "value": steps_run / sec, "unit": "molecule-steps/s", "ms_per_launch": 1e3 * sec,
            "contact_pairs_per_s": rate,
            "roofline": {"bound": precision + "_pipe", "achieved": 9.0 * rate / 1e12, "peak": peak, "unit": "TFLOP/s",
                         "frac": 9.0 * rate / 1e12 / peak if peak else None,
                         "note": "9 flops per contact pair credited (the plain form: 3 sub + mul + 2 fma + compare/select); "
                                 "the fast form executes 6 flops + 1 min"},
        }
    del cpos
    # compute_potential for a batch (reference optimizer.py:114-142): 4096 molecules x 1000 atoms, all pairs
    pmol, pbonds, _ = build_workload("cube1000")
    n_p = pmol.get_num_atoms()
    ppos = torch.from_numpy(np.stack([pmol.get_position_matrix()] * 4096)).to(device)
    pb = torch.tensor(np.asarray(pbonds, dtype=np.int32).reshape(-1), dtype=torch.int32, device=device)
    popt = mch.Optimizer(step_size=0.25, target_bond_length=1.2, num_steps=2)
    for precision, peak in (("fp64", peak64), ("fp32", peak32)):
        sec = best_of(lambda: engine.potential_on_device(ppos, pb, len(pbonds), popt._params(), precision))
        rate = 4096 * (n_p * (n_p - 1) // 2) / sec
        out["potential_" + precision] = {
            "workload": f"compute_potential: 4096 molecules x {n_p} atoms, all pairs + {len(pbonds)} bond terms, device resident",
            "value": 4096 / sec, "unit": "molecules/s", "ms_per_launch": 1e3 * sec, "pair_evals_per_s": rate,
            "roofline": {"bound": precision + "_pipe", "achieved": FLOPS_PER_PAIR * rate / 1e12, "peak": peak,
                         "unit": "TFLOP/s", "frac": FLOPS_PER_PAIR * rate / 1e12 / peak if peak else None},
        }
    del ppos
    # --- strong scaling: the fixed 4096-chain batch over all ranks
    total = 4096
    num_steps = WORKLOADS["cube1000"]["num_steps"]
    if world == 1 and main_chains == total:
        value, ms = main_value, None
    else:
        per = total // world
        sec, _, _ = resident("cube1000", per, num_steps, "fp32", 5)
        value, ms = per * world * (num_steps - 1) / sec, 1e3 * sec
    out["strong_4096"] = {
        "workload": f"cube1000: {total} chains in all = {total // world} per GPU, num_steps={num_steps}, fast mode",
        "value": value, "unit": UNIT, "ms_per_launch": ms, "scaling": "strong",
        "note": "efficiency = value / (n_gpus x the 1-GPU value of this same line)",
    }
    # the same fixed batch with early rejection (opt-in in fast mode; same results bit for bit)
    if world == 1 and main_chains == total:
        value_e, ms_e = out["fp32_early_rejection"]["value"], out["fp32_early_rejection"]["ms_per_launch"]
    else:
        sec, _, _ = resident("cube1000", total // world, num_steps, "fp32", 5, early=True)
        value_e, ms_e = (total // world) * world * (num_steps - 1) / sec, 1e3 * sec
    out["strong_4096_early_rejection"] = {
        "workload": out["strong_4096"]["workload"] + ", MCH_FLAG_EARLY_REJECTION",
        "value": value_e, "unit": UNIT, "ms_per_launch": ms_e, "scaling": "strong",
    }
    return out


def cuda_arm(args):
    import numpy as np
    import torch

    import mchammer_b200 as mch
    from mchammer_b200 import native

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise native.NativeLibraryError("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=device)

    spec = WORKLOADS[args.workload]
    chains = args.chains or spec["chains"]
    num_steps = args.num_steps or spec["num_steps"]
    mol, long_bonds, subunits = build_workload(args.workload)
    n_atoms = mol.get_num_atoms()
    opt = mch.Optimizer(step_size=0.25, target_bond_length=1.2, num_steps=num_steps,
                        precision=args.precision, device=local)
    early = {"auto": None, "on": True, "off": False}[args.early_reject]
    plan = opt.prepare_batch(mol, long_bonds, subunits, chains, devices=[local],
                             per_chain_positions=True, threads_per_chain=args.threads, early_reject=early)
    seeds = [1000 + rank * chains + c for c in range(chains)]
    moves = num_steps - 1

    # ---------------- device-resident timing (value, roofline)
    plan.stage_inputs(seeds=seeds)
    plan.upload()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)  # > 126 MB L2
    for _ in range(max(args.warmup, 3)):
        plan.launch()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    events = []
    t_begin = time.perf_counter()
    torch.cuda.synchronize()
    for _ in range(args.steps):
        flush.zero_()  # untimed: evict the previous step's lines from L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.launch()
        e1.record()
        events.append((e0, e1))
    torch.cuda.synchronize()
    t_end = time.perf_counter()
    kernel_ms = [a.elapsed_time(b) for a, b in events]
    dev_seconds = sum(kernel_ms) * 1e-3
    pairs_per_launch = int(plan.shards[0].out.pair_evals.sum().item())
    gpu_launches_e2e = len(plan._slices(chains))
    # in-run validity of what was just timed: the potential every chain CARRIED through its moves against a fresh
    # float64 all-pairs evaluation (the potential kernel) of the chain's own final geometry
    from mchammer_b200 import engine as _engine
    _out = plan.shards[0].out
    _sample = min(chains, 512)
    _bdev = torch.tensor(np.asarray(long_bonds, dtype=np.int32).reshape(-1), dtype=torch.int32, device=device)
    _fresh, _, _ = _engine.potential_on_device(_out.positions[:_sample].to(torch.float64), _bdev, len(long_bonds),
                                               opt._params(), "fp64")
    _carried = _out.system_potential[:_sample]
    carried_vs_fresh = float(((_carried - _fresh).abs() / _fresh.abs()).max().item())

    # ---------------- end to end through the public API (host buffers in, host buffers out)
    host_positions = plan.h_pos_in  # pinned [C,N,3]; per-chain start geometries
    rng_words = mch.rngstate.seed_states(seeds)
    for _ in range(2):
        plan.run(rng_words=rng_words, positions=host_positions)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e2e_t0 = time.perf_counter()
    h2d = d2h = 0
    for _ in range(args.steps):
        res = plan.run(rng_words=rng_words, positions=host_positions)
        h2d, d2h = res.h2d_bytes, res.d2h_bytes
    torch.cuda.synchronize()
    e2e_seconds = time.perf_counter() - e2e_t0
    clocks = sampler.stop(t_begin, time.perf_counter())
    accepted = float(res.n_accepted.mean())

    # ---------------- the cold public call: seeds in, results out, nothing prepared
    # (plan allocation + pinning + numpy-compatible seeding + copies + kernel + result copy)
    cold_t0 = time.perf_counter()
    cold = opt.get_results_batch(mol, long_bonds, subunits, seeds=seeds, devices=[local])
    cold_seconds = time.perf_counter() - cold_t0
    cold_ok = bool(np.array_equal(cold.n_accepted, res.n_accepted))
    del cold

    # ---------------- the other configurations the metric names (fp64, M12L24, strong scaling)
    secondary = None
    if not args.no_secondary and args.workload == "cube1000" and args.precision == "fp32":
        del plan, res
        torch.cuda.empty_cache()
        secondary = secondary_lines(torch, dist, device, world, rank, flush,
                                    args.steps * chains * moves / dev_seconds, chains)

    # ---------------- reduce over ranks: slowest rank sets the time
    if dist is not None:
        t = torch.tensor([dev_seconds, e2e_seconds], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_seconds, e2e_seconds = float(t[0]), float(t[1])
    total_chain_steps = args.steps * chains * moves * world
    value = total_chain_steps / dev_seconds
    e2e_value = total_chain_steps / e2e_seconds

    if rank == 0:
        peaks, peaks_src = measured_peaks()
        dtype_code = native.MCH_F32 if args.precision == "fp32" else native.MCH_F64
        pipe_peak = probe_pipe_peak(torch, dtype_code, device)
        if args.precision == "fp32":  # the packed FFMA2 stream is the faster of the two fp32 probes
            pipe_peak = max(pipe_peak, probe_pipe_peak(torch, 3, device))
        mufu_peak = probe_pipe_peak(torch, 2, device) if args.precision == "fp32" else None
        launch_s = dev_seconds / args.steps
        achieved_tflops = FLOPS_PER_PAIR * pairs_per_launch / launch_s / 1e12
        io_bytes = chains * n_atoms * 3 * 8 * 2 + chains * 48 * 2
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * dev_seconds / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == "fp32" else "f64", "data": "synthetic",
            "config": {
                "workload": f"{args.workload}: {chains} chains/GPU x {n_atoms} atoms "
                            f"({len(subunits)} subunits, {len(long_bonds)} long bonds), num_steps={num_steps}, "
                            f"seeds 1000+c, identical start geometry per chain",
                "chains_per_gpu": chains, "num_steps": num_steps, "precision": args.precision,
                "io_dtype": "f64", "l2": "flushed between timed steps (256 MiB memset, untimed)",
                "threads_per_chain": args.threads or "auto",
                "early_rejection": _engine.resolve_early_reject(early, args.precision),
            },
            "pair_evals_per_s": pairs_per_launch * world / launch_s,
            "reference_equivalent_pair_evals_per_s": value * (n_atoms * (n_atoms - 1) // 2),
            "accepted_moves_per_chain": accepted,
            "carried_potential_check": {
                "max_rel_diff": carried_vs_fresh, "chains": min(chains, 512),
                "what": "after the timed launch: |U carried through all moves - fresh float64 all-pairs potential of the "
                        "chain's final geometry| / U, max over the sampled chains (tolerance of the fast mode: 1e-4; "
                        "parity mode: 1e-9)"},
            "roofline": {
                "bound": "fp32_pipe" if args.precision == "fp32" else "fp64_pipe",
                "achieved": achieved_tflops, "peak": pipe_peak, "unit": "TFLOP/s",
                "frac": achieved_tflops / pipe_peak if pipe_peak else None,
                "traffic": profiled_traffic(args.precision) if args.workload == "cube1000" and chains == 4096 else None,
                "kernel": "mch_optimize_kernel",
                "note": f"achieved = {FLOPS_PER_PAIR:.0f} flops/pair x {pairs_per_launch} pairs per launch / "
                        "CUDA-event launch time; peak = best dependent-FMA probe (mch_probe_fma: FFMA and packed "
                        "FFMA2 streams) measured in this run; the kernel is bound by the CUDA-core pipes (FP32/FP64 "
                        "FMA and, in fp32 mode, one MUFU.RSQ per pair), not by HBM or tensor cores (SURVEY.md 8-d)",
                "mufu_rsqrt_per_s_peak": mufu_peak * 1e12 if mufu_peak else None,
                "mufu_frac": pairs_per_launch / launch_s / (mufu_peak * 1e12) if mufu_peak else None,
                "hbm": {"achieved_gbs": io_bytes / launch_s / 1e9, "peak_gbs": peaks.get("hbm_gbs"),
                        "peak_source": peaks_src},
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "Optimizer.prepare_batch(...).run(rng_words, positions) with pinned host buffers; "
                           "the call pipelines H2D / kernel / D2H over 4 slices of the batch",
                    "cold_public_call": {
                        "api": "Optimizer.get_results_batch(mol, bonds, subunits, seeds): plan allocation + pinning + "
                               "PCG64 seeding + H2D + kernel + D2H + result copy, one call per rank on its own GPU (rank 0's time)",
                        "seconds": cold_seconds, "value": chains * moves / cold_seconds, "unit": UNIT,
                        "same_accept_counts_as_prepared_plan": cold_ok}},
            "secondary": secondary,
            "gpu_launches": args.steps,
            "gpu_launches_e2e": args.steps * gpu_launches_e2e,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.workload, args.cpu_seconds)
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.cpu_worker:
        cpu_worker_main(args)
    elif args.impl == "reference":
        reference_arm(args)
    else:
        cuda_arm(args)


if __name__ == "__main__":
    main()
