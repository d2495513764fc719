"""GPU parity tests of the split-state kernel (csrc/mch_wide.cuh; through the C ABI): overlapping
subunits with the reference's semantics, chains whose state is split over a thread-block
cluster, and molecules beyond what one SM's shared memory holds.

Goldens are outputs of the unmodified reference (tests/golden/make_golden.py, round-2 cases);
everything else is compared with the CPU oracle on the same seeded inputs.
"""

from __future__ import annotations

import numpy as np
import pytest

import mchammer_b200 as mch
from mchammer_b200 import engine, native, rngstate, synthetic
from mchammer_b200.lowering import lower_topology
from oracle import mch_oracle as orc
from tests.conftest import golden_bonds, golden_subunits, load_golden
from tests.test_gpu_parity import RTOL32, RTOL64, golden_molecule, golden_optimizer, rel_close

pytestmark = pytest.mark.gpu


def run_split(pos, bonds, subunits, params, steps, seed, precision="fp64", cluster=0, threads=0, trace=True):
    """One chain through the split-state kernel (forced), straight through the engine."""
    import torch

    topo = lower_topology(pos.shape[0], bonds, subunits, allow_overlap=True)
    dtopo = engine.upload_topology(topo, 0)
    dev = torch.device("cuda", 0)
    rng = engine.rng_words_to_device(rngstate.seed_states([seed]), dev)
    out = engine.optimize_on_device(
        dtopo, torch.from_numpy(np.ascontiguousarray(pos, dtype=np.float64)).to(dev), 1, rng, params, steps,
        precision, want_trace=trace, want_trajectory=trace, cluster_size=cluster, threads_per_chain=threads,
        force_split_state=True)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("case", ["benzene_overlap_seed9", "benzene_dup_seed4", "poc_overlap_seed3"])
def test_overlapping_subunits_match_the_reference(case):
    # reference optimizer.py:220-226, :248-252: first-owner rule for the bond ends, every listed atom
    # moves, centroid over the listed ids.  Through the public API (the lowering marks the topology general).
    g = load_golden(case)
    opt = golden_optimizer(g)
    mol = golden_molecule(g)
    out_mol, result = opt.get_trajectory(mol, golden_bonds(g), golden_subunits(g))
    props = [p for _, p in result.get_steps_properties()]
    passed = np.array([-1 if p["passed"] is None else int(p["passed"]) for p in props])
    assert np.array_equal(passed, g["passed"])
    assert rel_close([p["system_potential"] for p in props], g["system_potential"], RTOL64)
    assert rel_close([p["nonbonded_potential"] for p in props], g["nonbonded_potential"], RTOL64)
    assert rel_close([p["max_bond_distance"] for p in props], g["max_bond_distance"], RTOL64)
    assert np.allclose(out_mol.get_position_matrix(), g["final_positions"], rtol=RTOL64, atol=1e-9)
    if "positions" in g.files:
        frames = np.array([f for _, f in result.get_trajectory()])
        assert np.allclose(frames, g["positions"], rtol=RTOL64, atol=1e-9)
    # same chain through get_result (no trace buffers) and through the batched entry
    mol2, last = golden_optimizer(g).get_result(mol, golden_bonds(g), golden_subunits(g))
    assert np.allclose(mol2.get_position_matrix(), g["final_positions"], rtol=RTOL64, atol=1e-9)
    assert rel_close(last.system_potential, g["system_potential"][-1], RTOL64)
    batch = golden_optimizer(g).get_results_batch(mol, golden_bonds(g), golden_subunits(g), seeds=[int(g["seed"])] * 3)
    assert np.all(batch.n_accepted == int((g["passed"] == 1).sum()))


@pytest.mark.parametrize("cluster", [1, 2, 4, 8])
@pytest.mark.parametrize("case", ["poc_graph_seed7", "poc_bb_seed1000", "benzene_seed1000"])
def test_split_state_kernel_on_reference_goldens(case, cluster):
    # partition topologies forced through the split-state kernel, state over 1 / 2 / 4 / 8 CTAs
    # (benzene on 8 CTAs: most CTAs own no atom at all)
    g = load_golden(case)
    opt = golden_optimizer(g)
    out = run_split(g["pos0"], golden_bonds(g), golden_subunits(g), opt._params(), int(g["num_steps"]),
                    int(g["seed"]), cluster=cluster)
    from mchammer_b200.batch import unpack_info
    _, passed, _, _ = unpack_info(out.trace_info[0].cpu().numpy())
    assert np.array_equal(passed, g["passed"])
    trace = out.trace_potentials[0].cpu().numpy()
    assert rel_close(trace[:, 0], g["system_potential"], RTOL64)
    assert rel_close(trace[:, 1], g["nonbonded_potential"], RTOL64)
    assert rel_close(trace[:, 2], g["max_bond_distance"], RTOL64)
    assert np.allclose(out.positions[0].cpu().numpy(), g["final_positions"], rtol=RTOL64, atol=1e-9)
    if "positions" in g.files:
        assert np.allclose(out.trajectory[0].cpu().numpy(), g["positions"], rtol=RTOL64, atol=1e-9)


@pytest.mark.parametrize("precision,cluster", [("fp64", 8), ("fp64", 2), ("fp32", 4), ("fp32", 16)])
def test_split_state_kernel_on_the_bench_cage(precision, cluster):
    mol, long_bonds, subunits = synthetic.cube_cage()
    steps = 60
    p = orc.OptimizerParams(step_size=0.25, target_bond_length=1.2, num_steps=steps)
    want = orc.optimizer_run(mol.get_position_matrix(), long_bonds, subunits, p, np.random.default_rng(77))
    opt = mch.Optimizer(0.25, 1.2, steps, precision=precision)
    out = run_split(mol.get_position_matrix(), long_bonds, subunits, opt._params(), steps, 77, precision, cluster)
    from mchammer_b200.batch import unpack_info
    _, passed, _, _ = unpack_info(out.trace_info[0].cpu().numpy())
    assert np.array_equal(passed, want.passed)
    rtol = RTOL64 if precision == "fp64" else RTOL32
    trace = out.trace_potentials[0].cpu().numpy()
    assert rel_close(trace[:, 0], want.system_potential, rtol)
    assert rel_close(trace[:, 1], want.nonbonded_potential, rtol)
    assert np.allclose(out.positions[0].cpu().numpy(), want.final_positions, rtol=0,
                       atol=1e-9 if precision == "fp64" else 1e-3)


@pytest.mark.parametrize("seed", range(10))
def test_random_overlapping_subunits_against_oracle(seed):
    # seeded fuzz: random clusters, every subunit additionally lists a few atoms of other subunits (and
    # sometimes one of its own twice), some atoms in no subunit, random bond ends, random cluster size
    rng = np.random.default_rng(2600 + seed)
    n_sub = int(rng.integers(2, 8))
    sizes = [int(rng.integers(1, 60)) for _ in range(n_sub)]
    centres = rng.normal(scale=9.0, size=(n_sub, 3))
    pos = np.concatenate([c + rng.normal(scale=1.6, size=(k, 3)) for c, k in zip(centres, sizes)]
                         + [rng.normal(scale=8.0, size=(int(rng.integers(0, 4)), 3))])
    n = pos.shape[0]
    n_owned = sum(sizes)
    order = rng.permutation(n)
    inv = np.argsort(order)
    pos = pos[order]
    start = np.concatenate([[0], np.cumsum(sizes)])
    subunits = {}
    for k in rng.permutation(n_sub):
        ids = [int(inv[i]) for i in range(start[k], start[k + 1])]
        extra = rng.choice(n_owned, size=int(rng.integers(0, 4)), replace=False)
        ids += [int(inv[i]) for i in extra if int(inv[i]) not in ids]
        if rng.random() < 0.3:
            ids.append(ids[0])  # listed twice
        rng.shuffle(ids)
        subunits[int(5 * k + 1)] = ids
    owned = sorted({a for ids in subunits.values() for a in ids})
    bonds = tuple((int(a), int(b)) for a, b in (rng.choice(owned, size=2, replace=False)
                                                for _ in range(int(rng.integers(1, 7)))))
    steps = 50
    params = dict(step_size=float(rng.uniform(0.1, 0.4)), target_bond_length=float(rng.uniform(1.0, 2.0)),
                  bond_epsilon=float(rng.uniform(10, 60)), nonbond_epsilon=float(rng.uniform(5, 30)),
                  nonbond_sigma=float(rng.uniform(1.0, 1.6)), nonbond_mu=float(rng.choice([2.0, 3.0, 3.0, 4.5, 6.0])),
                  beta=float(rng.uniform(0.5, 3.0)))
    want = orc.optimizer_run(pos, bonds, subunits, orc.OptimizerParams(num_steps=steps, **params),
                             np.random.default_rng(seed))
    opt = mch.Optimizer(num_steps=steps, random_seed=seed, **params)
    out = run_split(pos, bonds, subunits, opt._params(), steps, seed, cluster=int(rng.choice([1, 2, 4])),
                    threads=int(rng.choice([0, 32, 96])))
    from mchammer_b200.batch import unpack_info
    bond, passed, kind, moved = unpack_info(out.trace_info[0].cpu().numpy())
    assert np.array_equal(passed, want.passed)
    assert np.array_equal(bond, want.bond_index)
    assert np.array_equal(moved, want.moving_subunit)
    trace = out.trace_potentials[0].cpu().numpy()
    assert rel_close(trace[:, 0], want.system_potential, RTOL64)
    assert rel_close(trace[:, 2], want.max_bond_distance, RTOL64)
    assert np.allclose(out.positions[0].cpu().numpy(), want.final_positions, rtol=RTOL64, atol=1e-9)


def big_blob_molecule(n_subunits: int, side: int, seed: int):
    """n_subunits rigid blobs of side^3 atoms (jittered 1.5 A lattice) on a jittered 3-D grid; bonds
    between lattice neighbours.  No chemistry implied; no two atoms closer than ~1 A."""
    rng = np.random.default_rng(seed)
    per = side ** 3
    grid = int(np.ceil(n_subunits ** (1 / 3)))
    cell = np.stack(np.meshgrid(*[np.arange(side)] * 3, indexing="ij"), axis=-1).reshape(-1, 3) * 1.5
    cell = cell - cell.mean(axis=0)
    spacing = side * 1.5 + 4.0
    coords, subunits, centres = [], {}, []
    for s in range(n_subunits):
        gx, gy, gz = s % grid, (s // grid) % grid, s // (grid * grid)
        centre = np.array([gx, gy, gz]) * spacing + rng.normal(scale=0.4, size=3)
        centres.append((gx, gy, gz))
        coords.append(centre + cell + rng.normal(scale=0.08, size=cell.shape))
        subunits[s] = list(range(s * per, (s + 1) * per))
    pos = np.concatenate(coords)
    index = {c: s for s, c in enumerate(centres)}
    bonds = []
    for s, (gx, gy, gz) in enumerate(centres):
        for d in ((1, 0, 0), (0, 1, 0), (0, 0, 1)):
            t = index.get((gx + d[0], gy + d[1], gz + d[2]))
            if t is not None:
                a = s * per + int(np.argmin(np.linalg.norm(coords[s] - coords[t].mean(axis=0), axis=1)))
                b = t * per + int(np.argmin(np.linalg.norm(coords[t] - coords[s].mean(axis=0), axis=1)))
                bonds.append((a, b))
    return pos, tuple(bonds), subunits


def test_molecule_beyond_one_sm_20000_atoms_300_subunits():
    # 300 subunits x 64 atoms = 19,200 atoms + 800 atoms in no subunit = 20,000 atoms: the one-CTA
    # kernel holds 6,744 atoms (fp64) and ~150 subunits; the public API must route this to the
    # split-state kernel by itself and reproduce the oracle's chain.
    pos, bonds, subunits = big_blob_molecule(300, 4, seed=12)
    rng = np.random.default_rng(3)
    loose = rng.normal(scale=6.0, size=(800, 3)) + pos.mean(axis=0) + np.array([0.0, 0.0, -40.0])
    pos = np.concatenate([pos, loose])
    assert pos.shape[0] == 20000 and len(subunits) == 300
    mol = mch.Molecule(atoms=[mch.Atom(i, "C") for i in range(pos.shape[0])], bonds=[], position_matrix=pos)
    steps = 16
    p = orc.OptimizerParams(step_size=0.25, target_bond_length=1.2, num_steps=steps)
    want = orc.optimizer_run(pos, bonds, subunits, p, np.random.default_rng(5))
    assert 0 < int((want.passed == 1).sum()) < steps - 1  # both outcomes occur
    opt = mch.Optimizer(0.25, 1.2, steps, random_seed=5)
    out_mol, result = opt.get_trajectory(mol, bonds, subunits)
    props = [q for _, q in result.get_steps_properties()]
    passed = np.array([-1 if q["passed"] is None else int(q["passed"]) for q in props])
    assert np.array_equal(passed, want.passed)
    assert rel_close([q["system_potential"] for q in props], want.system_potential, RTOL64)
    assert rel_close([q["nonbonded_potential"] for q in props], want.nonbonded_potential, RTOL64)
    assert np.allclose(out_mol.get_position_matrix(), want.final_positions, rtol=RTOL64, atol=1e-9)
    # fast mode on the same molecule: potentials within 1e-4 while the accept sequence agrees
    fast = mch.Optimizer(0.25, 1.2, steps, random_seed=5, precision="fp32")
    traj = fast.get_trajectories_batch(mol, bonds, subunits, seeds=[5], want_positions=False)
    same = traj.passed[0] == want.passed
    prefix = len(same) if same.all() else int(np.argmin(same))
    assert prefix >= 8
    assert rel_close(traj.system_potential[0][:prefix], want.system_potential[:prefix], RTOL32)


def test_too_large_for_a_cluster_is_still_refused():
    mol, long_bonds, subunits = synthetic.cube_cage(atoms_per_subunit=4, seed=1, spread=4.0)
    assert native.lib().mch_optimize_smem_bytes(400_000, 20, 24, 20_000, native.MCH_F64) == -2
