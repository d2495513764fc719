"""GPU parity tests: the CUDA path (through the C ABI) against
  (a) golden vectors produced by the unmodified reference (tests/golden/*.npz),
  (b) the CPU oracle on the same seeded inputs,
  (c) size-independent properties at BASELINE sizes.

Tolerances (BASELINE.json north_star): fp64 mode reproduces the reference's accept/reject
sequence exactly, potentials and positions within 1e-9 relative; fp32 mode keeps
potentials within 1e-4 relative.
"""

from __future__ import annotations

import numpy as np
import pytest

import mchammer_b200 as mch
from mchammer_b200 import synthetic
from oracle import mch_oracle as orc
from tests.conftest import golden_bonds, golden_subunits, load_golden

pytestmark = pytest.mark.gpu

RTOL64 = 1e-9
RTOL32 = 1e-4


def golden_molecule(g) -> mch.Molecule:
    return mch.Molecule(
        atoms=tuple(mch.Atom(id=i, element_string=str(e)) for i, e in enumerate(g["elements"])),
        bonds=tuple(mch.Bond(id=i, atom_ids=tuple(int(x) for x in b)) for i, b in enumerate(g["bonds"])),
        position_matrix=g["pos0"],
    )


def _num(value):
    """Golden scalars keep the Python type the reference was given (50 prints as '50', not '50.0')."""
    value = np.asarray(value)
    return int(value) if np.issubdtype(value.dtype, np.integer) else float(value)


def golden_optimizer(g, **kw) -> mch.Optimizer:
    return mch.Optimizer(
        step_size=_num(g["step_size"]), target_bond_length=_num(g["target_bond_length"]),
        num_steps=int(g["num_steps"]), bond_epsilon=_num(g["bond_epsilon"]),
        nonbond_epsilon=_num(g["nonbond_epsilon"]), nonbond_sigma=_num(g["nonbond_sigma"]),
        nonbond_mu=_num(g["nonbond_mu"]), beta=_num(g["beta"]), random_seed=int(g["seed"]), **kw,
    )


def rel_close(a, b, rtol):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.all(np.abs(a - b) <= rtol * np.maximum(np.abs(b), 1e-300))


# ------------------------------------------------------------------------- potentials
def test_known_answer_potentials_fp64():
    # reference tests/conftest.py:145-152, tests/optimizer/test_optimizer.py:37-62
    opt = mch.Optimizer(step_size=0.1, target_bond_length=2.0, num_steps=100)
    pos = np.array([[0, 1, 0], [1, 1, 0], [-1, 1, 0], [0, 10, 0], [1, 10, 0], [-1, 10, 0]])
    assert np.isclose(opt._compute_nonbonded_potential(pos), 147.2965949864993, rtol=1e-13, atol=0)
    mol = mch.Molecule(
        atoms=[mch.Atom(i, "C") for i in range(6)],
        bonds=[mch.Bond(0, (0, 1)), mch.Bond(1, (0, 2)), mch.Bond(2, (0, 3)), mch.Bond(3, (3, 4)), mch.Bond(4, (3, 5))],
        position_matrix=pos,
    )
    u, unb = opt.compute_potential(mol, bond_pair_ids=((0, 3),))
    assert np.isclose(unb, 147.2965949864993, rtol=1e-13, atol=0)
    assert np.isclose(u, 2597.2965949864993, rtol=1e-13, atol=0)


@pytest.mark.parametrize("precision,rtol", [("fp64", 1e-12), ("fp32", 1e-5)])
def test_potentials_random_geometries(precision, rtol):
    g = load_golden("potentials")
    opt = mch.Optimizer(step_size=0.1, target_bond_length=2.0, num_steps=2, precision=precision)
    for n in (2, 3, 17, 112, 300):
        mol = mch.Molecule(atoms=[mch.Atom(i, "C") for i in range(n)], bonds=[], position_matrix=g[f"pos_{n}"])
        u, unb = opt.compute_potential(mol, g[f"bonds_{n}"].tolist())
        assert np.isclose(u, float(g[f"u_{n}"]), rtol=rtol, atol=0), (n, u, float(g[f"u_{n}"]))
        assert np.isclose(unb, float(g[f"unb_{n}"]), rtol=rtol, atol=0)


@pytest.mark.parametrize("mu", [2.0, 2.5, 3.0, 6.0, 12.0])
@pytest.mark.parametrize("precision,rtol", [("fp64", 1e-11), ("fp32", 2e-5)])
def test_potential_mu_variants_against_oracle(mu, precision, rtol):
    rng = np.random.default_rng(5)
    pos = rng.normal(scale=4.0, size=(150, 3))
    bonds = [(0, 9), (3, 77), (140, 2)]
    p = orc.OptimizerParams(0.25, 1.2, 2, nonbond_mu=mu, nonbond_sigma=1.4, nonbond_epsilon=7.0)
    want_u, want_unb = orc.compute_potential(pos, bonds, p)
    opt = mch.Optimizer(0.25, 1.2, 2, nonbond_mu=mu, nonbond_sigma=1.4, nonbond_epsilon=7.0, precision=precision)
    mol = mch.Molecule(atoms=[mch.Atom(i, "C") for i in range(150)], bonds=[], position_matrix=pos)
    u, unb = opt.compute_potential(mol, bonds)
    assert np.isclose(unb, want_unb, rtol=rtol, atol=0)
    assert np.isclose(u, want_u, rtol=rtol, atol=0)


# ------------------------------------------------------------------------- MC traces vs reference
TRACE_CASES = ["benzene_seed1000", "benzene_seed7", "benzene_mu2p5", "benzene_mu6", "toy_seed1000",
               "poc_graph_seed1000", "poc_graph_seed7", "poc_graph_seed42", "poc_bb_seed1000"]


@pytest.mark.parametrize("case", TRACE_CASES)
def test_fp64_trajectory_matches_reference(case):
    g = load_golden(case)
    mol = golden_molecule(g)
    opt = golden_optimizer(g)
    out_mol, result = opt.get_trajectory(mol, golden_bonds(g), golden_subunits(g))
    props = list(result.get_steps_properties())
    assert len(props) == int(g["num_steps"])
    assert result.get_step_count() == int(g["num_steps"]) - 1
    passed = np.array([-1 if p["passed"] is None else int(p["passed"]) for _, p in props])
    assert np.array_equal(passed, g["passed"]), "accept/reject sequence differs from the reference"
    assert rel_close([p["system_potential"] for _, p in props], g["system_potential"], RTOL64)
    assert rel_close([p["nonbonded_potential"] for _, p in props], g["nonbonded_potential"], RTOL64)
    assert rel_close([p["max_bond_distance"] for _, p in props], g["max_bond_distance"], RTOL64)
    assert np.allclose(out_mol.get_position_matrix(), g["final_positions"], rtol=RTOL64, atol=1e-9)
    assert result.get_number_passed() == int(g["number_passed"])
    if "positions" in g.files:
        frames = np.array([f for _, f in result.get_trajectory()])
        assert np.allclose(frames, g["positions"], rtol=RTOL64, atol=1e-9)
    # log text: same lines as the reference (numbers compared numerically, words exactly)
    got_lines = result.get_log().splitlines()
    want_lines = str(g["log"]).splitlines()
    assert len(got_lines) == len(want_lines)
    for got, want in zip(got_lines, want_lines):
        if want.startswith("Total optimisation time"):
            assert got.startswith("Total optimisation time: ")
            continue
        gt, wt = got.split(), want.split()
        if len(wt) >= 6 and wt[0].isdigit():
            assert gt[0] == wt[0] and gt[4:] == wt[4:], (got, want)
            assert rel_close([float(x) for x in gt[1:4]], [float(x) for x in wt[1:4]], RTOL64)
        else:
            assert got == want


@pytest.mark.parametrize("case", ["benzene_seed1000", "poc_graph_seed1000", "poc_bb_seed1000"])
def test_fp64_get_result_matches_reference(case):
    g = load_golden(case)
    mol = golden_molecule(g)
    before = mol.get_position_matrix()
    out_mol, last = golden_optimizer(g).get_result(mol, golden_bonds(g), golden_subunits(g))
    assert np.array_equal(mol.get_position_matrix(), before), "input molecule was mutated"
    assert last.step == int(g["num_steps"]) - 1
    assert last.passed == bool(g["passed"][-1])
    assert rel_close(last.system_potential, g["system_potential"][-1], RTOL64)
    assert rel_close(last.nonbonded_potential, g["nonbonded_potential"][-1], RTOL64)
    assert rel_close(last.max_bond_distance, g["max_bond_distance"][-1], RTOL64)
    assert np.allclose(last.position_matrix, g["final_positions"], rtol=RTOL64, atol=1e-9)
    assert np.allclose(out_mol.get_position_matrix(), g["final_positions"], rtol=RTOL64, atol=1e-9)


def test_committed_example_molfiles():
    # reference examples/poc_opt.mol etc. hold 4 decimals
    ex = load_golden("reference_examples")
    for case, name in (("benzene_seed1000", "benzene_opt"), ("poc_graph_seed1000", "poc_opt"),
                       ("poc_bb_seed1000", "poc_opt_nci")):
        g = load_golden(case)
        out_mol, _ = golden_optimizer(g).get_result(golden_molecule(g), golden_bonds(g), golden_subunits(g))
        assert np.abs(out_mol.get_position_matrix() - ex[name]).max() < 1.5e-4


def test_generator_advances_across_calls():
    g = load_golden("benzene_two_calls")
    b = load_golden("benzene_seed1000")
    mol = golden_molecule(b)
    lb = golden_bonds(b)
    su = mol.get_subunits(bond_pair_ids=lb)
    opt = mch.Optimizer(step_size=0.25, target_bond_length=1.2, num_steps=40, random_seed=3)
    m1, r1 = opt.get_result(mol, lb, su)
    m2, r2 = opt.get_result(m1, lb, su)
    assert np.allclose(m1.get_position_matrix(), g["first_positions"], rtol=RTOL64, atol=1e-9)
    assert np.allclose(m2.get_position_matrix(), g["second_positions"], rtol=RTOL64, atol=1e-9)
    assert rel_close(r1.system_potential, float(g["first_u"]), RTOL64)
    assert rel_close(r2.system_potential, float(g["second_u"]), RTOL64)
    assert int(bool(r1.passed)) == int(g["first_passed"]) and int(bool(r2.passed)) == int(g["second_passed"])


def test_num_steps_one_and_zero():
    g = load_golden("benzene_seed1000")
    mol = golden_molecule(g)
    for steps in (1, 0):
        opt = mch.Optimizer(step_size=0.25, target_bond_length=1.2, num_steps=steps)
        out_mol, last = opt.get_result(mol, golden_bonds(g), golden_subunits(g))
        assert last.step == 0 and last.passed is None
        assert last.log.endswith("-- --\n")
        assert rel_close(last.system_potential, g["system_potential"][0], 1e-12)
        assert np.array_equal(out_mol.get_position_matrix(), g["pos0"])


# ------------------------------------------------------------------------- reference-style integration tests
def test_reference_style_toy_run():
    # mirrors reference tests/optimizer/test_optimizer.py:101-177
    pos = np.array([[0, 1, 0], [1, 1, 0], [-1, 1, 0], [0, 10, 0], [1, 10, 0], [-1, 10, 0]])
    bonds = [mch.Bond(0, (0, 1)), mch.Bond(1, (0, 2)), mch.Bond(2, (0, 3)), mch.Bond(3, (3, 4)), mch.Bond(4, (3, 5))]
    molecule = mch.Molecule(atoms=[mch.Atom(i, "C") for i in range(6)], bonds=bonds, position_matrix=pos)
    optimizer = mch.Optimizer(step_size=0.1, target_bond_length=2.0, num_steps=100)
    subunits = molecule.get_subunits(bond_pair_ids=((0, 3),))
    original = molecule.get_position_matrix()
    mol, results = optimizer.get_trajectory(mol=molecule, bond_pair_ids=((0, 3),), subunits=subunits)
    assert results.get_step_count() == 99
    assert len(tuple(results.get_steps_properties())) == 100
    final = results.get_final_position_matrix()
    length = np.linalg.norm(mch.get_bond_vector(position_matrix=final, bond_pair=(0, 3)))
    assert 1.5 < length < 2.5
    for bond in molecule.get_bonds():
        pair = (bond.get_atom1_id(), bond.get_atom2_id())
        if pair != (0, 3):
            assert np.isclose(mch.get_atom_distance(final, *pair), mch.get_atom_distance(original, *pair), atol=1e-6)


# ------------------------------------------------------------------------- batched chains vs oracle
@pytest.mark.parametrize("precision", ["fp64"])
def test_batch_matches_oracle_on_synthetic_cage(precision):
    mol, long_bonds, subunits = synthetic.cube_cage(atoms_per_subunit=10, seed=4, spread=6.0)
    seeds = [1000 + c for c in range(6)]
    steps = 80
    opt = mch.Optimizer(step_size=0.25, target_bond_length=1.2, num_steps=steps, precision=precision)
    got = opt.get_results_batch(mol, long_bonds, subunits, seeds=seeds)
    p = orc.OptimizerParams(step_size=0.25, target_bond_length=1.2, num_steps=steps)
    for c, seed in enumerate(seeds):
        want = orc.optimizer_run(mol.get_position_matrix(), long_bonds, subunits, p, np.random.default_rng(seed))
        assert int(got.n_accepted[c]) == int((want.passed == 1).sum())
        assert int(got.last_passed[c]) == int(want.passed[-1])
        assert int(got.last_bond_index[c]) == int(want.bond_index[-1])
        assert rel_close(got.system_potential[c], want.system_potential[-1], RTOL64)
        assert rel_close(got.nonbonded_potential[c], want.nonbonded_potential[-1], RTOL64)
        assert rel_close(got.max_bond_distance[c], want.max_bond_distance[-1], RTOL64)
        assert np.allclose(got.positions[c], want.final_positions, rtol=RTOL64, atol=1e-9)


def test_batch_with_per_chain_positions_and_ragged_subunits():
    # ragged subunits (1, 3, 17, 40 atoms), an atom in no subunit, subunit dict order != atom order
    rng = np.random.default_rng(8)
    sizes = [17, 1, 40, 3]
    centres = np.array([[0, 0, 0], [9, 0, 0], [0, 11, 0], [0, 0, 8.0]])
    coords, subunits, offset = [], {}, 0
    for k, (size, centre) in enumerate(zip(sizes, centres)):
        coords.append(centre + rng.normal(scale=1.8, size=(size, 3)))
        subunits[10 - k] = list(range(offset, offset + size))[::-1]
        offset += size
    coords.append(np.array([[5.0, 5.0, 5.0]]))  # loose atom: static, still interacts
    pos = np.concatenate(coords)
    n = pos.shape[0]
    mol = mch.Molecule(atoms=[mch.Atom(i, "C") for i in range(n)], bonds=[], position_matrix=pos)
    long_bonds = ((16, 17), (57, 20), (0, 58), (18, 60))
    steps = 60
    chains = 3
    positions = np.stack([pos + rng.normal(scale=0.05, size=pos.shape) for _ in range(chains)])
    seeds = [5, 6, 7]
    opt = mch.Optimizer(step_size=0.3, target_bond_length=1.5, num_steps=steps)
    got = opt.get_results_batch(mol, long_bonds, subunits, seeds=seeds, positions=positions)
    p = orc.OptimizerParams(step_size=0.3, target_bond_length=1.5, num_steps=steps)
    for c in range(chains):
        want = orc.optimizer_run(positions[c], long_bonds, subunits, p, np.random.default_rng(seeds[c]))
        assert int(got.n_accepted[c]) == int((want.passed == 1).sum())
        assert rel_close(got.system_potential[c], want.system_potential[-1], RTOL64)
        assert np.allclose(got.positions[c], want.final_positions, rtol=RTOL64, atol=1e-9)


@pytest.mark.parametrize("seed", range(12))
def test_random_molecules_against_oracle(seed):
    # seeded fuzz: random sizes, ragged random partitions (some atoms in no subunit), random
    # bond lists (intra- and cross-subunit, repeated atoms), parameters and exponents -- fp64
    # mode must give the oracle's accept/reject sequence, potentials and geometry
    rng = np.random.default_rng(1234 + seed)
    n = int(rng.integers(2, 260))
    n_sub = int(rng.integers(1, min(n, 24) + 1))
    labels = rng.integers(0, n_sub, size=n)
    loose = rng.random(n) < (0.1 if seed % 3 == 0 else 0.0)
    members = {k: [int(i) for i in np.flatnonzero((labels == k) & ~loose)] for k in range(n_sub)}
    subunits = {int(100 - k): (v if k % 2 else v[::-1]) for k, v in members.items() if v}
    owned = sorted(i for v in subunits.values() for i in v)
    if len(owned) < 2:
        pytest.skip("degenerate draw")
    # well separated random points: jittered grid, so no pair is closer than ~0.6
    side = int(np.ceil(n ** (1 / 3)))
    grid = np.array([(x, y, z) for x in range(side) for y in range(side) for z in range(side)], dtype=float)
    pos = 1.7 * grid[rng.permutation(len(grid))[:n]] + rng.uniform(-0.4, 0.4, size=(n, 3)) + rng.normal(scale=30.0, size=3)
    n_bonds = int(rng.integers(1, 9))
    bonds = []
    while len(bonds) < n_bonds:
        a, b = (int(x) for x in rng.choice(owned, size=2, replace=False))
        bonds.append((a, b) if rng.random() < 0.5 else [b, a])
    mu = float(rng.choice([3.0, 3.0, 2.0, 4.0, 2.5, 6.0]))
    steps = int(rng.integers(2, 90))
    kwargs = dict(step_size=float(rng.uniform(0.05, 0.5)), target_bond_length=float(rng.uniform(0.8, 3.0)),
                  num_steps=steps, bond_epsilon=float(rng.uniform(1, 80)), nonbond_epsilon=float(rng.uniform(1, 40)),
                  nonbond_sigma=float(rng.uniform(0.8, 1.6)), nonbond_mu=mu, beta=float(rng.uniform(0.2, 4.0)))
    mol = mch.Molecule(atoms=[mch.Atom(i, "C") for i in range(n)], bonds=[], position_matrix=pos)
    opt = mch.Optimizer(random_seed=77 + seed, **kwargs)
    out_mol, result = opt.get_trajectory(mol, bonds, subunits)
    want = orc.optimizer_run(pos, bonds, subunits, orc.OptimizerParams(**kwargs), np.random.default_rng(77 + seed))
    props = list(result.get_steps_properties())
    passed = np.array([-1 if p["passed"] is None else int(p["passed"]) for _, p in props])
    assert np.array_equal(passed, want.passed), "accept/reject sequence differs from the oracle"
    assert rel_close([p["system_potential"] for _, p in props], want.system_potential, RTOL64)
    assert rel_close([p["nonbonded_potential"] for _, p in props], want.nonbonded_potential, RTOL64)
    assert rel_close([p["max_bond_distance"] for _, p in props], want.max_bond_distance, RTOL64)
    assert np.allclose(out_mol.get_position_matrix(), want.final_positions, rtol=RTOL64, atol=1e-9)
    # the same chain through the batched entry point in fp32: potential within the stated tolerance
    # as long as the two precisions took the same decisions
    fast = mch.Optimizer(precision="fp32", **kwargs).get_results_batch(mol, bonds, subunits, seeds=[77 + seed])
    if int(fast.n_accepted[0]) == int((want.passed == 1).sum()):
        assert rel_close(fast.system_potential[0], want.system_potential[-1], RTOL32)


def test_fp32_mode_tolerance_against_reference():
    # fp32 fast mode: potentials within 1e-4 relative (north_star).  Trajectories may fork
    # once an accept test lands within fp32 noise, so compare the common prefix of the
    # accept sequence and require that prefix to be long.
    for case in ("benzene_seed1000", "poc_graph_seed1000"):
        g = load_golden(case)
        opt = golden_optimizer(g, precision="fp32")
        _, result = opt.get_trajectory(golden_molecule(g), golden_bonds(g), golden_subunits(g))
        props = [p for _, p in result.get_steps_properties()]
        passed = np.array([-1 if p["passed"] is None else int(p["passed"]) for p in props])
        same = passed == g["passed"]
        prefix = len(same) if same.all() else int(np.argmin(same))
        assert prefix >= min(100, len(same)), f"{case}: fp32 accept sequence forked at step {prefix}"
        u = np.array([p["system_potential"] for p in props])[:prefix]
        unb = np.array([p["nonbonded_potential"] for p in props])[:prefix]
        assert rel_close(u, g["system_potential"][:prefix], RTOL32)
        assert rel_close(unb, g["nonbonded_potential"][:prefix], RTOL32)


def test_fp32_expanded_form_energy_accuracy():
    # fp32 fast mode forms r^2 in the expanded form around the moving subunit's centroid
    # (csrc/mch_device.cuh); step 0 of the step kernel sums every pair that way.  Stated
    # tolerance 1e-4 relative; required here: 1e-5, also far from the origin and on the
    # 5004-atom cage (radius ~45 A).
    exact = mch.Optimizer(0.25, 1.2, 1, precision="fp64")
    fast = mch.Optimizer(0.25, 1.2, 1, precision="fp32")
    cases = [synthetic.cube_cage(), synthetic.cuboctahedron_cage()]
    mol, lb, su = cases[0]
    cases.append((mol.with_displacement(np.array([1234.5, -987.25, 4321.0])), lb, su))
    for mol, long_bonds, subunits in cases:
        want_u, want_unb = exact.compute_potential(mol, long_bonds)
        got = fast.get_results_batch(mol, long_bonds, subunits, seeds=[1])
        assert np.isclose(got.nonbonded_potential[0], want_unb, rtol=1e-5, atol=0)
        assert np.isclose(got.system_potential[0], want_u, rtol=1e-5, atol=0)
        ref64 = exact.get_results_batch(mol, long_bonds, subunits, seeds=[1])
        assert np.isclose(ref64.nonbonded_potential[0], want_unb, rtol=1e-12, atol=0)


@pytest.mark.parametrize("case", ["benzene_mu2p5", "benzene_mu6"])
def test_fp32_mode_other_exponents(case):
    # fp32 pair terms for real mu (exp2/log2) and integer mu != 3 (square-and-multiply)
    g = load_golden(case)
    opt = golden_optimizer(g, precision="fp32")
    _, result = opt.get_trajectory(golden_molecule(g), golden_bonds(g), golden_subunits(g))
    props = [p for _, p in result.get_steps_properties()]
    passed = np.array([-1 if p["passed"] is None else int(p["passed"]) for p in props])
    same = passed == g["passed"]
    prefix = len(same) if same.all() else int(np.argmin(same))
    assert prefix >= 20
    u = np.array([p["system_potential"] for p in props])[:prefix]
    assert rel_close(u, g["system_potential"][:prefix], RTOL32)


def test_batch_trajectories_match_reference_and_single_chain():
    # chain 0 of a batched trajectory = the reference's seeded get_trajectory (golden, seed 1000);
    # chain 1 (seed 7) = the other golden of the same molecule; streamed in slices of ONE chain
    g, g7 = load_golden("poc_graph_seed1000"), load_golden("poc_graph_seed7")
    mol, bonds, subunits = golden_molecule(g), golden_bonds(g), golden_subunits(g)
    opt = golden_optimizer(g)
    n, steps = mol.get_num_atoms(), int(g["num_steps"])
    frame_bytes = steps * n * 3 * 8
    traj = opt.get_trajectories_batch(mol, bonds, subunits, seeds=[1000, 7, 1000],
                                      slice_bytes=frame_bytes, devices=[0, 0])
    assert len(traj) == 3 and traj.positions.shape == (3, steps, n, 3)
    for c, want in ((0, g), (1, g7), (2, g)):
        assert np.array_equal(traj.passed[c], want["passed"]), "accept/reject sequence differs from the reference"
        assert rel_close(traj.system_potential[c], want["system_potential"], RTOL64)
        assert rel_close(traj.nonbonded_potential[c], want["nonbonded_potential"], RTOL64)
        assert rel_close(traj.max_bond_distance[c], want["max_bond_distance"], RTOL64)
        assert np.allclose(traj.final_positions[c], want["final_positions"], rtol=RTOL64, atol=1e-9)
        assert np.allclose(traj.positions[c, -1], want["final_positions"], rtol=RTOL64, atol=1e-9)
        if "positions" in want.files:
            assert np.allclose(traj.positions[c], want["positions"], rtol=RTOL64, atol=1e-9)
    assert np.array_equal(traj.positions[0], traj.positions[2])
    # the rebuilt Result carries the reference's log text
    result = opt.trajectory_result(traj, 0, bonds, subunits)
    assert result.get_step_count() == steps - 1
    assert result.get_number_passed() == int(g["number_passed"])
    got_lines, want_lines = result.get_log().splitlines(), str(g["log"]).splitlines()
    assert len(got_lines) == len(want_lines)
    for got, want in zip(got_lines, want_lines):
        gt, wt = got.split(), want.split()
        if len(wt) >= 6 and wt[0].isdigit():
            assert gt[0] == wt[0] and gt[4:] == wt[4:], (got, want)
    # without frames: same traces, nothing bulky comes back
    lean = opt.get_trajectories_batch(mol, bonds, subunits, seeds=[1000, 7], want_positions=False)
    assert lean.positions is None
    assert np.array_equal(lean.system_potential, traj.system_potential[:2])
    assert lean.d2h_bytes < traj.d2h_bytes / 50


def test_batch_io_float32_and_two_shards_on_one_device():
    mol, long_bonds, subunits = synthetic.cube_cage(atoms_per_subunit=10, seed=4, spread=6.0)
    seeds = list(range(40, 47))
    opt = mch.Optimizer(step_size=0.25, target_bond_length=1.2, num_steps=50)
    whole = opt.get_results_batch(mol, long_bonds, subunits, seeds=seeds)
    # the same chains split over two shards (both on device 0): identical results, host gather
    split = opt.get_results_batch(mol, long_bonds, subunits, seeds=seeds, devices=[0, 0])
    assert np.array_equal(split.positions, whole.positions)
    assert np.array_equal(split.system_potential, whole.system_potential)
    assert np.array_equal(split.n_accepted, whole.n_accepted)
    # float32 I/O with fp64 compute: inputs/outputs rounded to fp32, arithmetic unchanged
    f32 = opt.get_results_batch(mol, long_bonds, subunits, seeds=seeds, io_dtype=np.float32)
    assert f32.positions.dtype == np.float32
    start32 = mol.get_position_matrix().astype(np.float32).astype(np.float64)
    p = orc.OptimizerParams(step_size=0.25, target_bond_length=1.2, num_steps=50)
    want = orc.optimizer_run(start32, long_bonds, subunits, p, np.random.default_rng(seeds[0]))
    assert rel_close(f32.system_potential[0], want.system_potential[-1], RTOL64)
    assert np.allclose(f32.positions[0], want.final_positions, rtol=0, atol=2e-6)
    # a prepared plan can be re-run; slices of the pipelined run do not change results
    plan = opt.prepare_batch(mol, long_bonds, subunits, len(seeds), n_slices=3)
    again = plan.run(seeds=seeds).copy()
    assert np.array_equal(again.positions, whole.positions)


def test_single_subunit_and_intra_subunit_bond():
    # every atom in one subunit: no cross pairs; moves translate the whole molecule rigidly,
    # so the potential never changes -- the chain must still follow the reference exactly
    rng = np.random.default_rng(12)
    pos = rng.normal(scale=2.5, size=(9, 3))
    mol = mch.Molecule(atoms=[mch.Atom(i, "C") for i in range(9)], bonds=[], position_matrix=pos)
    subunits = {0: list(range(9))}
    long_bonds = ((0, 5), (3, 8))
    opt = mch.Optimizer(step_size=0.2, target_bond_length=1.5, num_steps=30, random_seed=77)
    out_mol, last = opt.get_result(mol, long_bonds, subunits)
    p = orc.OptimizerParams(step_size=0.2, target_bond_length=1.5, num_steps=30)
    want = orc.optimizer_run(pos, long_bonds, subunits, p, np.random.default_rng(77))
    assert bool(last.passed) == bool(want.passed[-1])
    assert rel_close(last.system_potential, want.system_potential[-1], RTOL64)
    assert np.allclose(out_mol.get_position_matrix(), want.final_positions, rtol=RTOL64, atol=1e-9)
    with pytest.raises(ValueError):
        opt.get_result(mol, (), subunits)  # nothing to optimise (the reference's choice([]) raises too)
    with pytest.raises(StopIteration):
        opt.get_result(mol, long_bonds, {0: [0, 1, 2, 3, 4]})  # bond end in no subunit


def test_large_cage_against_oracle_both_precisions():
    # BASELINE config 5 geometry (5004 atoms, 36 subunits incl. 12 one-atom nodes): the fp64
    # chain fills an SM alone; the fp32 chain runs without the deferred-control buffer so that
    # two fit (csrc/mch_api.cu plan_chain_smem).  Few steps: the oracle needs ~0.25 s per step.
    mol, long_bonds, subunits = synthetic.cuboctahedron_cage()
    steps = 9
    p = orc.OptimizerParams(step_size=0.25, target_bond_length=1.2, num_steps=steps)
    want = orc.optimizer_run(mol.get_position_matrix(), long_bonds, subunits, p, np.random.default_rng(21))
    exact = mch.Optimizer(0.25, 1.2, steps, precision="fp64").get_results_batch(mol, long_bonds, subunits, seeds=[21, 22])
    assert int(exact.n_accepted[0]) == int((want.passed == 1).sum())
    assert rel_close(exact.system_potential[0], want.system_potential[-1], RTOL64)
    assert np.allclose(exact.positions[0], want.final_positions, rtol=RTOL64, atol=1e-9)
    fast = mch.Optimizer(0.25, 1.2, steps, precision="fp32").get_results_batch(mol, long_bonds, subunits, seeds=[21, 22])
    assert int(fast.n_accepted[0]) == int((want.passed == 1).sum())
    assert rel_close(fast.system_potential[0], want.system_potential[-1], RTOL32)
    assert np.allclose(fast.positions[0], want.final_positions, rtol=0, atol=1e-3)


@pytest.mark.parametrize("cluster", [2, 4, 8, 16])
@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_chain_on_a_thread_block_cluster(cluster, precision):
    # one chain spread over 2 / 4 / 8 / 16 CTAs (16 = the non-portable cluster size) (every CTA runs the control section redundantly, the sweep's
    # columns are split, totals and new energy-table rows travel through distributed shared memory):
    # same decisions and, up to the grouping of the sums, the same numbers as the one-CTA chain;
    # fp64 also against the oracle
    from mchammer_b200 import engine, rngstate
    from mchammer_b200.lowering import lower_topology
    import torch

    mol, long_bonds, subunits = synthetic.cube_cage()
    steps, seeds = 90, [11, 12, 13]
    opt = mch.Optimizer(step_size=0.25, target_bond_length=1.2, num_steps=steps, precision=precision)
    topo = lower_topology(mol.get_num_atoms(), long_bonds, subunits)
    dtopo = engine.upload_topology(topo, 0)
    pos = torch.from_numpy(mol.get_position_matrix()).cuda()

    def run(g):
        rng = engine.rng_words_to_device(rngstate.seed_states(seeds), torch.device("cuda", 0))
        out = engine.optimize_on_device(dtopo, pos, len(seeds), rng, opt._params(), steps, precision,
                                        want_trace=True, cluster_size=g)
        return (out.system_potential.cpu().numpy(), out.n_accepted.cpu().numpy(), out.positions.cpu().numpy(),
                out.trace_info.cpu().numpy(), out.rng_state.cpu().numpy())

    u1, acc1, p1, info1, rng1 = run(1)
    ug, accg, pg, infog, rngg = run(cluster)
    assert np.array_equal(info1, infog), "a chain on a cluster took different decisions"
    assert np.array_equal(acc1, accg) and np.array_equal(rng1, rngg)
    assert rel_close(ug, u1, 1e-12 if precision == "fp64" else 1e-5)
    assert np.allclose(pg, p1, rtol=0, atol=1e-9 if precision == "fp64" else 1e-4)
    if precision == "fp64":
        p = orc.OptimizerParams(step_size=0.25, target_bond_length=1.2, num_steps=steps)
        want = orc.optimizer_run(mol.get_position_matrix(), long_bonds, subunits, p, np.random.default_rng(seeds[0]))
        assert int(accg[0]) == int((want.passed == 1).sum())
        assert rel_close(ug[0], want.system_potential[-1], RTOL64)
        assert np.allclose(pg[0], want.final_positions, rtol=RTOL64, atol=1e-9)


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_threads_per_chain_does_not_change_the_chain(precision):
    # block size only changes how the pair sum is partitioned (rounding at the 1e-16 / 1e-7 level)
    mol, long_bonds, subunits = synthetic.cube_cage(atoms_per_subunit=24, seed=6, spread=8.0)
    opt = mch.Optimizer(0.25, 1.2, 60, precision=precision)
    seeds = [3, 4, 5]
    base = opt.get_results_batch(mol, long_bonds, subunits, seeds=seeds, threads_per_chain=32)
    limit = 512 if precision == "fp64" else 1024
    for threads in (64, 96, 128, 256, limit):
        got = opt.get_results_batch(mol, long_bonds, subunits, seeds=seeds, threads_per_chain=threads)
        assert np.array_equal(got.n_accepted, base.n_accepted), threads
        assert rel_close(got.system_potential, base.system_potential, 1e-12 if precision == "fp64" else 1e-5)
    with pytest.raises(mch.native.NativeLibraryError):
        opt.get_results_batch(mol, long_bonds, subunits, seeds=seeds, threads_per_chain=limit + 32)
    with pytest.raises(mch.native.NativeLibraryError):
        opt.get_results_batch(mol, long_bonds, subunits, seeds=seeds, threads_per_chain=48)


# ------------------------------------------------------------------------- properties at BASELINE size
@pytest.mark.parametrize("precision,rtol,atol", [("fp64", 1e-9, 1e-9), ("fp32", 1e-4, 2e-4)])
def test_full_size_cage_properties(precision, rtol, atol):
    # BASELINE config 4 geometry (20 subunits x 50 atoms), fewer chains/steps than the bench
    mol, long_bonds, subunits = synthetic.cube_cage()
    chains, steps = 48, 120
    opt = mch.Optimizer(step_size=0.25, target_bond_length=1.2, num_steps=steps, precision=precision)
    seeds = list(range(chains))
    got = opt.get_results_batch(mol, long_bonds, subunits, seeds=seeds)
    start = mol.get_position_matrix()
    assert got.positions.shape == (chains, 1000, 3)
    assert np.all(got.n_accepted >= 1) and np.all(got.n_accepted <= steps - 1)
    # step 0 evaluates every pair once; each move evaluates n_moving x (N - n_moving) pairs
    assert np.all(got.pair_evals == 1000 * 999 // 2 + (steps - 1) * 50 * 950)
    # 1. rigidity: every subunit moved as a rigid body (pure translation)
    for ids in subunits.values():
        delta = got.positions[:, ids, :] - start[ids]
        assert np.abs(delta - delta[:, :1, :]).max() < atol
    # 2. the potential carried incrementally through all steps equals a fresh evaluation
    fresh = mch.Optimizer(0.25, 1.2, 2, precision="fp64")
    for c in (0, chains // 2, chains - 1):
        u, unb = fresh.compute_potential(mol.with_position_matrix(got.positions[c]), long_bonds)
        assert np.isclose(got.system_potential[c], u, rtol=rtol, atol=0)
        assert np.isclose(got.nonbonded_potential[c], unb, rtol=rtol, atol=0)
    # 3. same seed -> same chain; different seeds -> different chains
    again = opt.get_results_batch(mol, long_bonds, subunits, seeds=seeds[:4])
    assert np.array_equal(again.positions, got.positions[:4])
    assert len({float(u) for u in got.system_potential}) > chains // 2
    # 4. the optimisation does optimise: the system potential drops, the longest bond does not grow
    u0, _ = fresh.compute_potential(mol, long_bonds)
    assert got.system_potential.mean() < 0.9 * u0
    assert got.max_bond_distance.mean() < max(np.linalg.norm(start[a] - start[b]) for a, b in long_bonds)


# ------------------------------------------------------------------------- Collapser
@pytest.mark.parametrize("case", ["collapser_poc", "collapser_poc_noscale", "collapser_toy"])
def test_collapser_matches_reference(case):
    g = load_golden(case)
    mol = golden_molecule(g)
    coll = mch.Collapser(step_size=float(g["step_size"]), distance_threshold=float(g["distance_threshold"]),
                         scale_steps=bool(g["scale_steps"]))
    subunits, bonds = golden_subunits(g), golden_bonds(g)
    out_mol, result = coll.get_trajectory(mol, bonds, subunits)
    assert result.get_step_count() == int(g["steps_run"])
    far = np.array([p["max_bond_distance"] for _, p in result.get_steps_properties()])
    assert rel_close(far, g["max_bond_distance"], RTOL64)
    frames = np.array([f for _, f in result.get_trajectory()])
    assert np.allclose(frames, g["positions"], rtol=RTOL64, atol=1e-9)
    assert np.allclose(out_mol.get_position_matrix(), g["final_positions"], rtol=RTOL64, atol=1e-9)
    assert f"Steps run: {int(g['steps_run'])}\n" in result.get_log()
    logged = float(result.get_log().split("Minimum inter-subunit distance: ")[1].split()[0])
    assert rel_close(logged, float(g["min_distance_logged"]), RTOL64)
    out_mol2, last = coll.get_result(mol, bonds, subunits)
    assert last.step == int(g["steps_run"])
    assert np.allclose(out_mol2.get_position_matrix(), g["final_positions"], rtol=RTOL64, atol=1e-9)
    assert rel_close(last.max_bond_distance, g["max_bond_distance"][-1], RTOL64)


def test_collapser_known_answers():
    # reference tests/collapser/test_collapser.py:69-160
    pos = np.array([[0, 1, 0], [1, 1, 0], [-1, 1, 0], [0, 10, 0], [1, 10, 0], [-1, 10, 0]])
    bonds = [mch.Bond(0, (0, 1)), mch.Bond(1, (0, 2)), mch.Bond(2, (0, 3)), mch.Bond(3, (3, 4)), mch.Bond(4, (3, 5))]
    mol = mch.Molecule(atoms=[mch.Atom(i, "C") for i in range(6)], bonds=bonds, position_matrix=pos)
    coll = mch.Collapser(step_size=0.05, distance_threshold=1.5, scale_steps=True)
    subunits = mol.get_subunits(bond_pair_ids=((0, 3),))
    test_mol, results = coll.get_trajectory(mol=mol, bond_pair_ids=((0, 3),), subunits=subunits)
    assert results.get_step_count() == 35
    final_min = min(coll._get_subunit_distances(test_mol, subunits))
    assert np.isclose(1.4947505, final_min, atol=1e-8)
    want = np.array([[0.0, 4.75262477, 0.0], [1.0, 4.75262477, 0.0], [-1.0, 4.75262477, 0.0],
                     [0.0, 6.24737523, 0.0], [1.0, 6.24737523, 0.0], [-1.0, 6.24737523, 0.0]])
    assert np.allclose(test_mol.get_position_matrix(), want, atol=1e-8)


def test_collapser_batch_with_hydrogens_against_oracle():
    mol, long_bonds, subunits = synthetic.cube_cage(atoms_per_subunit=12, seed=9, spread=7.0, h_fraction=0.3)
    rng = np.random.default_rng(2)
    base = mol.get_position_matrix()
    positions = np.stack([base * s for s in (1.0, 1.1, 1.25)])
    is_h = [a.get_element_string() == "H" for a in mol.get_atoms()]
    assert any(is_h)
    coll = mch.Collapser(step_size=0.04, distance_threshold=1.4, scale_steps=True)
    got = coll.get_results_batch(mol, long_bonds, subunits, positions)
    for m in range(positions.shape[0]):
        want = orc.collapser_run(positions[m], is_h, long_bonds, subunits, 0.04, 1.4, True)
        assert int(got.steps_run[m]) == want.steps_run
        assert rel_close(got.min_distance[m], want.min_distance, RTOL64)
        assert np.allclose(got.positions[m], want.final_positions, rtol=RTOL64, atol=1e-9)
        assert rel_close(got.max_bond_distance[m], want.max_bond_distance[-1], RTOL64)
    del rng


@pytest.mark.parametrize("threads", [0, 96, 256])
def test_collapser_full_size_cage_block_dealing(threads):
    # 1000 atoms, ~700 heavy: six 128-column tiles x up to 21 row blocks of the contact-test
    # triangle, dealt over 4 / 3 / 8 warps -- steps, contact distance and geometry vs the oracle
    from mchammer_b200 import engine
    from mchammer_b200.lowering import lower_topology
    import torch

    mol, long_bonds, subunits = synthetic.cube_cage(h_fraction=0.3)
    base = mol.get_position_matrix()
    elements = [a.get_element_string() for a in mol.get_atoms()]
    is_h = [e == "H" for e in elements]
    topo = lower_topology(mol.get_num_atoms(), long_bonds, subunits, elements=elements, require_bond_subunits=False)
    dtopo = engine.upload_topology(topo, 0)
    positions = np.stack([base, base * 1.03])
    out = engine.collapse_on_device(dtopo, torch.from_numpy(positions).cuda(), 2, 0.02, 1.5, True, "fp64",
                                    threads_per_mol=threads)
    for m in range(2):
        want = orc.collapser_run(positions[m], is_h, long_bonds, subunits, 0.02, 1.5, True)
        assert int(out.steps_run[m].item()) == want.steps_run
        assert rel_close(out.min_distance[m].item(), want.min_distance, RTOL64)
        assert np.allclose(out.positions[m].cpu().numpy(), want.final_positions, rtol=RTOL64, atol=1e-9)


@pytest.mark.parametrize("seed", range(8))
def test_collapser_random_clusters_against_oracle(seed):
    # seeded fuzz: 2-9 clusters of 1-70 atoms with random hydrogens (a cluster may be all H but one),
    # subunit ids interleaved with atom order, random step / threshold / scaling
    rng = np.random.default_rng(4321 + seed)
    n_sub = int(rng.integers(2, 10))
    sizes = [int(rng.integers(1, 71)) for _ in range(n_sub)]
    centres = rng.normal(scale=14.0, size=(n_sub, 3))
    pos = np.concatenate([c + rng.normal(scale=1.5, size=(k, 3)) for c, k in zip(centres, sizes)])
    n = pos.shape[0]
    owner = np.repeat(np.arange(n_sub), sizes)
    order = rng.permutation(n)           # atom ids are shuffled with respect to the clusters
    pos, owner = pos[order], owner[order]
    elements = ["H" if rng.random() < 0.35 else "C" for _ in range(n)]
    for k in range(n_sub):               # every subunit keeps at least one heavy atom
        ids = np.flatnonzero(owner == k)
        if all(elements[i] == "H" for i in ids):
            elements[int(ids[0])] = "N"
    subunits = {int(7 * k + 1): [int(i) for i in np.flatnonzero(owner == k)] for k in rng.permutation(n_sub)}
    bonds = tuple((int(a), int(b)) for a, b in (rng.choice(n, size=2, replace=False) for _ in range(int(rng.integers(1, 6)))))
    step, thr, scale = float(rng.uniform(0.02, 0.12)), float(rng.uniform(1.0, 2.2)), bool(rng.random() < 0.5)
    mol = mch.Molecule(atoms=[mch.Atom(i, e) for i, e in enumerate(elements)], bonds=[], position_matrix=pos)
    is_h = [e == "H" for e in elements]
    want = orc.collapser_run(pos, is_h, bonds, subunits, step, thr, scale)
    coll = mch.Collapser(step_size=step, distance_threshold=thr, scale_steps=scale)
    got = coll.get_results_batch(mol, bonds, subunits, pos[None])
    assert int(got.steps_run[0]) == want.steps_run
    assert rel_close(got.min_distance[0], want.min_distance, RTOL64)
    assert np.allclose(got.positions[0], want.final_positions, rtol=RTOL64, atol=1e-9)
    if want.steps_run:
        assert rel_close(got.max_bond_distance[0], want.max_bond_distance[-1], RTOL64)


def test_collapser_never_touching_raises():
    pos = np.array([[0.0, 0, 0], [1.5, 0, 0], [3.0, 0, 0]])
    mol = mch.Molecule(atoms=[mch.Atom(i, "C") for i in range(3)], bonds=[], position_matrix=pos)
    coll = mch.Collapser(step_size=0.05, distance_threshold=0.5, max_steps=50)
    with pytest.raises(RuntimeError):
        coll.get_result(mol, ((0, 1),), {0: [0, 1, 2]})
