"""GPU parity tests added in round 2 (all through the C ABI): the configurations round 1 left
untested -- Collapser fast mode, fast mode at the bench configuration (1000 atoms x 500 steps),
the 5004-atom cage over >= 50 steps, a real two-device plan, the stk-shaped batched entry --
plus the advisor's findings (atoms outside every subunit in the Collapser, explicit block sizes
with the automatic cluster policy, bond ids out of range, slice-independent launch choices).

Tolerances (BASELINE.json north_star): fp64 = reference accept/reject sequence, 1e-9 relative;
fp32 fast mode = potentials within 1e-4 relative while the accept sequence agrees, and a carried
potential within 1e-4 relative of a fresh fp64 evaluation of the chain's own geometry afterwards.
"""

from __future__ import annotations

import numpy as np
import pytest

import mchammer_b200 as mch
from mchammer_b200 import stk_helpers, synthetic
from mchammer_b200.lowering import lower_topology
from oracle import mch_oracle as orc
from tests.conftest import golden_bonds, golden_subunits, load_golden
from tests.test_gpu_parity import RTOL32, RTOL64, golden_molecule, rel_close

pytestmark = pytest.mark.gpu


def same_chains(got, want, precision, max_forked=0):
    """Chain-by-chain equality of two runs of the same seeds.  fp64: identical decisions and
    potentials to 1e-12.  fp32: the grouping of the fp32 partial sums depends on the block size,
    so a chain may fork when an accept test lands within ~1e-7 relative of its threshold; at most
    `max_forked` chains may, all others must agree to 1e-6."""
    equal = got.n_accepted == want.n_accepted
    if precision == "fp64":
        assert equal.all()
        assert rel_close(got.system_potential, want.system_potential, 1e-12)
        return
    assert (~equal).sum() <= max_forked, f"{(~equal).sum()} chains forked"
    assert rel_close(got.system_potential[equal], want.system_potential[equal], 1e-6)


def common_prefix(a, b) -> int:
    same = np.asarray(a) == np.asarray(b)
    return len(same) if same.all() else int(np.argmin(same))


# ------------------------------------------------------------------------- Collapser
def test_collapser_atoms_outside_every_subunit_match_reference():
    # reference collapser.py:56-104: such atoms never move, are not contact-tested and do not enter
    # the step scales; they only count in the molecule's centroid (golden = the reference's output)
    g = load_golden("collapser_poc_loose")
    mol = golden_molecule(g)
    subunits, bonds = golden_subunits(g), golden_bonds(g)
    loose = sorted(set(range(mol.get_num_atoms())) - {a for ids in subunits.values() for a in ids})
    assert len(loose) == 16
    coll = mch.Collapser(step_size=float(g["step_size"]), distance_threshold=float(g["distance_threshold"]),
                         scale_steps=bool(g["scale_steps"]))
    out_mol, result = coll.get_trajectory(mol, bonds, subunits)
    assert result.get_step_count() == int(g["steps_run"])
    frames = np.array([f for _, f in result.get_trajectory()])
    assert np.allclose(frames, g["positions"], rtol=RTOL64, atol=1e-9)
    assert np.array_equal(out_mol.get_position_matrix()[loose], g["pos0"][loose])  # bit for bit
    far = np.array([p["max_bond_distance"] for _, p in result.get_steps_properties()])
    assert rel_close(far, g["max_bond_distance"], RTOL64)
    logged = float(result.get_log().split("Minimum inter-subunit distance: ")[1].split()[0])
    assert rel_close(logged, float(g["min_distance_logged"]), RTOL64)


@pytest.mark.parametrize("seed", range(4))
def test_collapser_loose_atoms_fuzz_against_oracle(seed):
    rng = np.random.default_rng(900 + seed)
    n_sub = int(rng.integers(2, 7))
    sizes = [int(rng.integers(2, 40)) for _ in range(n_sub)]
    centres = rng.normal(scale=12.0, size=(n_sub, 3))
    pos = np.concatenate([c + rng.normal(scale=1.4, size=(k, 3)) for c, k in zip(centres, sizes)]
                         + [rng.normal(scale=3.0, size=(int(rng.integers(1, 9)), 3))])  # loose atoms near the centre
    n = pos.shape[0]
    owner = np.concatenate([np.repeat(np.arange(n_sub), sizes), np.full(n - sum(sizes), -1)])
    order = rng.permutation(n)
    pos, owner = pos[order], owner[order]
    elements = ["H" if rng.random() < 0.3 else "C" for _ in range(n)]
    for k in range(n_sub):
        ids = np.flatnonzero(owner == k)
        if all(elements[i] == "H" for i in ids):
            elements[int(ids[0])] = "O"
    subunits = {int(3 * k + 2): [int(i) for i in np.flatnonzero(owner == k)] for k in rng.permutation(n_sub)}
    bonds = ((0, n - 1), (1, n // 2))
    step, thr = float(rng.uniform(0.03, 0.1)), float(rng.uniform(1.0, 2.0))
    mol = mch.Molecule(atoms=[mch.Atom(i, e) for i, e in enumerate(elements)], bonds=[], position_matrix=pos)
    is_h = [e == "H" for e in elements]
    want = orc.collapser_run(pos, is_h, bonds, subunits, step, thr, True)
    got = mch.Collapser(step_size=step, distance_threshold=thr).get_results_batch(mol, bonds, subunits, pos[None])
    assert int(got.steps_run[0]) == want.steps_run
    assert rel_close(got.min_distance[0], want.min_distance, RTOL64)
    assert np.allclose(got.positions[0], want.final_positions, rtol=RTOL64, atol=1e-9)


@pytest.mark.parametrize("case", ["collapser_poc", "collapser_poc_noscale", "collapser_toy", "collapser_poc_loose"])
def test_collapser_fp32_mode_against_reference(case):
    # Collapser(precision="fp32"): positions and squared distances in float32 (centred frame),
    # centroids / scales / step vectors in float64.  The step count is a discrete outcome of
    # `min distance < threshold` on an fp32 r^2 (relative error ~1e-6); it must equal the
    # reference's unless the reference's own deciding distance lies within 1e-5 relative of the
    # threshold (then +-1 step is allowed).  With equal step counts: positions within 1e-4 A,
    # contact distance and longest bond within 1e-4 relative.
    g = load_golden(case)
    mol = golden_molecule(g)
    subunits, bonds = golden_subunits(g), golden_bonds(g)
    thr = float(g["distance_threshold"])
    coll = mch.Collapser(step_size=float(g["step_size"]), distance_threshold=thr,
                         scale_steps=bool(g["scale_steps"]), precision="fp32")
    out_mol, result = coll.get_trajectory(mol, bonds, subunits)
    want_steps = int(g["steps_run"])
    is_h = [str(e) == "H" for e in g["elements"]]
    knife_edge = False
    for frame in [g["pos0"], *g["positions"]]:
        d = orc.min_subunit_distance(np.asarray(frame), is_h, subunits)
        knife_edge |= abs(d - thr) < 1e-5 * thr or abs(d - thr / 2) < 1e-5 * thr
    if knife_edge:
        assert abs(result.get_step_count() - want_steps) <= 1
        return
    assert result.get_step_count() == want_steps
    frames = np.array([f for _, f in result.get_trajectory()])
    assert np.allclose(frames, g["positions"], rtol=0, atol=1e-4)
    assert np.allclose(out_mol.get_position_matrix(), g["final_positions"], rtol=0, atol=1e-4)
    far = np.array([p["max_bond_distance"] for _, p in result.get_steps_properties()])
    assert rel_close(far, g["max_bond_distance"], RTOL32)
    logged = float(result.get_log().split("Minimum inter-subunit distance: ")[1].split()[0])
    assert rel_close(logged, float(g["min_distance_logged"]), RTOL32)


def test_collapser_fp32_batch_with_hydrogens_against_oracle():
    mol, long_bonds, subunits = synthetic.cube_cage(atoms_per_subunit=12, seed=9, spread=7.0, h_fraction=0.3)
    base = mol.get_position_matrix()
    positions = np.stack([base * s for s in (1.0, 1.1, 1.25)])
    is_h = [a.get_element_string() == "H" for a in mol.get_atoms()]
    coll = mch.Collapser(step_size=0.04, distance_threshold=1.4, scale_steps=True, precision="fp32")
    got = coll.get_results_batch(mol, long_bonds, subunits, positions)
    for m in range(positions.shape[0]):
        want = orc.collapser_run(positions[m], is_h, long_bonds, subunits, 0.04, 1.4, True)
        assert int(got.steps_run[m]) == want.steps_run  # not a knife-edge case (checked when the test was written)
        assert rel_close(got.min_distance[m], want.min_distance, RTOL32)
        assert np.allclose(got.positions[m], want.final_positions, rtol=0, atol=1e-4)
        assert rel_close(got.max_bond_distance[m], want.max_bond_distance[-1], RTOL32)


# ------------------------------------------------------------------------- fast mode at bench size
@pytest.mark.parametrize("seed", [1000, 1001, 1002, 1003])
def test_fp32_bench_configuration_against_oracle(seed):
    # BASELINE configs[3]: the 1000-atom cube cage, num_steps = 500, seeds of the bench (1000 + c).
    # The fast-mode chain is compared with the fp64 oracle over the common prefix of the
    # accept/reject sequence (a fork needs an accept test within fp32 noise), then its carried
    # potentials are checked against a fresh fp64 evaluation of its own final geometry.
    mol, long_bonds, subunits = synthetic.cube_cage()
    steps = 500
    p = orc.OptimizerParams(step_size=0.25, target_bond_length=1.2, num_steps=steps)
    want = orc.optimizer_run(mol.get_position_matrix(), long_bonds, subunits, p, np.random.default_rng(seed))
    fast = mch.Optimizer(0.25, 1.2, steps, precision="fp32")
    traj = fast.get_trajectories_batch(mol, long_bonds, subunits, seeds=[seed], want_positions=False)
    prefix = common_prefix(traj.passed[0], want.passed)
    assert prefix >= 100, f"fast mode forked from the oracle at step {prefix}"
    assert np.array_equal(traj.bond_index[0][:prefix], want.bond_index[:prefix])
    assert rel_close(traj.system_potential[0][:prefix], want.system_potential[:prefix], RTOL32)
    assert rel_close(traj.nonbonded_potential[0][:prefix], want.nonbonded_potential[:prefix], RTOL32)
    assert rel_close(traj.max_bond_distance[0][:prefix], want.max_bond_distance[:prefix], RTOL32)
    # after all 500 steps: carried potential == potential of the chain's own geometry
    final = mol.with_position_matrix(np.array(traj.final_positions[0], dtype=np.float64))
    u, unb = mch.Optimizer(0.25, 1.2, 1, precision="fp64").compute_potential(final, long_bonds)
    assert rel_close(traj.system_potential[0][-1], u, RTOL32)
    assert rel_close(traj.nonbonded_potential[0][-1], unb, RTOL32)


def test_fp32_and_fp64_accept_statistics_agree_at_bench_size():
    # 4096 chains x 500 steps, both modes, same seeds: chains fork individually, the ensemble must not
    mol, long_bonds, subunits = synthetic.cube_cage()
    seeds = list(range(1000, 1000 + 4096))
    got = {}
    for precision in ("fp32", "fp64"):
        opt = mch.Optimizer(0.25, 1.2, 500, precision=precision)
        got[precision] = opt.get_results_batch(mol, long_bonds, subunits, seeds=seeds)
    a32, a64 = got["fp32"].n_accepted.astype(float), got["fp64"].n_accepted.astype(float)
    assert abs(a32.mean() - a64.mean()) < 0.005 * a64.mean()
    assert abs(a32.std() - a64.std()) < 0.05 * a64.std()
    u32, u64 = got["fp32"].system_potential, got["fp64"].system_potential
    assert abs(u32.mean() - u64.mean()) < 0.005 * abs(u64.mean())
    # most chains have not forked at all over 500 steps
    assert (got["fp32"].n_accepted == got["fp64"].n_accepted).mean() > 0.5


def test_m12l24_fast_mode_over_sixty_steps():
    # BASELINE configs[4] geometry, one chain, 60 steps (the oracle needs ~0.25 s per step)
    mol, long_bonds, subunits = synthetic.cuboctahedron_cage()
    steps = 60
    p = orc.OptimizerParams(step_size=0.25, target_bond_length=1.2, num_steps=steps)
    want = orc.optimizer_run(mol.get_position_matrix(), long_bonds, subunits, p, np.random.default_rng(31))
    for precision, rtol in (("fp32", RTOL32), ("fp64", RTOL64)):
        opt = mch.Optimizer(0.25, 1.2, steps, precision=precision)
        traj = opt.get_trajectories_batch(mol, long_bonds, subunits, seeds=[31], want_positions=False)
        prefix = common_prefix(traj.passed[0], want.passed)
        assert prefix == steps, f"{precision}: forked at step {prefix}"
        assert rel_close(traj.system_potential[0], want.system_potential, rtol)
        assert rel_close(traj.nonbonded_potential[0], want.nonbonded_potential, rtol)
        assert np.allclose(traj.final_positions[0], want.final_positions, rtol=0,
                           atol=1e-9 if precision == "fp64" else 1e-3)


# ------------------------------------------------------------------------- devices, slices, policies
def test_batch_plan_on_two_devices():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    mol, long_bonds, subunits = synthetic.cube_cage(atoms_per_subunit=10, seed=4, spread=6.0)
    seeds = list(range(70, 83))
    opt = mch.Optimizer(step_size=0.25, target_bond_length=1.2, num_steps=50)
    one = opt.get_results_batch(mol, long_bonds, subunits, seeds=seeds, devices=[0])
    two = opt.get_results_batch(mol, long_bonds, subunits, seeds=seeds, devices=[0, 1])
    assert np.array_equal(one.positions, two.positions)
    assert np.array_equal(one.system_potential, two.system_potential)
    assert np.array_equal(one.n_accepted, two.n_accepted)
    traj = opt.get_trajectories_batch(mol, long_bonds, subunits, seeds=seeds, devices=[1, 0])
    assert np.array_equal(traj.final_positions, one.positions)


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_slices_and_block_sizes_do_not_change_results(precision):
    # a medium batch of a large molecule: with per-slice heuristics the slices of a pipelined run
    # could flip into the latency / cluster choices; the batch's choices must hold for every slice
    mol, long_bonds, subunits = synthetic.cube_cage()
    seeds = list(range(300, 324))
    opt = mch.Optimizer(0.25, 1.2, 40, precision=precision)
    whole = opt.prepare_batch(mol, long_bonds, subunits, len(seeds), n_slices=1).run(seeds=seeds).copy()
    for n_slices in (3, 8):
        cut = opt.prepare_batch(mol, long_bonds, subunits, len(seeds), n_slices=n_slices).run(seeds=seeds).copy()
        same_chains(cut, whole, precision, max_forked=1)
    # explicit block sizes, including ones a chain on a cluster cannot take (advisor, round 1):
    # few chains of a 1000-atom cage used to be refused with threads_per_chain = 32
    few = seeds[:3]
    base = opt.get_results_batch(mol, long_bonds, subunits, seeds=few)
    for threads in (32, 64, 160, 256, 512) + ((1024,) if precision == "fp32" else ()):
        got = opt.get_results_batch(mol, long_bonds, subunits, seeds=few, threads_per_chain=threads)
        same_chains(got, base, precision, max_forked=0)


@pytest.mark.parametrize("chains", [148, 300, 512, 1024])
def test_medium_batches_match_the_full_batch_choice(chains):
    # the block size now follows chains per SM (149 .. ~1184 chains used to get the full-batch choice):
    # whatever is chosen, chain c must be the same chain
    mol, long_bonds, subunits = synthetic.cube_cage()
    opt = mch.Optimizer(0.25, 1.2, 30, precision="fp32")
    seeds = list(range(5000, 5000 + chains))
    got = opt.get_results_batch(mol, long_bonds, subunits, seeds=seeds)
    want = opt.get_results_batch(mol, long_bonds, subunits, seeds=seeds, threads_per_chain=64)
    same_chains(got, want, "fp32", max_forked=max(1, chains // 200))


def test_compute_potential_rejects_bond_ids_outside_the_molecule():
    opt = mch.Optimizer(step_size=0.1, target_bond_length=2.0, num_steps=10)
    pos = np.array([[0, 1, 0], [1, 1, 0], [-1, 1, 0], [0, 10, 0.0]])
    mol = mch.Molecule(atoms=[mch.Atom(i, "C") for i in range(4)], bonds=[], position_matrix=pos)
    with pytest.raises(IndexError):
        opt.compute_potential(mol, ((0, 4),))
    with pytest.raises(IndexError):
        opt.compute_potential(mol, ((-5, 1),))
    u, unb = opt.compute_potential(mol, ((0, -1),))  # numpy-style negative id = last atom, like the reference
    want = orc.compute_potential(pos, [(0, 3)], orc.OptimizerParams(step_size=0.1, target_bond_length=2.0, num_steps=10))
    assert rel_close(u, want[0], 1e-12) and rel_close(unb, want[1], 1e-12)


# ------------------------------------------------------------------------- stk-shaped batched entry
class _Atom:
    def __init__(self, i): self.i = i
    def get_id(self): return self.i


class C(_Atom): pass  # noqa: E701  (stk: the atom's class name is its element)


class N(_Atom): pass  # noqa: E701


class _Bond:
    def __init__(self, a, b): self.a, self.b = _Atom(a), _Atom(b)
    def get_atom1(self): return self.a
    def get_atom2(self): return self.b


class _BondInfo:
    def __init__(self, a, b, bb): self.bond, self.bb = _Bond(a, b), bb
    def get_bond(self): return self.bond
    def get_building_block(self): return self.bb


class _AtomInfo:
    def __init__(self, i, bb): self.atom, self.bb = _Atom(i), bb
    def get_atom(self): return self.atom
    def get_building_block_id(self): return self.bb


class _State:
    """What stk_helpers needs of an stk ConstructedMolecule / ConstructionState."""

    def __init__(self, pos, owner, intra_bonds, long_bonds):
        self.pos, self.owner, self.intra, self.long = np.asarray(pos, float), owner, intra_bonds, long_bonds

    def get_atoms(self): return [(C if i % 3 else N)(i) for i in range(len(self.owner))]
    def get_atom_infos(self): return [_AtomInfo(i, int(b)) for i, b in enumerate(self.owner)]
    def get_bond_infos(self):
        return ([_BondInfo(a, b, "bb") for a, b in self.intra] + [_BondInfo(a, b, None) for a, b in self.long])
    def get_position_matrix(self): return self.pos.copy()
    def with_position_matrix(self, position_matrix):
        return _State(position_matrix, self.owner, self.intra, self.long)


def test_stk_optimize_batch_against_oracle():
    # reference tests/stk_reproduction/test_stk.py:97-135 wiring, for a small screening set: four
    # conformers of one topology and one molecule of another topology, all in one call
    rng = np.random.default_rng(77)
    mol, long_bonds, subunits = synthetic.cube_cage(atoms_per_subunit=8, seed=5, spread=5.5)
    base = mol.get_position_matrix()
    owner = np.zeros(base.shape[0], dtype=int)
    for k, ids in subunits.items():
        owner[ids] = k
    intra = [(b.get_atom1_id(), b.get_atom2_id()) for b in mol.get_bonds()]
    long_unsorted = [(b, a) for a, b in long_bonds]  # stk lists bond ends in either order; the helper sorts
    states = [_State(base + rng.normal(scale=0.03, size=base.shape), owner, intra, long_unsorted) for _ in range(4)]
    mol2, lb2, su2 = synthetic.cube_cage(atoms_per_subunit=5, seed=8, spread=5.0)
    owner2 = np.zeros(mol2.get_num_atoms(), dtype=int)
    for k, ids in su2.items():
        owner2[ids] = k
    states.insert(2, _State(mol2.get_position_matrix(), owner2,
                            [(b.get_atom1_id(), b.get_atom2_id()) for b in mol2.get_bonds()], list(lb2)))
    opt = mch.Optimizer(step_size=0.25, target_bond_length=1.2, num_steps=60)
    results = stk_helpers.optimize_batch(states, opt, seeds=1000)
    assert len(results) == len(states)
    p = orc.OptimizerParams(step_size=0.25, target_bond_length=1.2, num_steps=60)
    for state, (new_state, record) in zip(states, results):
        bonds = stk_helpers.get_long_bond_ids(state)
        want = orc.optimizer_run(state.get_position_matrix(), bonds, dict(stk_helpers.get_subunits(state)), p,
                                 np.random.default_rng(1000))
        assert isinstance(new_state, _State)
        assert np.allclose(new_state.get_position_matrix(), want.final_positions, rtol=RTOL64, atol=1e-9)
        assert rel_close(record["system_potential"], want.system_potential[-1], RTOL64)
        assert record["n_accepted"] == int((want.passed == 1).sum())
        assert record["passed"] == bool(want.passed[-1])


def test_topology_cache_returns_what_a_fresh_lowering_would():
    # the drop-in calls keep the lowered + uploaded topology of recent (bonds, subunits) inputs; the key is the
    # exact input (order of the bond pairs, dict order and member order of the subunits)
    from mchammer_b200 import engine

    mol, long_bonds, subunits = synthetic.cube_cage(atoms_per_subunit=6, seed=3, spread=5.0)
    n = mol.get_num_atoms()
    engine._TOPOLOGY_CACHE.clear()
    a = engine.cached_topology(n, long_bonds, subunits, 0, allow_overlap=True)
    b = engine.cached_topology(n, [tuple(p) for p in long_bonds], {k: list(v) for k, v in subunits.items()}, 0,
                               allow_overlap=True)
    assert a is b and len(engine._TOPOLOGY_CACHE) == 1
    flipped = [tuple(long_bonds[0][::-1])] + [tuple(p) for p in long_bonds[1:]]  # the sign of one bond vector
    c = engine.cached_topology(n, flipped, subunits, 0, allow_overlap=True)
    assert c is not a and not np.array_equal(c.host.bonds, a.host.bonds)
    reordered = dict(reversed(list(subunits.items())))  # dict order decides which subunit owns a bond end
    d = engine.cached_topology(n, long_bonds, reordered, 0, allow_overlap=True)
    assert d is not a
    fresh = lower_topology(n, long_bonds, subunits, allow_overlap=True)
    assert np.array_equal(a.perm.cpu().numpy(), fresh.perm) and np.array_equal(a.su_offsets.cpu().numpy(), fresh.su_offsets)
    assert np.array_equal(a.row_slots.cpu().numpy(), fresh.row_slots) and np.array_equal(a.runs.cpu().numpy(), fresh.runs.reshape(-1))
    # same chain with a warm and with a cold cache
    opt1 = mch.Optimizer(0.25, 1.2, 40, random_seed=5)
    warm = opt1.get_result(mol, long_bonds, subunits)[1]
    engine._TOPOLOGY_CACHE.clear()
    cold = mch.Optimizer(0.25, 1.2, 40, random_seed=5).get_result(mol, long_bonds, subunits)[1]
    assert warm.system_potential == cold.system_potential and np.array_equal(warm.position_matrix, cold.position_matrix)
    for k in range(engine._TOPOLOGY_CACHE_SIZE + 3):  # bounded
        engine.cached_topology(n, long_bonds[: 1 + k % len(long_bonds)] * (1 + k // len(long_bonds)), subunits, 0, allow_overlap=True)
    assert len(engine._TOPOLOGY_CACHE) <= engine._TOPOLOGY_CACHE_SIZE


def test_long_fp64_chain_keeps_the_reference_sequence():
    # 3000 steps of poc.mol (six times the reference's own example length): the accept/reject sequence still
    # equals the oracle's, potentials and positions within 1e-9 -- no drift in the carried energy table, the
    # coordinate sums or the random stream
    g = load_golden("poc_graph_seed1000")
    mol = golden_molecule(g)
    bonds, subunits = golden_bonds(g), golden_subunits(g)
    steps, seed = 3000, 77
    p = orc.OptimizerParams(step_size=0.25, target_bond_length=1.2, num_steps=steps)
    want = orc.optimizer_run(mol.get_position_matrix(), bonds, subunits, p, np.random.default_rng(seed))
    opt = mch.Optimizer(0.25, 1.2, steps, random_seed=seed)
    out_mol, result = opt.get_trajectory(mol, bonds, subunits)
    props = [q for _, q in result.get_steps_properties()]
    passed = np.array([-1 if q["passed"] is None else int(q["passed"]) for q in props])
    assert np.array_equal(passed, want.passed)
    assert rel_close([q["system_potential"] for q in props], want.system_potential, RTOL64)
    assert rel_close([q["nonbonded_potential"] for q in props], want.nonbonded_potential, RTOL64)
    assert np.allclose(out_mol.get_position_matrix(), want.final_positions, rtol=RTOL64, atol=1e-9)
    # fast mode over the same 3000 steps: the carried potential still equals a fresh evaluation of its own geometry
    fast = mch.Optimizer(0.25, 1.2, steps, random_seed=seed, precision="fp32")
    fmol, last = fast.get_result(mol, bonds, subunits)
    fresh_u, _ = mch.Optimizer(0.25, 1.2, 2).compute_potential(fmol, bonds)
    assert rel_close(last.system_potential, fresh_u, RTOL32)


@pytest.mark.parametrize("n_per_subunit", [12, 50])
def test_collapser_decides_at_the_threshold_like_the_reference(n_per_subunit):
    # The contact pass works in a fast form that cannot resolve 1e-12; every pair it puts near the threshold is
    # re-evaluated in the plain form (csrc/mch_collapse.cuh).  Thresholds placed a hair above / below the minimum
    # distance the reference sees after step k must end the collapse at step k / let it run on, exactly like the
    # oracle -- fp64 decides such cases identically, fp32 within its stated +-1 step.
    mol, long_bonds, subunits = synthetic.cube_cage(atoms_per_subunit=n_per_subunit, seed=4, spread=7.0, h_fraction=0.25)
    pos0 = mol.get_position_matrix()
    is_h = [a.get_element_string() == "H" for a in mol.get_atoms()]
    step_size, k = 0.03, 6
    pos = pos0.copy()
    for _ in range(k):
        pos = orc.collapser_step(pos, subunits, step_size, True)
    d_k = orc.min_subunit_distance(pos, is_h, subunits)
    for factor in (1.0 + 1e-12, 1.0 - 1e-12, 1.0 + 1e-9, 1.0 - 1e-9):
        thr = d_k * factor
        want = orc.collapser_run(pos0, is_h, long_bonds, subunits, step_size, thr, True)
        coll = mch.Collapser(step_size=step_size, distance_threshold=thr, scale_steps=True)
        got_mol, result = coll.get_trajectory(mol, long_bonds, subunits)
        assert result.get_step_count() == want.steps_run, (factor, result.get_step_count(), want.steps_run)
        assert np.allclose(got_mol.get_position_matrix(), want.final_positions, rtol=0, atol=1e-9)
        fast = mch.Collapser(step_size=step_size, distance_threshold=thr, scale_steps=True, precision="fp32")
        fmol, fres = fast.get_trajectory(mol, long_bonds, subunits)
        assert abs(fres.get_step_count() - want.steps_run) <= 1


def test_stk_collapse_batch_against_oracle():
    # stk's Collapser optimizer wiring for a small set: three conformers of one topology and one molecule of
    # another, one call; every molecule must end where the oracle's Collapser puts it
    rng = np.random.default_rng(12)
    mol, long_bonds, subunits = synthetic.cube_cage(atoms_per_subunit=8, seed=5, spread=6.5)
    base = mol.get_position_matrix()
    owner = np.zeros(base.shape[0], dtype=int)
    for k, ids in subunits.items():
        owner[ids] = k
    intra = [(b.get_atom1_id(), b.get_atom2_id()) for b in mol.get_bonds()]
    states = [_State(base * s + rng.normal(scale=0.01, size=base.shape), owner, intra, list(long_bonds)) for s in (1.0, 1.1, 1.2)]
    mol2, lb2, su2 = synthetic.cube_cage(atoms_per_subunit=5, seed=8, spread=6.0)
    owner2 = np.zeros(mol2.get_num_atoms(), dtype=int)
    for k, ids in su2.items():
        owner2[ids] = k
    states.insert(1, _State(mol2.get_position_matrix(), owner2,
                            [(b.get_atom1_id(), b.get_atom2_id()) for b in mol2.get_bonds()], list(lb2)))
    coll = mch.Collapser(step_size=0.05, distance_threshold=1.5, scale_steps=True)
    results = stk_helpers.collapse_batch(states, coll)
    assert len(results) == len(states)
    for state, (new_state, record) in zip(states, results):
        is_h = [type(a).__name__ == "H" for a in state.get_atoms()]
        want = orc.collapser_run(state.get_position_matrix(), is_h, stk_helpers.get_long_bond_ids(state),
                                 dict(stk_helpers.get_subunits(state)), 0.05, 1.5, True)
        assert isinstance(new_state, _State)
        assert record["steps_run"] == want.steps_run
        assert np.allclose(new_state.get_position_matrix(), want.final_positions, rtol=0, atol=1e-9)
        assert rel_close(record["min_distance"], want.min_distance, RTOL64)
