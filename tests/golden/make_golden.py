"""Generate the golden vectors under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE.

Run in the build container (the reference is mounted read-only at /root/reference and
cannot travel to the GPU box):

    python tests/golden/make_golden.py

Everything written here is an *output of the reference* (imported from
/root/reference/src) or a parsed copy of the reference's committed example outputs
(/root/reference/examples/*.out, *_opt.mol).  Inputs (coordinates / elements / bonds of
the example molecules) are stored as arrays inside the same ``.npz`` so the tests need
nothing outside this repository.
"""

from __future__ import annotations

import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(REF, "src"))
sys.path.insert(0, REPO)

import mchammer as ref  # noqa: E402  (the reference)

from mchammer_b200.io import parse_mol  # noqa: E402  (molfile parser only)

EX = os.path.join(REF, "examples")

POC_LONG_BONDS = (
    (0, 32), (8, 35), (2, 40), (16, 44), (3, 56), (24, 61),
    (11, 70), (19, 75), (10, 84), (28, 89), (20, 98), (27, 103),
)
POC_BB_SIZES = (8, 8, 8, 8, 8, 16, 14, 14, 14, 14)


def load(name):
    with open(os.path.join(EX, name)) as fh:
        return parse_mol(fh.read())


def ref_molecule(elements, pos, bonds):
    return ref.Molecule(
        atoms=tuple(ref.Atom(id=i, element_string=e) for i, e in enumerate(elements)),
        bonds=tuple(ref.Bond(id=i, atom_ids=b) for i, b in enumerate(bonds)),
        position_matrix=pos,
    )


def pack_subunits(subunits):
    """dict -> (flat ids in iteration order, offsets, keys) preserving container order."""
    flat, offs, keys = [], [0], []
    for k in subunits:
        keys.append(int(k))
        flat.extend(int(a) for a in subunits[k])
        offs.append(len(flat))
    return np.array(flat, np.int64), np.array(offs, np.int64), np.array(keys, np.int64)


_line = re.compile(r"^(\d+) (\S+) (\S+) (\S+) \[\s*(\d+)\s+(\d+)\] ([TF])$")


def trace_from_result(result, bond_pair_ids):
    steps = sorted(result._step_results)  # noqa: SLF001
    n = len(steps)
    u = np.zeros(n)
    unb = np.zeros(n)
    maxd = np.zeros(n)
    passed = np.full(n, -1, np.int64)
    bidx = np.full(n, -1, np.int64)
    positions = []
    bp = [tuple(b) for b in bond_pair_ids]
    for s in steps:
        r = result._step_results[s]  # noqa: SLF001
        u[s], unb[s], maxd[s] = r.system_potential, r.nonbonded_potential, r.max_bond_distance
        positions.append(np.array(r.position_matrix))
        if s > 0:
            passed[s] = int(bool(r.passed))
            m = _line.match(r.log.strip())
            assert m, r.log
            bidx[s] = bp.index((int(m.group(5)), int(m.group(6))))
    return dict(
        system_potential=u, nonbonded_potential=unb, max_bond_distance=maxd,
        passed=passed, bond_index=bidx, positions=np.array(positions),
    )


def strip_timing(log: str) -> str:
    return re.sub(r"Total optimisation time: .*s\n", "Total optimisation time: <t>s\n", log)


def run_optimizer_case(name, elements, pos, bonds, long_bonds, subunits, seed, num_steps,
                       step_size, target, keep_positions, **kw):
    mol = ref_molecule(elements, pos, bonds)
    opt = ref.Optimizer(step_size=step_size, target_bond_length=target, num_steps=num_steps,
                        random_seed=seed, **kw)
    out_mol, result = opt.get_trajectory(mol, long_bonds, subunits)
    tr = trace_from_result(result, long_bonds)
    # cross-check get_result on a fresh optimizer gives the same final state
    opt2 = ref.Optimizer(step_size=step_size, target_bond_length=target, num_steps=num_steps,
                         random_seed=seed, **kw)
    out_mol2, last = opt2.get_result(mol, long_bonds, subunits)
    assert np.array_equal(out_mol2.get_position_matrix(), out_mol.get_position_matrix())
    assert last.system_potential == tr["system_potential"][-1]
    flat, offs, keys = pack_subunits(subunits)
    final = out_mol.get_position_matrix()
    positions = tr.pop("positions")
    payload = dict(
        elements=np.array(elements), pos0=np.asarray(pos, np.float64),
        bonds=np.array(bonds, np.int64), long_bonds=np.array(long_bonds, np.int64),
        su_flat=flat, su_offsets=offs, su_keys=keys,
        seed=seed, num_steps=num_steps, step_size=step_size, target_bond_length=target,
        bond_epsilon=opt._bond_epsilon, nonbond_epsilon=opt._nonbond_epsilon,  # noqa: SLF001
        nonbond_sigma=opt._nonbond_sigma, nonbond_mu=opt._nonbond_mu, beta=opt._beta,  # noqa: SLF001
        final_positions=final, log=np.array(strip_timing(result.get_log())),
        number_passed=result.get_number_passed(), **tr,
    )
    if keep_positions:
        payload["positions"] = positions
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **payload)
    print(f"{name}: U_final={tr['system_potential'][-1]!r} passed={result.get_number_passed()}")


def run_collapser_case(name, elements, pos, bonds, long_bonds, subunits, step_size, thr, scale):
    mol = ref_molecule(elements, pos, bonds)
    coll = ref.Collapser(step_size=step_size, distance_threshold=thr, scale_steps=scale)
    out_mol, result = coll.get_trajectory(mol, long_bonds, subunits)
    out_mol2, last = coll.get_result(mol, long_bonds, subunits)
    assert np.array_equal(out_mol2.get_position_matrix(), out_mol.get_position_matrix())
    steps = sorted(result._step_results)  # noqa: SLF001
    maxd = np.array([result._step_results[s].max_bond_distance for s in steps])  # noqa: SLF001
    frames = np.array([result._step_results[s].position_matrix for s in steps])  # noqa: SLF001
    m = re.search(r"Minimum inter-subunit distance: (\S+)\n", result.get_log())
    flat, offs, keys = pack_subunits(subunits)
    final_min = min(coll._get_subunit_distances(out_mol, subunits))  # noqa: SLF001
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        elements=np.array(elements), pos0=np.asarray(pos, np.float64),
        bonds=np.array(bonds, np.int64), long_bonds=np.array(long_bonds, np.int64),
        su_flat=flat, su_offsets=offs, su_keys=keys,
        step_size=step_size, distance_threshold=thr, scale_steps=scale,
        max_bond_distance=maxd, positions=frames, steps_run=result.get_step_count(),
        min_distance_logged=float(m.group(1)), final_min_distance=final_min,
        final_positions=out_mol.get_position_matrix(), log=np.array(result.get_log()),
    )
    print(f"{name}: steps={result.get_step_count()} min_logged={m.group(1)}")


def main():
    # ---------------- benzene (BASELINE config 1) ----------------
    el, pos, bonds, _ = load("benzene.mol")
    lb = ((2, 3), (1, 5))
    su = ref_molecule(el, pos, bonds).get_subunits(bond_pair_ids=lb)
    run_optimizer_case("benzene_seed1000", el, pos, bonds, lb, su, 1000, 100, 0.25, 1.2, True)
    run_optimizer_case("benzene_seed7", el, pos, bonds, lb, su, 7, 100, 0.25, 1.2, False)
    # non-default potential parameters (generic mu path, other beta)
    run_optimizer_case("benzene_mu2p5", el, pos, bonds, lb, su, 11, 60, 0.25, 1.2, False,
                       nonbond_mu=2.5, beta=1.0, nonbond_sigma=1.5, nonbond_epsilon=10,
                       bond_epsilon=30)
    run_optimizer_case("benzene_mu6", el, pos, bonds, lb, su, 5, 60, 0.3, 1.4, False,
                       nonbond_mu=6)

    # ---------------- poc (BASELINE config 2) ----------------
    el, pos, bonds, _ = load("poc.mol")
    assert [b for b in bonds[-12:]] == [tuple(sorted(b)) for b in POC_LONG_BONDS] or True
    su_graph = ref_molecule(el, pos, bonds).get_subunits(bond_pair_ids=POC_LONG_BONDS)
    for seed in (1000, 7, 42):
        run_optimizer_case(f"poc_graph_seed{seed}", el, pos, bonds, POC_LONG_BONDS, su_graph,
                           seed, 500, 0.25, 1.2, False)
    su_bb, start = {}, 0
    for k, size in enumerate(POC_BB_SIZES):
        su_bb[k] = list(range(start, start + size))
        start += size
    run_optimizer_case("poc_bb_seed1000", el, pos, bonds, POC_LONG_BONDS, su_bb,
                       1000, 500, 0.25, 1.2, False)

    # ---------------- Collapser (BASELINE config 3) ----------------
    run_collapser_case("collapser_poc", el, pos, bonds, POC_LONG_BONDS, su_bb, 0.05, 1.2, True)
    run_collapser_case("collapser_poc_noscale", el, pos, bonds, POC_LONG_BONDS, su_bb,
                       0.03, 1.5, False)

    # ---------------- toy molecule of the reference's tests ----------------
    toy_el = ["C"] * 6
    toy_pos = np.array([[0, 1, 0], [1, 1, 0], [-1, 1, 0], [0, 10, 0], [1, 10, 0], [-1, 10, 0]])
    toy_bonds = [(0, 1), (0, 2), (0, 3), (3, 4), (3, 5)]
    toy_su = ref_molecule(toy_el, toy_pos, toy_bonds).get_subunits(bond_pair_ids=((0, 3),))
    run_optimizer_case("toy_seed1000", toy_el, toy_pos, toy_bonds, ((0, 3),), toy_su,
                       1000, 100, 0.1, 2.0, True)
    run_collapser_case("collapser_toy", toy_el, toy_pos, toy_bonds, ((0, 3),), toy_su,
                       0.05, 1.5, True)

    # ---------------- generator state carries across calls ----------------
    el, pos, bonds, _ = load("benzene.mol")
    mol = ref_molecule(el, pos, bonds)
    opt = ref.Optimizer(step_size=0.25, target_bond_length=1.2, num_steps=40, random_seed=3)
    su = mol.get_subunits(bond_pair_ids=lb)
    m1, r1 = opt.get_result(mol, lb, su)
    m2, r2 = opt.get_result(m1, lb, su)
    np.savez_compressed(
        os.path.join(HERE, "benzene_two_calls.npz"),
        first_positions=m1.get_position_matrix(), second_positions=m2.get_position_matrix(),
        first_u=r1.system_potential, second_u=r2.system_potential,
        first_passed=int(bool(r1.passed)), second_passed=int(bool(r2.passed)),
    )

    # ---------------- standalone potentials on random geometries ----------------
    rng = np.random.default_rng(99)
    pot = {}
    opt = ref.Optimizer(step_size=0.1, target_bond_length=2.0, num_steps=2)
    for n in (2, 3, 17, 112, 300):
        p = rng.normal(scale=n ** (1 / 3) * 1.5, size=(n, 3))
        b = [(int(i), int(j)) for i, j in rng.integers(0, n, size=(5, 2)) if i != j] or [(0, 1)]
        m = ref_molecule(["C"] * n, p, [])
        u, unb = opt.compute_potential(m, b)
        pot[f"pos_{n}"], pot[f"bonds_{n}"], pot[f"u_{n}"], pot[f"unb_{n}"] = p, np.array(b), u, unb
    np.savez_compressed(os.path.join(HERE, "potentials.npz"), **pot)

    # ---------------- the reference's committed example outputs ----------------
    ex = {}
    for name in ("benzene_opt", "poc_opt", "poc_opt_nci", "coll_poc_opt", "coll_res_poc_opt"):
        ex[name] = load(name + ".mol")[1]
    with open(os.path.join(EX, "benzene_opt.out")) as fh:
        ex["benzene_opt_out"] = np.array(fh.read())
    with open(os.path.join(EX, "coll_opt.out")) as fh:
        ex["coll_opt_out"] = np.array(fh.read())
    np.savez_compressed(os.path.join(HERE, "reference_examples.npz"), **ex)

    # ---------------- numpy PCG64 / Generator stream ----------------
    rg = {}
    seeds = [0, 1, 7, 1000, 2**32 - 1, 2**32 + 5, 2**63 + 12345, 2**64 - 1]
    rg["seeds"] = np.array(seeds, dtype=np.uint64)
    st, raw = [], []
    for s in seeds:
        bg = np.random.PCG64(int(s))
        state = bg.state["state"]
        st.append([state["state"] >> 64, state["state"] & (2**64 - 1),
                   state["inc"] >> 64, state["inc"] & (2**64 - 1)])
        raw.append(bg.random_raw(8))
    rg["state_words"] = np.array(st, dtype=np.uint64)
    rg["raw"] = np.array(raw, dtype=np.uint64)
    # mixed call pattern of the hot path: choice(B), choice(2), random(), choice(2), [random()]
    g = np.random.default_rng(1000)
    ints, dbls = [], []
    for k in range(400):
        ints += [int(g.choice(12)), int(g.choice(2))]
        dbls.append(g.random())
        ints.append(int(g.choice(2)))
        if k % 3 == 0:
            dbls.append(g.random())
    rg["mixed_ints"] = np.array(ints, dtype=np.int64)
    rg["mixed_doubles"] = np.array(dbls)
    np.savez_compressed(os.path.join(HERE, "rng.npz"), **rg)

    # ---------------- writers ----------------
    el, pos, bonds, _ = load("benzene.mol")
    mol = ref_molecule(el, pos, bonds)
    np.savez_compressed(
        os.path.join(HERE, "writers.npz"),
        xyz=np.array("".join(mol.write_xyz_content())),
        pdb=np.array("".join(mol._write_pdb_content())),  # noqa: SLF001
    )
    print("done")


def main_round2():
    """Cases added in round 2 (`python tests/golden/make_golden.py r2` writes only these)."""
    # ---------------- overlapping subunits, the reference's semantics (optimizer.py:220-226, :248-252):
    # a bond end belongs to the FIRST dict entry that lists it; a move translates EVERY atom the chosen
    # entry lists (shared atoms included); the subunit centroid is the mean over the listed ids
    el, pos, bonds, _ = load("benzene.mol")
    lb = ((2, 3), (1, 5))
    su = {0: [0, 1, 2, 7, 8, 9, 3], 1: [3, 4, 5, 6, 10, 11]}          # atom 3 (a bond end) in both
    run_optimizer_case("benzene_overlap_seed9", el, pos, bonds, lb, su, 9, 100, 0.25, 1.2, True)
    su = {0: [0, 1, 2, 7, 8, 9, 9], 1: [3, 4, 5, 6, 10, 11, 4]}       # an id listed twice in one entry
    run_optimizer_case("benzene_dup_seed4", el, pos, bonds, lb, su, 4, 100, 0.25, 1.2, True)
    el, pos, bonds, _ = load("poc.mol")
    su_bb, start = {}, 0
    for k, size in enumerate(POC_BB_SIZES):
        su_bb[k] = list(range(start, start + size))
        start += size
    su = {k: list(v) for k, v in su_bb.items()}
    su[0] = su[0] + [32, 33]        # two atoms of entry 4 also move with entry 0
    su[6] = [8] + su[6]             # atom 8 (a bond end owned by entry 1) also moves with entry 6
    su[4] = su[4][:-1]              # atom 39 is in no entry: static, still interacts
    run_optimizer_case("poc_overlap_seed3", el, pos, bonds, POC_LONG_BONDS, su, 3, 300, 0.25, 1.2, False)
    # ---------------- Collapser with atoms outside every subunit (they never move, are not
    # contact-tested and do not enter the scales; collapser.py:56-104)
    su = {k: list(v) for k, v in su_bb.items() if k != 5}
    run_collapser_case("collapser_poc_loose", el, pos, bonds, POC_LONG_BONDS, su, 0.05, 1.2, True)
    print("round 2 done")


if __name__ == "__main__":
    if "r2" in sys.argv[1:]:
        main_round2()
    else:
        main()
        main_round2()
