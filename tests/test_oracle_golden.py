"""Pin the CPU oracle (oracle/mch_oracle.py) to outputs of the unmodified reference.

Golden vectors: tests/golden/*.npz, produced by tests/golden/make_golden.py from the
reference at /root/reference.  Known-answer literals: reference tests/conftest.py:127-152.
Bit-exact (``==``) wherever the reference is bit-reproducible on this numpy/scipy.
"""

from __future__ import annotations

import numpy as np
import pytest

from oracle import mch_oracle as orc
from tests.conftest import golden_bonds, golden_subunits, load_golden

OPT_CASES = [
    "benzene_seed1000", "benzene_seed7", "benzene_mu2p5", "benzene_mu6",
    "poc_graph_seed1000", "poc_graph_seed7", "poc_graph_seed42", "poc_bb_seed1000",
    "toy_seed1000",
    # round 2: overlapping subunits / an id listed twice / an atom in no subunit (reference semantics)
    "benzene_overlap_seed9", "benzene_dup_seed4", "poc_overlap_seed3",
]


def params_of(g) -> orc.OptimizerParams:
    return orc.OptimizerParams(
        step_size=float(g["step_size"]), target_bond_length=float(g["target_bond_length"]),
        num_steps=int(g["num_steps"]), bond_epsilon=float(g["bond_epsilon"]),
        nonbond_epsilon=float(g["nonbond_epsilon"]), nonbond_sigma=float(g["nonbond_sigma"]),
        nonbond_mu=float(g["nonbond_mu"]), beta=float(g["beta"]),
    )


def test_known_answer_potentials():
    # reference tests/conftest.py:54-65, :127-152 ; tests/optimizer/test_optimizer.py:19-62
    p = orc.OptimizerParams(step_size=0.1, target_bond_length=2.0, num_steps=100)
    pos = np.array([[0, 1, 0], [1, 1, 0], [-1, 1, 0], [0, 10, 0], [1, 10, 0], [-1, 10, 0]])
    for d, want in zip(range(1, 8), [50, 0, 50, 200, 450, 800, 1250]):
        assert np.isclose(orc.bond_potential(d, p), want, atol=1e-5)
    nb = [34.559999999999995, 4.319999999999999, 1.2799999999999998, 0.5399999999999999,
          0.27647999999999995, 0.15999999999999998, 0.10075801749271138]
    for d, want in zip(range(1, 8), nb):
        assert np.isclose(orc.nonbond_potential(d, p), want, atol=1e-5)
    assert orc.nonbonded_potential(pos, p) == 147.2965949864993
    u, unb = orc.compute_potential(pos, ((0, 3),), p)
    assert unb == 147.2965949864993
    assert u == 2597.2965949864993


def test_potentials_random_geometries():
    g = load_golden("potentials")
    p = orc.OptimizerParams(step_size=0.1, target_bond_length=2.0, num_steps=2)
    for n in (2, 3, 17, 112, 300):
        u, unb = orc.compute_potential(g[f"pos_{n}"], g[f"bonds_{n}"].tolist(), p)
        assert u == float(g[f"u_{n}"])
        assert unb == float(g[f"unb_{n}"])


@pytest.mark.parametrize("case", OPT_CASES)
def test_optimizer_trace_bit_exact(case):
    g = load_golden(case)
    tr = orc.optimizer_run(
        g["pos0"], golden_bonds(g), golden_subunits(g), params_of(g),
        np.random.default_rng(int(g["seed"])), keep_positions="positions" in g.files,
    )
    assert np.array_equal(tr.passed, g["passed"])
    assert np.array_equal(tr.bond_index, g["bond_index"])
    assert np.array_equal(tr.system_potential, g["system_potential"])
    assert np.array_equal(tr.nonbonded_potential, g["nonbonded_potential"])
    assert np.array_equal(tr.max_bond_distance, g["max_bond_distance"])
    assert np.array_equal(tr.final_positions, g["final_positions"])
    if "positions" in g.files:
        assert np.array_equal(tr.positions, g["positions"])


def test_benzene_against_committed_example_log():
    # reference examples/benzene_opt.out:26-125 (written by the author's numpy build:
    # identical decisions, potentials to <= 1 ulp-level)
    ex = load_golden("reference_examples")
    g = load_golden("benzene_seed1000")
    rows = [ln.split() for ln in str(ex["benzene_opt_out"]).splitlines()
            if ln[:1].isdigit() and len(ln.split()) >= 6]
    assert len(rows) == 100
    bonds = golden_bonds(g)
    for k, row in enumerate(rows):
        assert int(row[0]) == k
        assert np.isclose(float(row[1]), g["system_potential"][k], rtol=1e-13)
        assert np.isclose(float(row[2]), g["nonbonded_potential"][k], rtol=1e-13)
        assert np.isclose(float(row[3]), g["max_bond_distance"][k], rtol=1e-13)
        if k:
            text = " ".join(row[4:-1]).strip("[]").split()
            assert bonds[int(g["bond_index"][k])] == (int(text[0]), int(text[1]))
            assert row[-1] == ("T" if g["passed"][k] else "F")
    assert "23 steps passed: 23.0%" in str(ex["benzene_opt_out"])
    assert int(g["number_passed"]) == 23


def test_final_positions_against_committed_example_molfiles():
    # reference examples/benzene_opt.mol, poc_opt.mol, poc_opt_nci.mol (4-decimal files)
    ex = load_golden("reference_examples")
    for case, name in (("benzene_seed1000", "benzene_opt"), ("poc_graph_seed1000", "poc_opt"),
                       ("poc_bb_seed1000", "poc_opt_nci")):
        g = load_golden(case)
        assert np.abs(g["final_positions"] - ex[name]).max() < 1.5e-4


@pytest.mark.parametrize("case", ["collapser_poc", "collapser_poc_noscale", "collapser_toy", "collapser_poc_loose"])
def test_collapser_bit_exact(case):
    g = load_golden(case)
    is_h = [e == "H" for e in g["elements"]]
    tr = orc.collapser_run(
        g["pos0"], is_h, golden_bonds(g), golden_subunits(g), float(g["step_size"]),
        float(g["distance_threshold"]), bool(g["scale_steps"]), keep_positions=True,
    )
    assert tr.steps_run == int(g["steps_run"])
    assert np.array_equal(tr.max_bond_distance, g["max_bond_distance"])
    assert np.array_equal(tr.positions, g["positions"])
    assert np.array_equal(tr.final_positions, g["final_positions"])
    assert np.isclose(tr.min_distance, float(g["min_distance_logged"]), rtol=1e-15, atol=0)


def test_collapser_known_answers():
    # reference tests/collapser/test_collapser.py:89,129 ; tests/conftest.py:288-299
    g = load_golden("collapser_toy")
    assert int(g["steps_run"]) == 35
    assert np.isclose(float(g["final_min_distance"]), 1.4947505, atol=1e-8)
    want = np.array([[0.0, 4.75262477, 0.0], [1.0, 4.75262477, 0.0], [-1.0, 4.75262477, 0.0],
                     [0.0, 6.24737523, 0.0], [1.0, 6.24737523, 0.0], [-1.0, 6.24737523, 0.0]])
    assert np.allclose(g["final_positions"], want, atol=1e-8)
    # reference examples/coll_opt.out:28 -- 13 steps on poc.mol
    assert int(load_golden("collapser_poc")["steps_run"]) == 13
    ex = load_golden("reference_examples")
    assert "Steps run: 13" in str(ex["coll_opt_out"])
    assert np.abs(load_golden("collapser_poc")["final_positions"] - ex["coll_poc_opt"]).max() < 1.5e-4


def test_collapser_subunit_distance_order():
    # reference tests/conftest.py:216-229, tests/collapser/test_collapser.py:7-22
    pos = np.array([[0, 1, 0], [1, 1, 0], [-1, 1, 0], [0, 10, 0], [1, 10, 0], [-1, 10, 0]], float)
    want = [9, 9.055385138137417, 9.055385138137417, 9.055385138137417, 9,
            9.219544457292887, 9.055385138137417, 9.219544457292887, 9]
    got = list(orc.subunit_distances(pos, [False] * 6, {0: {0, 1, 2}, 1: {3, 4, 5}}))
    assert np.allclose(got, want, atol=1e-6)
    vec, sc = orc.subunit_vectors(pos, {0: {0, 1, 2}, 1: {3, 4, 5}}, True)
    assert np.array_equal(vec[0], [0, -4.5, 0]) and np.array_equal(vec[1], [0, 4.5, 0])
    assert sc == {0: 1.0, 1: 1.0}


def test_pcg64_model_matches_numpy_stream():
    g = load_golden("rng")
    for seed, words, raw in zip(g["seeds"], g["state_words"], g["raw"]):
        m = orc.PCG64Model.from_seed(int(seed))
        assert m.state == (int(words[0]) << 64 | int(words[1]))
        assert m.inc == (int(words[2]) << 64 | int(words[3]))
        assert [m.next64() for _ in range(8)] == [int(x) for x in raw]
    m = orc.PCG64Model.from_seed(1000)
    ints, dbls = [], []
    for k in range(400):
        ints += [m.bounded(12), m.bounded(2)]
        dbls.append(m.random())
        ints.append(m.bounded(2))
        if k % 3 == 0:
            dbls.append(m.random())
    assert ints == g["mixed_ints"].tolist()
    assert dbls == g["mixed_doubles"].tolist()


def test_pcg64_model_matches_live_numpy():
    for seed in (0, 5, 123456789012345678901234567890):
        bg = np.random.PCG64(seed)
        m = orc.PCG64Model.from_seed(seed)
        st = bg.state["state"]
        assert (m.state, m.inc) == (st["state"], st["inc"])
        gen = np.random.Generator(bg)
        for n in (1, 2, 3, 12, 24, 48, 1000, 2**31 + 7):
            assert int(gen.choice(n)) == m.bounded(n)
            assert gen.random() == m.random()
