"""CPU tests (no GPU): data model, lowering, RNG seeding in the C library, the C-ABI
surface, result raising, sharding (incl. a world_size-2 gloo run), and the guarantee that
the product path refuses to run without CUDA instead of falling back to a CPU path."""

from __future__ import annotations

import ctypes as CT
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import mchammer_b200 as mch
from mchammer_b200 import io as mio
from mchammer_b200 import lowering, native, rngstate, stk_helpers, synthetic
from tests.conftest import REPO, golden_bonds, golden_subunits, load_golden


# ------------------------------------------------------------------ data model (reference tests/atom, bond, molecule, utilities)
@pytest.mark.parametrize("idx,el,radius", [(0, "N", 1.4882656711484), (65, "P", 1.925513788), (2, "C", 1.60775485914852)])
def test_atom(idx, el, radius):
    atom = mch.Atom(id=idx, element_string=el)
    assert atom.get_id() == idx and atom.get_element_string() == el
    assert atom.get_radius() == radius == mch.get_radius(el)
    assert str(atom) == f"{el}(id={idx})"


def test_bond_sorts_and_validates():
    bond = mch.Bond(id=3, atom_ids=(9, 0))
    assert (bond.get_id(), bond.get_atom1_id(), bond.get_atom2_id()) == (3, 0, 9)
    assert repr(bond) == "Bond(id=3, atom1_id=0, atom2_id=9)"
    with pytest.raises(ValueError):
        mch.Bond(id=0, atom_ids=(1, 1))


@pytest.fixture()
def toy():
    pos = np.array([[0, 1, 0], [1, 1, 0], [-1, 1, 0], [0, 10, 0], [1, 10, 0], [-1, 10, 0]])
    bonds = [mch.Bond(0, (0, 1)), mch.Bond(1, (0, 2)), mch.Bond(2, (0, 3)), mch.Bond(3, (3, 4)), mch.Bond(4, (3, 5))]
    return mch.Molecule(atoms=[mch.Atom(i, "C") for i in range(6)], bonds=bonds, position_matrix=pos)


def test_molecule_accessors(toy):
    assert toy.get_num_atoms() == 6
    assert toy.get_position_matrix().dtype == np.float64
    assert np.array_equal(toy.get_centroid(), [0, 5.5, 0])
    assert np.array_equal(toy.get_centroid(atom_ids={3, 4, 5}), [0, 10, 0])
    with pytest.raises(ValueError):
        toy.get_centroid(atom_ids=())
    moved = toy.with_displacement(np.array([0, 1, 0]))
    assert np.array_equal(moved.get_position_matrix()[:, 1], toy.get_position_matrix()[:, 1] + 1)
    assert toy.get_subunits(bond_pair_ids=((0, 3),)) == {0: {0, 1, 2}, 1: {3, 4, 5}}
    # the reference tests membership of the SORTED TUPLE: a list entry does not match a tuple
    assert toy.get_subunits(bond_pair_ids=([0, 3],)) == {0: {0, 1, 2, 3, 4, 5}}
    pos = toy.get_position_matrix()
    pos[0, 0] = 99.0
    assert toy.get_position_matrix()[0, 0] == 0.0


def test_primitives(toy):
    pos = toy.get_position_matrix()
    assert np.array_equal(mch.get_bond_vector(pos, (0, 3)), [0, 9, 0])
    assert mch.get_atom_distance(pos, 0, 1) == 1.0 and mch.get_atom_distance(pos, 1, 2) == 2.0
    shifted = mch.translate_atoms_along_vector(toy, (3, 4, 5), np.array([0, 5, 0]))
    assert np.array_equal(shifted.get_position_matrix()[3:, 1], [5, 5, 5])
    back = mch.translate_atoms_along_vector(shifted, (3, 4, 5), np.array([0, -5, 0]))
    assert np.array_equal(back.get_position_matrix(), pos)
    assert mch.test_move(beta=2, curr_pot=-1, new_pot=-2, generator=np.random.default_rng())
    rot = mch.rotation_matrix_arbitrary_axis(np.pi / 2, np.array([0.0, 0.0, 1.0]))
    assert np.allclose(rot @ np.array([1.0, 0, 0]), [0, 1, 0])
    turned = mch.rotate_molecule_by_angle(toy, np.pi, np.array([0.0, 0.0, 1.0]), np.array([0.0, 0, 0]))
    assert np.allclose(turned.get_position_matrix(), pos * [-1, -1, 1])
    assert np.array_equal(mch.translate_molecule_along_vector(toy, np.array([1, 0, 0])).get_position_matrix(), pos + [1, 0, 0])


def test_scalar_potentials_known_answers():
    # reference tests/conftest.py:127-142
    opt = mch.Optimizer(step_size=0.1, target_bond_length=2.0, num_steps=100)
    for d, want in zip(range(1, 8), [50, 0, 50, 200, 450, 800, 1250]):
        assert np.isclose(opt._bond_potential(d), want, atol=1e-5)
    nb = [34.559999999999995, 4.319999999999999, 1.2799999999999998, 0.5399999999999999,
          0.27647999999999995, 0.15999999999999998, 0.10075801749271138]
    for d, want in zip(range(1, 8), nb):
        assert np.isclose(opt._nonbond_potential(d), want, atol=1e-5)


def test_collapser_host_helpers(toy):
    # reference tests/collapser/test_collapser.py:7-66, tests/conftest.py:216-267
    coll = mch.Collapser(step_size=0.05, distance_threshold=1.5, scale_steps=True)
    su = {0: {0, 1, 2}, 1: {3, 4, 5}}
    want = [9, 9.055385138137417, 9.055385138137417, 9.055385138137417, 9, 9.219544457292887,
            9.055385138137417, 9.219544457292887, 9]
    assert np.allclose(list(coll._get_subunit_distances(toy, su)), want, atol=1e-6)
    vec, sc = coll._get_subunit_vectors(toy, su)
    assert np.array_equal(vec[0], [0, -4.5, 0]) and np.array_equal(vec[1], [0, 4.5, 0])
    assert sc == {0: 1.0, 1: 1.0}
    new = coll._get_new_position_matrix(
        toy, su, {0: np.array([2, -1.5, 0]), 1: np.array([0, 2.5, 0])}, {0: 0.2, 1: 1.0}, 1.5)
    assert np.allclose(new, [[-0.6, 1.45, 0], [0.4, 1.45, 0], [-1.6, 1.45, 0], [0, 6.25, 0], [1, 6.25, 0], [-1, 6.25, 0]], atol=1e-6)
    assert not coll._has_short_contacts(toy, su)


def test_writers_match_reference_text():
    g = load_golden("writers")
    b = load_golden("benzene_seed1000")
    mol = mch.Molecule(
        atoms=[mch.Atom(i, str(e)) for i, e in enumerate(b["elements"])],
        bonds=[mch.Bond(i, tuple(int(x) for x in p)) for i, p in enumerate(b["bonds"])],
        position_matrix=b["pos0"],
    )
    assert "".join(mol.write_xyz_content()) == str(g["xyz"])
    assert "".join(mol._write_pdb_content()) == str(g["pdb"])


def test_molfile_round_trip(tmp_path):
    b = load_golden("poc_graph_seed1000")
    mol = mch.Molecule(
        atoms=[mch.Atom(i, str(e)) for i, e in enumerate(b["elements"])],
        bonds=[mch.Bond(i, tuple(int(x) for x in p)) for i, p in enumerate(b["bonds"])],
        position_matrix=b["pos0"],
    )
    path = tmp_path / "poc.mol"
    mio.write_mol(mol, str(path))
    again = mio.read_mol(str(path))
    assert again.get_num_atoms() == 112 and len(list(again.get_bonds())) == 113
    assert np.allclose(again.get_position_matrix(), b["pos0"], atol=5e-5)
    assert [a.get_element_string() for a in again.get_atoms()] == [str(e) for e in b["elements"]]
    v2000 = "\n".join([
        "", " test", "", "  2  1  0  0  0  0  0  0  0  0999 V2000",
        "   -4.1844   -0.5357   -0.0000 C   0  0  0  0  0  0  0  0  0  0  0  0",
        "    4.5677    1.9724    1.0072 Cl  0  0  0  0  0  0  0  0  0  0  0  0",
        "  1  2  1  0  0  0  0", "M  END", ""])
    el, pos, bonds, orders = mio.parse_mol(v2000)
    assert el == ["C", "Cl"] and bonds == [(0, 1)] and orders == [1]
    assert np.allclose(pos, [[-4.1844, -0.5357, 0.0], [4.5677, 1.9724, 1.0072]])


# ------------------------------------------------------------------ lowering
def test_lowering_subunit_order_and_first_owner():
    g = load_golden("poc_bb_seed1000")
    su, lb = golden_subunits(g), golden_bonds(g)
    topo = lowering.lower_topology(112, lb, su)
    assert topo.n_subunits == 10 and topo.max_subunit_atoms == 16 and topo.n_bonds == 12
    assert sorted(topo.perm.tolist()) == list(range(112))
    for k, key in enumerate(su):
        lo, hi = topo.su_offsets[k], topo.su_offsets[k + 1]
        assert topo.perm[lo:hi].tolist() == list(su[key])
    for b, (a1, a2) in enumerate(lb):
        s1 = next(k for k, key in enumerate(su) if a1 in su[key])
        s2 = next(k for k, key in enumerate(su) if a2 in su[key])
        assert topo.bond_su[b].tolist() == [s1, s2]
        assert topo.bonds[b].tolist() == [a1, a2]


def test_lowering_edge_cases():
    su = {"b": [4, 3], "a": (0,)}
    topo = lowering.lower_topology(6, [(3, 0)], su)
    assert topo.n_subunits == 3 and topo.n_user_subunits == 2  # atoms 1, 2, 5 form the static group
    assert topo.perm.tolist() == [4, 3, 0, 1, 2, 5]
    assert topo.su_offsets.tolist() == [0, 2, 3, 6]
    assert topo.bond_su.tolist() == [[0, 1]]
    with pytest.raises(StopIteration):
        lowering.lower_topology(6, [(3, 5)], su)  # reference: bare next() on an empty generator
    assert lowering.lower_topology(6, [(3, 5)], su, require_bond_subunits=False).bond_su.tolist() == [[0, -1]]
    with pytest.raises(ValueError):
        lowering.lower_topology(6, [(3, 0)], {0: [0, 1], 1: [1, 2]})  # overlap
    with pytest.raises(ValueError):
        lowering.lower_topology(6, [(3, 9)], su)
    with pytest.raises(ValueError):
        lowering.lower_topology(6, [(3, 0)], {0: [], 1: [0, 3]})
    heavy = lowering.lower_topology(4, [], {0: [0, 1], 1: [2, 3]}, elements=["H", "C", "N", "H"],
                                    require_bond_subunits=False)
    assert heavy.perm.tolist() == [1, 0, 2, 3] and heavy.n_heavy.tolist() == [1, 1]


# ------------------------------------------------------------------ the C library on the host
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(REPO, "include", "mchammer_b200.h")).read()
    declared = set(re.findall(r"\b(mch_[a-z0-9_]+)\s*\(", header))
    declared -= {"mch_params", "mch_optimize_desc", "mch_collapse_desc"}
    assert declared == set(native.EXPORTS)
    handle = native.lib()
    for name in declared:
        assert getattr(handle, name) is not None
    assert handle.mch_version() >= 100
    out = subprocess.run(["nm", "-D", "--defined-only", native.LIB_PATH], capture_output=True, text=True).stdout
    for name in declared:
        assert re.search(rf"\bT {name}\b", out), name
    # the check build (device asserts + cluster race detector) has the same C ABI
    check = os.path.join(os.path.dirname(native.LIB_PATH), "libmchammer_b200_check.so")
    if os.path.exists(check):
        out = subprocess.run(["nm", "-D", "--defined-only", check], capture_output=True, text=True).stdout
        for name in declared:
            assert re.search(rf"\bT {name}\b", out), name


def test_ctypes_structs_match_header_layout(tmp_path):
    # the C compiler's view of include/mchammer_b200.h against the ctypes mirrors: size of every
    # struct and offset of every field
    structs = {"mch_params": native.MchParams, "mch_optimize_desc": native.OptimizeDesc,
               "mch_collapse_desc": native.CollapseDesc}
    lines = []
    for cname, ctype in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for field, _ in ctype._fields_:
            lines.append(f'printf("{cname}.{field} %zu\\n", offsetof({cname}, {field}));')
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "mchammer_b200.h"\nint main(void) {\n'
                   + "\n".join(lines) + "\nreturn 0; }\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(REPO, "include"), "-o", str(exe), str(src)], check=True)
    seen = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True,
                                                        check=True).stdout.splitlines())
    for cname, ctype in structs.items():
        assert int(seen[cname]) == CT.sizeof(ctype), cname
        for field, _ in ctype._fields_:
            assert int(seen[f"{cname}.{field}"]) == getattr(ctype, field).offset, (cname, field)
    # ... and the header's constants against the Python mirror's
    header = open(os.path.join(REPO, "include", "mchammer_b200.h")).read()
    for name, value in (("MCH_F64", native.MCH_F64), ("MCH_F32", native.MCH_F32),
                        ("MCH_FLAG_EARLY_REJECTION", native.MCH_FLAG_EARLY_REJECTION)):
        assert int(re.search(rf"#define {name} (\d+)", header).group(1)) == value, name


def test_c_seeding_matches_numpy():
    g = load_golden("rng")
    words = rngstate.seed_states(g["seeds"].tolist())
    assert np.array_equal(words[:, :4], g["state_words"])
    assert np.all(words[:, 4:] == 0)
    rng = np.random.default_rng(1)
    seeds = [int(s) for s in rng.integers(0, 2**63, size=200)] + [0, 1, 2**32, 2**64 - 1, 2**80 + 3]
    got = rngstate.seed_states(seeds)
    for s, w in zip(seeds, got):
        assert np.array_equal(w, rngstate.words_from_bit_generator(np.random.PCG64(s)))


def test_c_stream_matches_numpy_generator():
    g = load_golden("rng")
    state = rngstate.seed_states([1000])[0].copy()
    rounds = 400
    ints = np.zeros(3 * rounds, dtype=np.int64)
    dbls = np.zeros(rounds + (rounds + 2) // 3, dtype=np.float64)
    native.check(native.lib().mch_pcg64_stream_check(
        state.ctypes.data_as(CT.c_void_p), rounds, 12, ints.ctypes.data_as(CT.c_void_p),
        dbls.ctypes.data_as(CT.c_void_p)), "stream_check")
    assert np.array_equal(ints, g["mixed_ints"])
    assert np.array_equal(dbls, g["mixed_doubles"])
    # the advanced state can be put back into numpy and both continue identically
    gen = np.random.default_rng(0)
    rngstate.words_to_bit_generator(state, gen.bit_generator)
    ref = np.random.default_rng(1000)
    for k in range(rounds):
        ref.choice(12), ref.choice(2), ref.random(), ref.choice(2)
        if k % 3 == 0:
            ref.random()
    assert gen.random() == ref.random() and int(gen.choice(7)) == int(ref.choice(7))
    # bound 1 draws nothing; odd bounds exercise the rejection branch bookkeeping
    for bound in (1, 2, 3, 24, 48, 65535):
        live = np.random.default_rng(77)
        st = rngstate.words_from_bit_generator(live.bit_generator).copy()
        ints = np.zeros(3 * 50, dtype=np.int64)
        dbls = np.zeros(50 + 17, dtype=np.float64)
        native.check(native.lib().mch_pcg64_stream_check(
            st.ctypes.data_as(CT.c_void_p), 50, bound, ints.ctypes.data_as(CT.c_void_p),
            dbls.ctypes.data_as(CT.c_void_p)), "stream_check")
        want_i, want_d = [], []
        for k in range(50):
            want_i += [int(live.choice(bound)), int(live.choice(2))]
            want_d.append(live.random())
            want_i.append(int(live.choice(2)))
            if k % 3 == 0:
                want_d.append(live.random())
        assert ints.tolist() == want_i and dbls.tolist() == want_d


def test_smem_query_and_limits():
    handle = native.lib()
    small = handle.mch_optimize_smem_bytes(1000, 20, 24, 50, native.MCH_F32)
    big = handle.mch_optimize_smem_bytes(5004, 36, 48, 208, native.MCH_F64)
    assert 0 < small < 32 * 1024
    assert small < big <= 227 * 1024
    assert handle.mch_optimize_smem_bytes(20000, 36, 48, 208, native.MCH_F64) == -2
    assert b"shared memory" in handle.mch_last_error()
    assert handle.mch_optimize_smem_bytes(0, 1, 1, 1, 0) == -1
    assert handle.mch_optimize_batch(None, None) == -1
    assert handle.mch_collapse_batch(None, None) == -1


def test_product_path_refuses_to_run_without_cuda(toy):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    opt = mch.Optimizer(step_size=0.1, target_bond_length=2.0, num_steps=10)
    su = toy.get_subunits(bond_pair_ids=((0, 3),))
    with pytest.raises(native.NativeLibraryError):
        opt.get_result(toy, ((0, 3),), su)
    with pytest.raises(native.NativeLibraryError):
        opt.compute_potential(toy, ((0, 3),))
    with pytest.raises(native.NativeLibraryError):
        opt.get_results_batch(toy, ((0, 3),), su, seeds=[1, 2])
    with pytest.raises(native.NativeLibraryError):
        mch.Collapser(0.05, 1.5).get_result(toy, ((0, 3),), su)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(REPO, "mchammer_b200")
    for root, _, files in os.walk(pkg):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, name)).read()
                assert "oracle" not in text.replace("mch_oracle", "oracle") or name == "__never__", name


# ------------------------------------------------------------------ stk adapters (duck typed)
class _A:
    def __init__(self, i): self.i = i
    def get_id(self): return self.i


class C(_A): pass  # noqa: E701  element name comes from the class name (stk: atom class == element)


class _B:
    def __init__(self, a, b): self.a, self.b = _A(a), _A(b)
    def get_atom1(self): return self.a
    def get_atom2(self): return self.b


class _BI:
    def __init__(self, a, b, bb): self.b, self.bb = _B(a, b), bb
    def get_bond(self): return self.b
    def get_building_block(self): return self.bb


class _AI:
    def __init__(self, i, bb): self.a, self.bb = _A(i), bb
    def get_atom(self): return self.a
    def get_building_block_id(self): return self.bb


class _State:
    def get_bond_infos(self): return [_BI(1, 0, "x"), _BI(3, 1, None), _BI(2, 3, "y")]
    def get_atom_infos(self): return [_AI(0, 0), _AI(1, 0), _AI(2, 1), _AI(3, 1)]
    def get_atoms(self): return [C(i) for i in range(4)]
    def get_position_matrix(self): return np.arange(12.0).reshape(4, 3)


def test_stk_helpers():
    st = _State()
    assert stk_helpers.get_long_bond_ids(st) == ((1, 3),)
    assert dict(stk_helpers.get_subunits(st)) == {0: [0, 1], 1: [2, 3]}
    bonds = list(stk_helpers.get_mch_bonds(st))
    assert [(b.get_atom1_id(), b.get_atom2_id()) for b in bonds] == [(0, 1), (1, 3), (2, 3)]
    mol = stk_helpers.molecule_from_stk(st)
    assert mol.get_num_atoms() == 4 and [a.get_element_string() for a in mol.get_atoms()] == ["C"] * 4
    assert len(list(mol.get_bonds())) == 3


# ------------------------------------------------------------------ synthetic workloads
def test_synthetic_cages_are_deterministic_and_well_formed():
    mol, lb, su = synthetic.cube_cage()
    mol2, lb2, _ = synthetic.cube_cage()
    assert np.array_equal(mol.get_position_matrix(), mol2.get_position_matrix()) and lb == lb2
    assert mol.get_num_atoms() == 1000 and len(su) == 20 and len(lb) == 24
    assert {len(v) for v in su.values()} == {50}
    pos = mol.get_position_matrix()
    lengths = [np.linalg.norm(pos[a] - pos[b]) for a, b in lb]
    assert 3.5 < min(lengths) and max(lengths) < 7.5
    found = mol.get_subunits(bond_pair_ids=lb)
    assert sorted(map(sorted, found.values())) == sorted(map(sorted, su.values()))
    big, lb5, su5 = synthetic.cuboctahedron_cage()
    assert big.get_num_atoms() == 5004 and len(su5) == 36 and len(lb5) == 48


# ------------------------------------------------------------------ sharding
def test_shard_bounds():
    for n, g in ((4096, 8), (10, 3), (3, 8), (0, 2), (65536, 5)):
        bounds = mch.shard_bounds(n, g)
        assert bounds[0][0] == 0 and bounds[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(bounds, bounds[1:]))
        sizes = [hi - lo for lo, hi in bounds]
        assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_sharding_and_gather(tmp_path):
    """world_size-2 run on CPU: each rank takes its shard of the chains, 'processes' it with a
    stand-in that tags results with the chain's seed, and rank 0 gathers -- the same
    plumbing bench.py uses around the CUDA call (no collective on the data path)."""
    script = tmp_path / "worker.py"
    script.write_text(
        "import os, sys\n"
        f"sys.path.insert(0, {REPO!r})\n"
        "import numpy as np, torch, torch.distributed as dist\n"
        "from mchammer_b200 import shard_bounds\n"
        "dist.init_process_group('gloo')\n"
        "rank, world = dist.get_rank(), dist.get_world_size()\n"
        "seeds = np.arange(1000, 1011)\n"
        "lo, hi = shard_bounds(len(seeds), world)[rank]\n"
        "mine = torch.tensor(seeds[lo:hi] * 2.0)\n"
        "steps = torch.tensor([float(hi - lo)])\n"
        "dist.all_reduce(steps, op=dist.ReduceOp.SUM)\n"
        "t = torch.tensor([0.5 + rank])\n"
        "dist.all_reduce(t, op=dist.ReduceOp.MAX)\n"
        "parts = [None] * world\n"
        "dist.all_gather_object(parts, (lo, hi, mine.tolist()))\n"
        "if rank == 0:\n"
        "    full = np.zeros(len(seeds))\n"
        "    for lo_, hi_, vals in parts: full[lo_:hi_] = vals\n"
        "    assert np.array_equal(full, seeds * 2.0), full\n"
        "    assert steps.item() == len(seeds) and t.item() == 1.5\n"
        "    print('GATHER_OK')\n"
        "dist.destroy_process_group()\n"
    )
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    proc = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)],
        capture_output=True, text=True, env=env, timeout=240,
    )
    assert proc.returncode == 0, proc.stderr[-2000:]
    assert "GATHER_OK" in proc.stdout


def test_bench_reference_arm_prints_the_contract_line():
    # `bench.py --impl reference` (the CPU arm the driver runs beside the CUDA arm): one JSON line with the
    # contract's keys, the unmodified reference when baseline/_ref is present, no GPU touched
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--workload", "cube1000"], capture_output=True, text=True, env=env,
                         timeout=600, cwd=REPO)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "mc_chain_steps_per_s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["unit"] == "chain-steps/s" and line["n_gpus"] == 1
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    base = line["cpu_baseline"]
    assert base["cores"] >= 1 and base["kind"] in ("reference", "port")
    if os.path.isfile(os.path.join(REPO, "baseline", "_ref", "mchammer", "__init__.py")):
        assert base["kind"] == "reference"


def test_early_rejection_default_policy_and_environment_override(monkeypatch):
    from mchammer_b200 import engine as _engine

    assert _engine.resolve_early_reject(None, "fp64") is _engine.EARLY_REJECT_DEFAULT["fp64"]
    assert _engine.resolve_early_reject(None, "fp32") is _engine.EARLY_REJECT_DEFAULT["fp32"]
    assert _engine.resolve_early_reject(False, "fp64") is False and _engine.resolve_early_reject(True, "fp32") is True
    monkeypatch.setenv("MCHAMMER_B200_EARLY", "1")
    assert _engine.resolve_early_reject(None, "fp32") is True
    monkeypatch.setenv("MCHAMMER_B200_EARLY", "0")
    assert _engine.resolve_early_reject(None, "fp64") is False and _engine.resolve_early_reject(True, "fp64") is True
