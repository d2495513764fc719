"""Duck-typed adapters for stk ``ConstructedMolecule`` / ``ConstructionState`` objects.

They only call ``get_atoms() / get_bonds() / get_bond_infos() / get_atom_infos()`` and the
per-item accessors, so stk itself is not a dependency.  Behaviour follows the helpers the
reference ships in its examples and tests (examples/stk_example.py:8-43,
tests/stk_reproduction/utilities.py:7-70) and the wiring of tests/stk_reproduction/test_stk.py:97-135.
"""

from __future__ import annotations

from collections import defaultdict

from .structure import Atom, Bond, Molecule


def _bond_ends(bond_info) -> tuple[int, int]:
    bond = bond_info.get_bond()
    return bond.get_atom1().get_id(), bond.get_atom2().get_id()


def get_mch_bonds(state):
    """Yield a :class:`Bond` (sorted ids, id = running index) per stk bond info."""
    for index, info in enumerate(state.get_bond_infos()):
        yield Bond(id=index, atom_ids=tuple(sorted(_bond_ends(info))))


def get_long_bond_ids(state) -> tuple[tuple[int, int], ...]:
    """Sorted id pairs of the bonds created by construction (``get_building_block() is None``)."""
    return tuple(
        tuple(sorted(_bond_ends(info)))
        for info in state.get_bond_infos()
        if info.get_building_block() is None
    )


def get_subunits(state) -> dict:
    """``{building_block_id: [atom ids]}`` in atom order (stk's notion of a rigid subunit)."""
    groups: dict = defaultdict(list)
    for info in state.get_atom_infos():
        groups[info.get_building_block_id()].append(info.get_atom().get_id())
    return groups


def molecule_from_stk(state) -> Molecule:
    """Build a :class:`Molecule` from an stk molecule/state (element = atom class name)."""
    atoms = tuple(
        Atom(id=a.get_id(), element_string=type(a).__name__) for a in state.get_atoms()
    )
    if hasattr(state, "get_bond_infos"):
        bonds = tuple(get_mch_bonds(state))
    else:
        bonds = tuple(
            Bond(id=i, atom_ids=(b.get_atom1().get_id(), b.get_atom2().get_id()))
            for i, b in enumerate(state.get_bonds())
        )
    return Molecule(atoms=atoms, bonds=bonds, position_matrix=state.get_position_matrix())


def optimize_batch(states, optimizer, seeds=1000, devices=None):
    """Optimise many stk-shaped molecules/states in batched launches (NEW; no reference counterpart).

    Per state this is the wiring of reference tests/stk_reproduction/test_stk.py:97-135 (what
    stk's ``MCHammer`` optimizer does): atoms + sorted bonds -> :class:`Molecule`,
    ``bond_pair_ids = get_long_bond_ids(state)``, ``subunits = get_subunits(state)``, then
    ``Optimizer(..., random_seed=seed).get_result(...)``.  States that share a topology (same
    atom count, long bonds and building-block assignment -- e.g. conformers or a screening library
    built from one topology graph) go to the device as ONE ``get_results_batch`` call with
    per-chain start geometries; different topologies are grouped and run one group after another.

    optimizer: a :class:`mchammer_b200.Optimizer` (its potential parameters, num_steps and
        precision are used; its own generator is not touched).
    seeds: one integer for every state (stk constructs a fresh ``Optimizer(random_seed=1000)``
        per molecule) or one integer per state.

    Returns a list, in input order, of ``(new_state, record)``: ``new_state`` is
    ``state.with_position_matrix(final)`` when the state has that method (stk molecules do), else
    a :class:`Molecule`; ``record`` is a dict with the last step's ``system_potential``,
    ``nonbonded_potential``, ``max_bond_distance``, ``passed`` and ``n_accepted``.
    """
    import numpy as np

    states = list(states)
    if isinstance(seeds, (int, np.integer)):
        seeds = [int(seeds)] * len(states)
    seeds = [int(s) for s in seeds]
    if len(seeds) != len(states):
        raise ValueError("seeds must be one integer or one integer per state")
    groups: dict = {}
    lowered = []
    for index, state in enumerate(states):
        long_bonds = get_long_bond_ids(state)
        subunits = {k: list(v) for k, v in get_subunits(state).items()}
        positions = np.asarray(state.get_position_matrix(), dtype=np.float64)
        key = (positions.shape[0], long_bonds, tuple((k, tuple(v)) for k, v in subunits.items()))
        groups.setdefault(key, []).append(index)
        lowered.append((long_bonds, subunits, positions))
    out: list = [None] * len(states)
    for members in groups.values():
        first = members[0]
        long_bonds, subunits, _ = lowered[first]
        template = molecule_from_stk(states[first])
        stack = np.stack([lowered[i][2] for i in members])
        res = optimizer.get_results_batch(template, long_bonds, subunits, seeds=[seeds[i] for i in members],
                                          positions=stack, devices=devices)
        for slot, index in enumerate(members):
            final = np.array(res.positions[slot], dtype=np.float64)
            state = states[index]
            new_state = (state.with_position_matrix(final) if hasattr(state, "with_position_matrix")
                         else molecule_from_stk(state).with_position_matrix(final))
            out[index] = (new_state, {
                "system_potential": float(res.system_potential[slot]),
                "nonbonded_potential": float(res.nonbonded_potential[slot]),
                "max_bond_distance": float(res.max_bond_distance[slot]),
                "passed": None if res.last_passed[slot] < 0 else bool(res.last_passed[slot]),
                "n_accepted": int(res.n_accepted[slot]),
            })
    return out


def collapse_batch(states, collapser):
    """Collapse many stk-shaped molecules/states in batched launches (NEW; no reference counterpart).

    Per state this is what stk's ``Collapser`` optimizer does with the reference (the wiring of
    tests/stk_reproduction/test_stk.py:97-135 with ``mch.Collapser`` in place of ``mch.Optimizer``):
    atoms + sorted bonds -> :class:`Molecule`, ``bond_pair_ids = get_long_bond_ids(state)``,
    ``subunits = get_subunits(state)``, then ``Collapser(...).get_result(...)``.  States that share a
    topology (atom count, elements, long bonds, building-block assignment) go to the device as ONE
    ``Collapser.get_results_batch`` call with per-molecule start geometries.

    Returns a list, in input order, of ``(new_state, record)``; ``record`` holds ``steps_run``,
    ``min_distance`` (the value the reference tests against ``distance_threshold / 2``) and
    ``max_bond_distance`` after the last step.
    """
    import numpy as np

    states = list(states)
    groups: dict = {}
    lowered = []
    for index, state in enumerate(states):
        long_bonds = get_long_bond_ids(state)
        subunits = {k: list(v) for k, v in get_subunits(state).items()}
        positions = np.asarray(state.get_position_matrix(), dtype=np.float64)
        elements = tuple(type(a).__name__ for a in state.get_atoms())
        key = (positions.shape[0], elements, long_bonds, tuple((k, tuple(v)) for k, v in subunits.items()))
        groups.setdefault(key, []).append(index)
        lowered.append((long_bonds, subunits, positions))
    out: list = [None] * len(states)
    for members in groups.values():
        first = members[0]
        long_bonds, subunits, _ = lowered[first]
        template = molecule_from_stk(states[first])
        stack = np.stack([lowered[i][2] for i in members])
        res = collapser.get_results_batch(template, long_bonds, subunits, stack)
        for slot, index in enumerate(members):
            final = np.array(res.positions[slot], dtype=np.float64)
            state = states[index]
            new_state = (state.with_position_matrix(final) if hasattr(state, "with_position_matrix")
                         else molecule_from_stk(state).with_position_matrix(final))
            out[index] = (new_state, {
                "steps_run": int(res.steps_run[slot]),
                "min_distance": float(res.min_distance[slot]),
                "max_bond_distance": float(res.max_bond_distance[slot]),
            })
    return out
