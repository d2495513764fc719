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
