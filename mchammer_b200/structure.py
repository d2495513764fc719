"""Host-side data model: ``Atom``, ``Bond``, ``Molecule``.

These mirror the reference's public surface (reference: src/mchammer/_internal/atom.py:10-65,
bond.py:9-54, molecule.py:18-249) so that callers of ``mchammer`` can switch packages
unchanged.  They are plain host containers; nothing here is on the Monte-Carlo hot path.
The hot path consumes the flat arrays produced by :mod:`mchammer_b200.lowering`.
"""

from __future__ import annotations

from collections import deque
from dataclasses import dataclass
from typing import Iterable, Iterator

import numpy as np

from .elements import get_radius


@dataclass
class Atom:
    """An atom: integer id (== row of the position matrix) and element symbol.

    ``radius`` defaults to the STREUSEL table value (reference atom.py:41-45).
    ``sigma``/``epsilon``/``charge`` are carried for callers; no potential reads them.
    """

    id: int
    element_string: str
    radius: float | None = None
    sigma: float | None = None
    epsilon: float | None = None
    charge: float | None = None

    def __post_init__(self) -> None:
        if self.radius is None:
            self.radius = get_radius(self.element_string)

    def get_id(self) -> int:
        return self.id

    def get_element_string(self) -> str:
        return self.element_string

    def get_radius(self) -> float | None:
        return self.radius

    def __repr__(self) -> str:
        return f"{self.element_string}(id={self.id})"

    __str__ = __repr__


@dataclass
class Bond:
    """A bond between two distinct atoms; ids are stored sorted (reference bond.py:25-30)."""

    id: int
    atom_ids: tuple[int, int]

    def __post_init__(self) -> None:
        if len(set(self.atom_ids)) != 2:
            raise ValueError("Two distinct atom ids are required.")
        lo, hi = sorted(self.atom_ids)
        self.atom1_id = lo
        self.atom2_id = hi

    def get_id(self) -> int:
        return self.id

    def get_atom1_id(self) -> int:
        return self.atom1_id

    def get_atom2_id(self) -> int:
        return self.atom2_id

    def __repr__(self) -> str:
        return (
            f"{type(self).__name__}(id={self.id}, "
            f"atom1_id={self.atom1_id}, atom2_id={self.atom2_id})"
        )

    __str__ = __repr__


@dataclass
class Molecule:
    """Atoms + bonds + an ``(n, 3)`` position matrix (stored transposed, float64).

    Value semantics follow the reference (molecule.py:39-86): inputs are never mutated,
    ``with_*`` methods return clones, ``get_position_matrix`` returns a fresh copy.
    """

    atoms: tuple[Atom, ...]
    bonds: tuple[Bond, ...]
    position_matrix: np.ndarray

    def __post_init__(self) -> None:
        self.atoms = tuple(self.atoms)
        self.bonds = tuple(self.bonds)
        # (3, n) float64 storage, as reference molecule.py:43-46.
        self.position_matrix = np.array(
            np.asarray(self.position_matrix).T, dtype=np.float64
        )

    # -- geometry ---------------------------------------------------------------------
    def get_position_matrix(self) -> np.ndarray:
        """Fresh ``(n, 3)`` float64 copy of the coordinates."""
        return np.array(self.position_matrix.T)

    def with_position_matrix(self, position_matrix: np.ndarray) -> "Molecule":
        return Molecule(self.atoms, self.bonds, np.array(position_matrix))

    def with_displacement(self, displacement: np.ndarray) -> "Molecule":
        return Molecule(
            self.atoms, self.bonds, np.array(self.position_matrix.T + displacement)
        )

    def get_centroid(self, atom_ids: Iterable[int] | None = None) -> np.ndarray:
        """Mean position of ``atom_ids`` (all atoms if ``None``); reference molecule.py:178-207."""
        if atom_ids is None:
            atom_ids = range(len(self.atoms))
        elif not isinstance(atom_ids, (list, tuple)):
            atom_ids = tuple(atom_ids)
        count = len(atom_ids)
        if count == 0:
            raise ValueError("atom_ids was of length 0.")
        return np.divide(self.position_matrix[:, atom_ids].sum(axis=1), count)

    # -- topology ---------------------------------------------------------------------
    def get_atoms(self) -> Iterator[Atom]:
        yield from self.atoms

    def get_bonds(self) -> Iterator[Bond]:
        yield from self.bonds

    def get_num_atoms(self) -> int:
        return len(self.atoms)

    def get_subunits(self, bond_pair_ids: Iterable) -> dict[int, set[int]]:
        """Connected components after deleting the bonds listed in ``bond_pair_ids``.

        Same contract as reference molecule.py:209-238: a bond is deleted only if its
        *sorted tuple* ``(atom1_id, atom2_id)`` compares ``in bond_pair_ids``; components
        are numbered in order of their first atom (input atom order) and returned as sets.
        Implemented as a breadth-first sweep over an adjacency list (no graph library).
        """
        neighbours: dict[int, list[int]] = {a.get_id(): [] for a in self.atoms}
        for bond in self.bonds:
            key = (bond.get_atom1_id(), bond.get_atom2_id())
            if key in bond_pair_ids:
                continue
            neighbours[key[0]].append(key[1])
            neighbours[key[1]].append(key[0])

        subunits: dict[int, set[int]] = {}
        visited: set[int] = set()
        for start in neighbours:
            if start in visited:
                continue
            component = {start}
            frontier = deque([start])
            while frontier:
                here = frontier.popleft()
                for there in neighbours[here]:
                    if there not in component:
                        component.add(there)
                        frontier.append(there)
            visited |= component
            subunits[len(subunits)] = component
        return subunits

    # -- writers (reference molecule.py:88-164) ----------------------------------------
    def write_xyz_content(self) -> list[str]:
        coords = self.get_position_matrix()
        lines = [f"{len(self.atoms)}\n\n"]
        for atom in self.atoms:
            x, y, z = coords[atom.get_id()]
            lines.append(f"{atom.get_element_string()} {x:f} {y:f} {z:f}\n")
        return lines

    def write_xyz_file(self, path: str) -> None:
        with open(path, "w") as handle:
            handle.write("".join(self.write_xyz_content()))

    def _write_pdb_content(self) -> list[str]:
        coords = self.get_position_matrix()
        seen_per_element: dict[str, int] = {}
        present: set[int] = set()
        lines: list[str] = []
        for atom in self.atoms:
            atom_id = atom.get_id()
            present.add(atom_id)
            element = atom.get_element_string()
            seen_per_element[element] = seen_per_element.get(element, 0) + 1
            label = f"{element}{seen_per_element[element]}"
            x, y, z = coords[atom_id]
            lines.append(
                f"{'HETATM':<6}{atom_id + 1:>5} {label:<4}{'':<1}{'UNL':<3} {'':<1}"
                f"{'1':>4}{'':<1}    {x:>7.3f} {y:>7.3f} {z:>7.3f}"
                f"{'1.00':>6}{'0.00':>6}          {element:>2}{0:>2}\n"
            )
        for bond in self.bonds:
            a1, a2 = bond.get_atom1_id(), bond.get_atom2_id()
            if a1 in present and a2 in present:
                lines.append(f"{'CONECT':<6}{a1 + 1:>5}{a2 + 1:>5}               \n")
        lines.append("END\n")
        return lines

    def write_pdb_file(self, path: str) -> None:
        with open(path, "w") as handle:
            handle.write("".join(self._write_pdb_content()))

    def __repr__(self) -> str:
        return f"<{type(self).__name__}({len(self.atoms)} atoms) at {id(self)}>"

    __str__ = __repr__
