"""Minimal MDL ``.mol`` reader/writer (V2000 and V3000 CTABs) -- no stk / rdkit needed.

The reference's examples build their ``Molecule`` through stk
(reference examples/minimum_example.py:7-20, examples/stk_example.py:76-97); the files
they consume and emit are plain molfiles, so this module is the host-side data-format
adapter on the input/output edge of the hot path.

Conventions: atom id == 0-based file order, element symbol as
written, bond ids in file order with 0-based atom ids.
"""

from __future__ import annotations

import numpy as np

from .structure import Atom, Bond, Molecule


def parse_mol(text: str) -> tuple[list[str], np.ndarray, list[tuple[int, int]], list[int]]:
    """Return ``(elements, positions[n,3], bonds[(a1,a2)...], bond_orders)`` from molfile text."""
    lines = text.splitlines()
    if len(lines) < 4:
        raise ValueError("not a molfile: fewer than 4 lines")
    counts = lines[3]
    if "V3000" in counts:
        return _parse_v3000(lines)
    if "V2000" in counts or len(counts) >= 6:
        return _parse_v2000(lines)
    raise ValueError("unrecognised molfile counts line")


def _parse_v2000(lines: list[str]):
    counts = lines[3]
    n_atoms, n_bonds = int(counts[0:3]), int(counts[3:6])
    elements, coords = [], []
    for row in lines[4 : 4 + n_atoms]:
        coords.append((float(row[0:10]), float(row[10:20]), float(row[20:30])))
        elements.append(row[31:34].strip())
    bonds, orders = [], []
    for row in lines[4 + n_atoms : 4 + n_atoms + n_bonds]:
        bonds.append((int(row[0:3]) - 1, int(row[3:6]) - 1))
        orders.append(int(row[6:9]))
    return elements, np.array(coords, dtype=np.float64).reshape(-1, 3), bonds, orders


def _parse_v3000(lines: list[str]):
    elements, coords, bonds, orders = [], [], [], []
    section = None
    for raw in lines:
        if not raw.startswith("M  V30 "):
            continue
        fields = raw[7:].split()
        if not fields:
            continue
        if fields[0] == "BEGIN":
            section = fields[1]
            continue
        if fields[0] == "END":
            section = None
            continue
        if section == "ATOM":
            elements.append(fields[1])
            coords.append((float(fields[2]), float(fields[3]), float(fields[4])))
        elif section == "BOND":
            orders.append(int(fields[1]))
            bonds.append((int(fields[2]) - 1, int(fields[3]) - 1))
    return elements, np.array(coords, dtype=np.float64).reshape(-1, 3), bonds, orders


def read_mol(path: str) -> Molecule:
    """Build a :class:`Molecule` from a ``.mol`` file (ids follow file order)."""
    with open(path) as handle:
        elements, coords, bonds, _ = parse_mol(handle.read())
    return Molecule(
        atoms=tuple(Atom(id=i, element_string=e) for i, e in enumerate(elements)),
        bonds=tuple(Bond(id=i, atom_ids=b) for i, b in enumerate(bonds)),
        position_matrix=coords,
    )


def mol_v3000_text(mol: Molecule, bond_orders: list[int] | None = None) -> str:
    """Serialise ``mol`` as a V3000 molfile with 4-decimal coordinates (stk/rdkit style)."""
    coords = mol.get_position_matrix()
    atoms = list(mol.get_atoms())
    bonds = list(mol.get_bonds())
    if bond_orders is None:
        bond_orders = [1] * len(bonds)
    out = [
        "",
        "     mchammer_b200  3D",
        "",
        "  0  0  0  0  0  0  0  0  0  0999 V3000",
        "M  V30 BEGIN CTAB",
        f"M  V30 COUNTS {len(atoms)} {len(bonds)} 0 0 0",
        "M  V30 BEGIN ATOM",
    ]
    for k, atom in enumerate(atoms, 1):
        x, y, z = coords[atom.get_id()]
        out.append(f"M  V30 {k} {atom.get_element_string()} {x:.4f} {y:.4f} {z:.4f} 0")
    out += ["M  V30 END ATOM", "M  V30 BEGIN BOND"]
    for k, (bond, order) in enumerate(zip(bonds, bond_orders), 1):
        out.append(
            f"M  V30 {k} {order} {bond.get_atom1_id() + 1} {bond.get_atom2_id() + 1}"
        )
    out += ["M  V30 END BOND", "M  V30 END CTAB", "M  END", ""]
    return "\n".join(out)


def write_mol(mol: Molecule, path: str, bond_orders: list[int] | None = None) -> None:
    with open(path, "w") as handle:
        handle.write(mol_v3000_text(mol, bond_orders))
