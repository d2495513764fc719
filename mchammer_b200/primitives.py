"""Host-side primitive operations kept for API parity with the reference.

``get_atom_distance`` / ``get_bond_vector``   reference utilities.py:13-31
``translate_atoms_along_vector``             reference mc_operations.py:15-28
``translate_molecule_along_vector``          reference mc_operations.py:31-36
``test_move``                                reference mc_operations.py:39-52
rotation helpers                              reference mc_operations.py:55-114

The Monte-Carlo drivers in this package do NOT call these per step; the same
arithmetic runs inside the CUDA step kernel (csrc/mch_optimize.cuh).  They exist
so user code written against ``mchammer`` keeps working.
"""

from __future__ import annotations

from typing import Iterable

import numpy as np

from .structure import Molecule


def get_atom_distance(position_matrix: np.ndarray, atom1_id: int, atom2_id: int) -> float:
    """Euclidean distance between two rows of ``position_matrix``."""
    delta = np.asarray(position_matrix[atom1_id]) - np.asarray(position_matrix[atom2_id])
    return float(np.linalg.norm(delta, ord=2))


def get_bond_vector(position_matrix: np.ndarray, bond_pair: tuple[int, int]) -> np.ndarray:
    """Vector from atom ``bond_pair[0]`` to atom ``bond_pair[1]``."""
    return position_matrix[bond_pair[1]] - position_matrix[bond_pair[0]]


def translate_atoms_along_vector(
    mol: Molecule, atom_ids: Iterable[int], vector: np.ndarray
) -> Molecule:
    """Clone of ``mol`` with ``vector`` SUBTRACTED from the listed atoms (note the sign)."""
    coords = mol.get_position_matrix()
    chosen = set(atom_ids)
    rows = [a.get_id() for a in mol.get_atoms() if a.get_id() in chosen]
    if rows:
        coords[rows] = coords[rows] - vector
    return mol.with_position_matrix(coords)


def translate_molecule_along_vector(mol: Molecule, vector: np.ndarray) -> Molecule:
    """Clone of ``mol`` displaced by ``+vector``."""
    return mol.with_displacement(vector)


def test_move(
    beta: float, curr_pot: float, new_pot: float, generator: np.random.Generator
) -> bool:
    """Metropolis criterion.  Strictly-downhill moves pass without consuming a random draw."""
    if new_pot < curr_pot:
        return True
    boltzmann = np.exp(-beta * (new_pot - curr_pot))
    return boltzmann > generator.random()


test_move.__test__ = False  # not a pytest test, despite the name the reference gave it


def rotation_matrix_arbitrary_axis(angle: float, axis: np.ndarray) -> np.ndarray:
    """3x3 rotation by ``angle`` radians about unit vector ``axis`` (Euler-Rodrigues form)."""
    half = angle / 2.0
    a = np.cos(half)
    b, c, d = np.asarray(axis) * np.sin(half)
    aa, bb, cc, dd = a * a, b * b, c * c, d * d
    matrix = np.array(
        [
            [aa + bb - cc - dd, 2 * (b * c - a * d), 2 * (b * d + a * c)],
            [2 * (b * c + a * d), aa + cc - bb - dd, 2 * (c * d - a * b)],
            [2 * (b * d - a * c), 2 * (c * d + a * b), aa + dd - bb - cc],
        ]
    )
    # The reference re-orthonormalises through scipy's Rotation (mc_operations.py:91-93).
    from scipy.spatial.transform import Rotation

    return Rotation.from_matrix(matrix).as_matrix()


def rotate_molecule_by_angle(
    mol: Molecule, angle: float, axis: np.ndarray, origin: np.ndarray
) -> Molecule:
    """Clone of ``mol`` rotated by ``angle`` about ``axis`` through ``origin``."""
    rotation = rotation_matrix_arbitrary_axis(angle, axis)
    shifted = mol.get_position_matrix() - origin
    return mol.with_position_matrix((rotation @ shifted.T).T + origin)
