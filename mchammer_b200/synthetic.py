"""Deterministic synthetic cages for benchmarks and scale tests (no chemistry implied).

``cube_cage``        BASELINE config 4: 8 tritopic nodes on the cube vertices + 12 ditopic
                     linkers on the edge mid-points = 20 subunits x 50 atoms = 1000 atoms,
                     24 long bonds (each linker end <-> node).
``cuboctahedron_cage`` BASELINE config 5 (M12L24 style): 12 one-atom nodes + 24 linkers x 208
                     atoms = 5004 atoms, 36 subunits, 48 long bonds.
Subunit interiors are self-avoiding random walks with 1.4-1.5 A steps (bonded along the
walk, so ``Molecule.get_subunits`` recovers the subunits); long bonds start 4-7 A long,
like stk's freshly constructed cages (poc.mol: 5.4-6.6 A).  SURVEY.md section 8-d.
"""

from __future__ import annotations

import numpy as np

from .structure import Atom, Bond, Molecule


def _blob(rng: np.random.Generator, n_atoms: int, h_fraction: float):
    """Self-avoiding walk of ``n_atoms`` sites, recentred; returns (coords, elements)."""
    pts = np.zeros((n_atoms, 3))
    for k in range(1, n_atoms):
        for _ in range(200):
            step = rng.normal(size=3)
            step *= rng.uniform(1.4, 1.5) / np.linalg.norm(step)
            # bias toward the origin so the blob stays compact
            cand = pts[k - 1] + step - 0.08 * pts[k - 1]
            if np.min(np.linalg.norm(pts[:k] - cand, axis=1)) > 1.25:
                break
        pts[k] = cand
    pts -= pts.mean(axis=0)
    elements = ["H" if rng.random() < h_fraction else "C" for _ in range(n_atoms)]
    elements[0] = "C"
    return pts, elements


def _assemble(centres, sizes, links, seed: int, h_fraction: float, spread: float):
    """Place one blob per centre; ``links`` = (subunit_a, subunit_b) pairs -> long bonds
    between the atoms of a and b that are closest to each other's centre."""
    rng = np.random.default_rng(seed)
    coords, elements, bonds, subunits = [], [], [], {}
    offset = 0
    blobs = []
    for s, (centre, size) in enumerate(zip(centres, sizes)):
        pts, els = _blob(rng, size, h_fraction) if size > 1 else (np.zeros((1, 3)), ["Pd"])
        pts = pts + np.asarray(centre) * spread
        blobs.append((offset, pts))
        coords.append(pts)
        elements += els
        bonds += [(offset + k, offset + k + 1) for k in range(size - 1)]
        subunits[s] = list(range(offset, offset + size))
        offset += size
    long_bonds = []
    used: set[int] = set()
    for a, b in links:
        (oa, pa), (ob, pb) = blobs[a], blobs[b]
        ca, cb = pa.mean(axis=0), pb.mean(axis=0)
        order_a = np.argsort(np.linalg.norm(pa - cb, axis=1))
        order_b = np.argsort(np.linalg.norm(pb - ca, axis=1))
        ia = next(int(i) + oa for i in order_a if (int(i) + oa) not in used or len(pa) == 1)
        ib = next(int(i) + ob for i in order_b if (int(i) + ob) not in used or len(pb) == 1)
        used.update((ia, ib))
        long_bonds.append(tuple(sorted((ia, ib))))
    coords = np.concatenate(coords)
    all_bonds = bonds + long_bonds
    mol = Molecule(
        atoms=tuple(Atom(id=i, element_string=e) for i, e in enumerate(elements)),
        bonds=tuple(Bond(id=i, atom_ids=b) for i, b in enumerate(all_bonds)),
        position_matrix=coords,
    )
    return mol, tuple(long_bonds), subunits


def cube_cage(atoms_per_subunit: int = 50, seed: int = 20240, h_fraction: float = 0.0,
              spread: float = 12.3):
    """8 vertex nodes + 12 edge linkers.  Returns ``(Molecule, long_bond_ids, subunits)``."""
    verts = [np.array([x, y, z], float) for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)]
    centres = list(verts)
    links = []
    for i in range(8):
        for j in range(i + 1, 8):
            if np.sum(np.abs(verts[i] - verts[j])) == 2:  # cube edge
                centres.append((verts[i] + verts[j]) / 2)
                edge = len(centres) - 1
                links += [(edge, i), (edge, j)]
    sizes = [atoms_per_subunit] * len(centres)
    return _assemble(centres, sizes, links, seed, h_fraction, spread)


def cuboctahedron_cage(atoms_per_linker: int = 208, seed: int = 20240, h_fraction: float = 0.0,
                       spread: float = 16.0):
    """12 one-atom nodes on a cuboctahedron + 24 edge linkers (M12L24 style)."""
    verts = []
    for a in (-1, 1):
        for b in (-1, 1):
            verts += [np.array([a, b, 0.0]), np.array([a, 0.0, b]), np.array([0.0, a, b])]
    centres = list(verts)
    links = []
    for i in range(12):
        for j in range(i + 1, 12):
            if abs(np.linalg.norm(verts[i] - verts[j]) - np.sqrt(2.0)) < 1e-9:
                centres.append((verts[i] + verts[j]) / 2)
                edge = len(centres) - 1
                links += [(edge, i), (edge, j)]
    sizes = [1] * 12 + [atoms_per_linker] * 24
    return _assemble(centres, sizes, links, seed, h_fraction, spread)
