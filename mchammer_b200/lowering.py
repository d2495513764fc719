"""Lowering: ``Molecule`` + ``bond_pair_ids`` + ``subunits``  ->  flat arrays for the kernels.

The kernels want every subunit to be a contiguous *slot* range, so atoms are permuted
(``perm[slot] = atom id``); positions stay in the caller's atom order in HBM and are
gathered through ``perm`` once per chain.  Semantics preserved from the reference:

* subunit order == dict order of ``subunits``; a bond end belongs to the FIRST subunit that
  contains it (reference optimizer.py:220-221); an end in no subunit raises StopIteration,
  which is what the reference's bare ``next(...)`` does;
* the order inside a bond pair is the caller's (sign of the bond vector, utilities.py:24-31);
* atoms that are in no subunit never move but still count in potentials and centroids:
  they become one trailing static group;
* overlapping subunits (``allow_overlap=True``, the Optimizer): the slot ranges are the
  FIRST-OWNER partition (it decides ``bond_su``), and every subunit additionally gets its explicit
  *row list* -- the slots of every atom it lists, each once, with the multiplicity of the atom in
  the caller's list (the reference's subunit centroid is the mean over the LISTED ids,
  molecule.py:204-207, while a move translates each listed atom once, mc_operations.py:15-28);
* Collapser: hydrogens are skipped by the contact test (collapser.py:77) -> within each
  subunit the non-H atoms are put first and counted in ``n_heavy``.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, Mapping, Sequence

import numpy as np


@dataclass
class LoweredTopology:
    n_atoms: int
    n_subunits: int          # including the trailing static group, if any
    n_user_subunits: int     # entries of the caller's dict
    perm: np.ndarray         # int32 [N]   slot -> atom id
    su_offsets: np.ndarray   # int32 [S+1]
    bonds: np.ndarray        # int32 [B,2] atom ids, caller's order
    bond_su: np.ndarray      # int32 [B,2] subunit index of each end (-1 where unknown)
    n_heavy: np.ndarray      # int32 [S]   non-H atoms per subunit (== size if no elements)
    max_subunit_atoms: int
    subunit_keys: list       # caller's keys, dict order
    subunit_sizes: list[int]
    # explicit moving sets (one entry per USER subunit), always filled:
    general: bool = False            # True: some atom is listed by two subunits or twice by one
    row_offsets: np.ndarray = None   # int32 [S_user+1] into row_slots / row_weights
    row_slots: np.ndarray = None     # int32 [R] slots of the atoms subunit s lists, ascending, each once
    row_weights: np.ndarray = None   # int32 [R] how often the caller's list names that atom
    run_offsets: np.ndarray = None   # int32 [S_user+1] into runs
    runs: np.ndarray = None          # int32 [K,2] maximal contiguous slot ranges [begin, end) of each row list
    max_rows: int = 0                # longest row list

    @property
    def n_bonds(self) -> int:
        return int(self.bonds.shape[0])


def _as_pairs(bond_pair_ids: Iterable[Sequence[int]]) -> np.ndarray:
    pairs = [(int(b[0]), int(b[1])) for b in bond_pair_ids]
    return np.array(pairs, dtype=np.int32).reshape(-1, 2)


def lower_topology(
    n_atoms: int,
    bond_pair_ids: Iterable[Sequence[int]],
    subunits: Mapping,
    elements: Sequence[str] | None = None,
    require_bond_subunits: bool = True,
    allow_overlap: bool = False,
) -> LoweredTopology:
    """Build the slot permutation and the bond/subunit tables (host, O(N + B*S)).

    allow_overlap: accept subunits that share atoms (or list an atom twice) with the
    reference's semantics (optimizer.py:220-226, :248-252) instead of raising ValueError.
    """
    bonds = _as_pairs(bond_pair_ids)
    if bonds.size and (bonds.min() < 0 or bonds.max() >= n_atoms):
        raise ValueError("bond_pair_ids refers to an atom id outside the molecule")

    keys = list(subunits)
    members: list[np.ndarray] = []
    for key in keys:
        try:
            ids = np.fromiter(subunits[key], dtype=np.int64)
        except (TypeError, ValueError):
            ids = np.array([int(a) for a in subunits[key]], dtype=np.int64)
        if ids.size == 0:
            raise ValueError(f"subunit {key!r} has no atoms")
        members.append(ids)
    owner = np.full(n_atoms, -1, dtype=np.int64)
    flat = np.concatenate(members) if members else np.zeros(0, dtype=np.int64)
    clean = flat.size == 0 or (flat.min() >= 0 and flat.max() < n_atoms
                                and np.bincount(flat, minlength=n_atoms).max() <= 1)
    listed = [m.copy() for m in members]  # the caller's lists (with repeats), kept for the row lists
    general = False
    if clean:  # the usual case: a partition (of a subset) of the atoms
        owner[flat] = np.repeat(np.arange(len(members)), [m.size for m in members])
    elif allow_overlap and flat.min() >= 0 and flat.max() < n_atoms:
        # first-owner partition: an atom belongs to the first dict entry that lists it
        general = True
        owned = []
        for index, ids in enumerate(members):
            first = ids[np.sort(np.unique(ids, return_index=True)[1])]  # listed order, each once
            mine = first[owner[first] < 0]
            owner[mine] = index
            owned.append(mine)
        members = owned  # may contain empty ranges (a subunit whose atoms all belong to earlier ones)
    else:
        # something is wrong: walk the subunits in the caller's order so that the first offending
        # atom decides the error, as a per-atom loop would
        for index, (key, ids) in enumerate(zip(keys, members)):
            outside = (ids < 0) | (ids >= n_atoms)
            taken = np.where(outside, -1, owner[np.where(outside, 0, ids)])
            repeated = np.ones(ids.size, dtype=bool)
            repeated[np.unique(ids, return_index=True)[1]] = False
            bad = outside | repeated | (taken != -1)
            if bad.any():
                k = int(np.argmax(bad))
                atom = int(ids[k])
                if outside[k]:
                    raise ValueError(f"subunit {key!r} refers to atom {atom} outside the molecule")
                if repeated[k]:
                    raise ValueError(f"atom {atom} is listed twice in subunit {key!r}")
                raise ValueError(
                    f"atom {atom} is in two subunits ({keys[int(taken[k])]!r} and {key!r}); "
                    "overlapping subunits cannot be moved rigidly on the device"
                )
            owner[ids] = index
    loose = np.nonzero(owner < 0)[0]
    if loose.size:
        members.append(loose.astype(np.int64))  # static group: never selected, still interacts

    is_h = None
    if elements is not None:
        is_h = np.array([e == "H" for e in elements], dtype=bool)
        members = [np.concatenate([ids[~is_h[ids]], ids[is_h[ids]]]) for ids in members]

    perm = np.concatenate(members).astype(np.int32) if members else np.zeros(0, dtype=np.int32)
    sizes = [int(ids.size) for ids in members]
    # explicit row lists of the user subunits (slots ascending, multiplicities, contiguous runs)
    slot_of = np.empty(n_atoms, dtype=np.int64)
    slot_of[perm] = np.arange(perm.size)
    row_offsets, row_slots, row_weights, run_offsets, runs = [0], [], [], [0], []
    for ids in listed:
        slots, counts = np.unique(slot_of[ids], return_counts=True)
        row_slots.append(slots)
        row_weights.append(counts)
        row_offsets.append(row_offsets[-1] + slots.size)
        breaks = np.flatnonzero(np.diff(slots) != 1) + 1
        begins = np.concatenate([[0], breaks])
        ends = np.concatenate([breaks, [slots.size]])
        runs.append(np.stack([slots[begins], slots[ends - 1] + 1], axis=1))
        run_offsets.append(run_offsets[-1] + begins.size)
    cat = lambda parts, shape: (np.concatenate(parts) if parts else np.zeros(shape, dtype=np.int64)).astype(np.int32)  # noqa: E731
    su_offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    if is_h is None:
        n_heavy = np.array(sizes, dtype=np.int32)
    else:
        n_heavy = np.array([int((~is_h[ids]).sum()) for ids in members], dtype=np.int32)

    bond_su = np.full((bonds.shape[0], 2), -1, dtype=np.int32)
    n_user = len(keys)
    for b, (a1, a2) in enumerate(bonds):
        for side, atom in enumerate((a1, a2)):
            index = int(owner[atom])
            if index < 0 or index >= n_user:
                if require_bond_subunits:
                    # reference: next(i for i in subunits if atom in subunits[i]) -> StopIteration
                    raise StopIteration(f"atom {int(atom)} of bond {(int(a1), int(a2))} is in no subunit")
                continue
            bond_su[b, side] = index

    return LoweredTopology(
        n_atoms=n_atoms,
        n_subunits=len(members),
        n_user_subunits=n_user,
        perm=perm,
        su_offsets=su_offsets,
        bonds=bonds,
        bond_su=bond_su,
        n_heavy=n_heavy,
        max_subunit_atoms=max(sizes) if sizes else 0,
        subunit_keys=keys,
        subunit_sizes=[len(subunits[k]) for k in keys],
        general=general,
        row_offsets=np.array(row_offsets, dtype=np.int32),
        row_slots=cat(row_slots, (0,)),
        row_weights=cat(row_weights, (0,)),
        run_offsets=np.array(run_offsets, dtype=np.int32),
        runs=cat(runs, (0, 2)).reshape(-1, 2),
        max_rows=max((int(r.size) for r in row_slots), default=0),
    )
