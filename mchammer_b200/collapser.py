"""``Collapser``: rigid subunits slide toward the molecular centroid until they touch.

Same constructor and methods as reference src/mchammer/_internal/collapser.py:20-340.
``get_result`` / ``get_trajectory`` run the whole data-dependent loop inside
``mch_collapse_batch`` (CUDA).  The three private helpers the reference's tests call
(``_get_subunit_distances``, ``_get_subunit_vectors``, ``_get_new_position_matrix``) are
kept as small host conveniences; the drivers do not use them.
"""

from __future__ import annotations

import time
from dataclasses import dataclass
from itertools import combinations
from typing import Iterator

import numpy as np

from . import engine, native
from .lowering import lower_topology
from .primitives import get_atom_distance
from .records import Result, StepResult
from .structure import Molecule


@dataclass
class CollapseBatchResult:
    positions: np.ndarray          # [M,N,3]
    steps_run: np.ndarray          # [M]
    min_distance: np.ndarray       # [M] min cross-subunit non-H distance tested against threshold/2
    max_bond_distance: np.ndarray  # [M] after the last step
    seconds: float


class Collapser:
    """Moves rigid subunits toward the centroid of the molecule (on CUDA)."""

    def __init__(
        self,
        step_size: float,
        distance_threshold: float,
        scale_steps: bool = True,  # noqa: FBT001, FBT002
        *,
        precision: str = "fp64",
        device: int | None = None,
        max_steps: int = 1_000_000,
    ) -> None:
        """Reference arguments (collapser.py:23-47) plus keyword-only device options.

        max_steps bounds the loop: the reference never terminates when no pair of non-H
        atoms in distinct subunits can come within the threshold (e.g. one subunit);
        here that raises ``RuntimeError`` after ``max_steps`` steps.
        """
        self._step_size = step_size
        self._distance_threshold = distance_threshold
        self._scale_steps = scale_steps
        if precision not in engine.PRECISIONS:
            raise ValueError(f"precision must be 'fp64' or 'fp32', got {precision!r}")
        self._precision = precision
        self._device = device
        self._max_steps = int(max_steps)
        self._trace_capacity_guess = 64  # steps get_trajectory records in its first launch

    # ------------------------------------------------------------------ host conveniences
    def _get_subunit_distances(self, mol: Molecule, subunits: dict) -> Iterator[float]:
        """Cross-subunit distances, H skipped, in the reference's order (collapser.py:56-82)."""
        elements = [a.get_element_string() for a in mol.get_atoms()]
        coords = mol.get_position_matrix()
        for first, second in combinations(subunits, 2):
            for a1 in subunits[first]:
                if elements[a1] == "H":
                    continue
                for a2 in subunits[second]:
                    if elements[a2] != "H":
                        yield get_atom_distance(coords, a1, a2)

    def _has_short_contacts(self, mol: Molecule, subunits: dict) -> bool:
        return any(d < self._distance_threshold for d in self._get_subunit_distances(mol, subunits))

    def _get_subunit_vectors(self, mol: Molecule, subunits: dict) -> tuple[dict, dict]:
        """Centroid vectors and relative magnitudes per subunit (collapser.py:106-157)."""
        centre = mol.get_centroid()
        vectors = {k: mol.get_centroid(atom_ids=subunits[k]) - centre for k in subunits}
        if self._scale_steps:
            lengths = {k: np.linalg.norm(v) for k, v in vectors.items()}
            longest = max(lengths.values())
            scales = {k: float(lengths[k] / longest) for k in lengths}
        else:
            scales = {k: 1.0 for k in vectors}
        return vectors, scales

    def _get_new_position_matrix(self, mol, subunits, vectors, scales, step_size) -> np.ndarray:
        """pos[a] - step_size * v_s * scale_s for every atom of every subunit (collapser.py:84-104)."""
        if step_size is None:
            step_size = self._step_size
        coords = mol.get_position_matrix()
        moved = coords.copy()
        for k in subunits:
            shift = step_size * vectors[k] * scales[k]
            for atom in subunits[k]:
                moved[atom] = coords[atom] - shift
        return moved

    # ------------------------------------------------------------------ device runs
    def _torch_device(self):
        torch = native.require_cuda()
        return torch.device("cuda", torch.cuda.current_device() if self._device is None else self._device)

    def _run(self, mol: Molecule, positions, bond_pair_ids, subunits, n_mol, trace_capacity=0,
             want_trajectory=False):
        torch = native.require_cuda()
        dev = self._torch_device()
        elements = [a.get_element_string() for a in mol.get_atoms()]
        dtopo = engine.cached_topology(mol.get_num_atoms(), bond_pair_ids, subunits, dev.index, elements=elements,
                                       require_bond_subunits=False)
        topo = dtopo.host
        pos = torch.from_numpy(np.ascontiguousarray(positions)).to(dev)
        out = engine.collapse_on_device(
            dtopo, pos, n_mol, self._step_size, self._distance_threshold, self._scale_steps,
            self._precision, self._max_steps, trace_capacity, want_trajectory,
        )
        if int(out.status.max().item()) != 0:
            raise RuntimeError(
                f"Collapser: no contact below {self._distance_threshold} after {self._max_steps} steps"
            )
        return out

    def get_result(self, mol: Molecule, bond_pair_ids, subunits: dict) -> tuple[Molecule, StepResult]:
        """Collapsed molecule and the record of the last step (reference collapser.py:287-340)."""
        out = self._run(mol, mol.get_position_matrix(), bond_pair_ids, subunits, 1)
        steps = int(out.steps_run.item())
        if steps == 0:
            # the reference returns an unbound `step_result` here (UnboundLocalError)
            raise UnboundLocalError("Collapser.get_result: no step was needed, there is no step result")
        final = out.positions[0].cpu().numpy().astype(np.float64)
        far = float(out.max_bond_distance.item())
        record = StepResult(step=steps, position_matrix=final, max_bond_distance=far,
                            log=f"{steps} {far}\n")
        return mol.with_position_matrix(final), record

    def get_trajectory(self, mol: Molecule, bond_pair_ids, subunits: dict) -> tuple[Molecule, Result]:
        """Collapsed molecule and every step's record + log (reference collapser.py:200-285)."""
        result = Result(start_time=time.time())
        rule = "=" * 52 + "\n"
        result.update_log(
            rule + "                Collapser algorithm.\n" + rule
            + f"Step size: {self._step_size}\n"
            f"Scale steps?: {self._scale_steps}\n"
            f"Distance threshold: {self._distance_threshold}\n"
        )
        result.update_log(f"There are {len(bond_pair_ids)} bonds to optimize.\n")
        result.update_log(
            f"There are {len(subunits)} sub units with N atoms:\n"
            f"{[len(subunits[i]) for i in subunits]}\n"
        )
        result.update_log(rule + " " * 17 + "Running optimisation!" + " " * 14 + "\n" + rule + "\n")

        # the loop length is data dependent: the trace buffers are sized by a guess and the launch is
        # repeated only if the loop ran longer (the kernel records the first `trace_capacity` steps
        # and keeps counting)
        coords = mol.get_position_matrix()
        capacity = self._trace_capacity_guess
        out = self._run(mol, coords, bond_pair_ids, subunits, 1, trace_capacity=capacity, want_trajectory=True)
        steps = int(out.steps_run.item())
        if steps > capacity:
            out = self._run(mol, coords, bond_pair_ids, subunits, 1, trace_capacity=steps, want_trajectory=True)
        if steps > 0:
            far = out.trace_max_bond[0, :steps].cpu().numpy()
            frames = out.trajectory[0, :steps].cpu().numpy().astype(np.float64)
            for k in range(steps):
                result.add_step_result(
                    StepResult(step=k + 1, position_matrix=frames[k],
                               max_bond_distance=float(far[k]), log=f"{k + 1} {float(far[k])}\n")
                )
        min_dist = float(out.min_distance.item())
        result.update_log(
            rule + f"Steps run: {steps}\n"
            f"Minimum inter-subunit distance: {min_dist}\n" + rule
        )
        final = out.positions[0].cpu().numpy().astype(np.float64)
        return mol.with_position_matrix(final), result

    def get_results_batch(self, mol: Molecule, bond_pair_ids, subunits: dict,
                          positions: np.ndarray) -> CollapseBatchResult:
        """Collapse ``positions[m]`` ([M,N,3], same topology as ``mol``) independently (NEW)."""
        start = time.perf_counter()
        positions = np.ascontiguousarray(positions)
        if positions.ndim != 3 or positions.shape[1:] != (mol.get_num_atoms(), 3):
            raise ValueError("positions must have shape [M, n_atoms, 3]")
        out = self._run(mol, positions, bond_pair_ids, subunits, positions.shape[0])
        return CollapseBatchResult(
            positions=out.positions.cpu().numpy(),
            steps_run=out.steps_run.cpu().numpy().astype(np.int64),
            min_distance=out.min_distance.cpu().numpy(),
            max_bond_distance=out.max_bond_distance.cpu().numpy(),
            seconds=time.perf_counter() - start,
        )
