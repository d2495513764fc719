"""Result containers: ``StepResult``, ``MCStepResult``, ``Result``.

Surface parity with reference src/mchammer/_internal/results.py:14-152.  On this
framework the per-step records are *raised* from device trace buffers
(see :mod:`mchammer_b200.optimizer`), not produced one Python step at a time.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Iterator

import numpy as np


@dataclass
class StepResult:
    """One step of a deterministic (Collapser) run."""

    step: int
    position_matrix: np.ndarray
    max_bond_distance: float
    log: str

    def get_properties(self) -> dict:
        return {"max_bond_distance": self.max_bond_distance}


@dataclass
class MCStepResult(StepResult):
    """One Metropolis step: adds potentials and the accept flag (``None`` for step 0)."""

    step: int
    log: str
    position_matrix: np.ndarray
    system_potential: float
    nonbonded_potential: float
    max_bond_distance: float
    passed: bool | None = None

    def get_properties(self) -> dict:
        return {
            "max_bond_distance": self.max_bond_distance,
            "system_potential": self.system_potential,
            "nonbonded_potential": self.nonbonded_potential,
            "passed": self.passed,
        }


class Result:
    """Ordered collection of step results plus the text log (reference results.py:92-152)."""

    def __init__(self, start_time: float) -> None:
        self._start_time = start_time
        self._log = ""
        self._step_results: dict[int, StepResult] = {}
        self._step_count = 0

    def add_step_result(self, step_result: StepResult) -> None:
        self._step_results[step_result.step] = step_result
        self._log += step_result.log
        self._step_count = step_result.step

    def update_log(self, string: str) -> None:
        self._log += string

    def get_log(self) -> str:
        return self._log

    def get_number_passed(self) -> int:
        return sum(1 for r in self._step_results.values() if r.passed)  # type: ignore[attr-defined]

    def get_final_position_matrix(self) -> np.ndarray:
        return self._step_results[self._step_count].position_matrix

    def get_timing(self, time: float) -> float:
        return time - self._start_time

    def get_step_count(self) -> int:
        return self._step_count

    def get_steps_properties(self) -> Iterator[tuple[int, dict]]:
        for step, record in self._step_results.items():
            yield step, record.get_properties()

    def get_trajectory(self) -> Iterator[tuple[int, np.ndarray]]:
        for step, record in self._step_results.items():
            yield step, record.position_matrix
