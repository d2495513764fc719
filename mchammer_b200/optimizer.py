"""``Optimizer``: the reference's Metropolis driver, running on a B200.

Same constructor and methods as reference src/mchammer/_internal/optimizer.py:19-438
(``get_result``, ``get_trajectory``, ``compute_potential`` and the private helpers its
tests call), plus :meth:`Optimizer.get_results_batch` for many independent chains.
All arithmetic of the step loop happens in ``mch_optimize_batch`` (CUDA); this class
lowers the inputs, launches, and raises the device records into the reference's
``Molecule`` / ``MCStepResult`` / ``Result`` objects and log text.
"""

from __future__ import annotations

import time
from typing import Sequence

import numpy as np

from . import engine, native, rngstate
from .batch import BatchPlan, BatchResult, BatchTrajectory, shard_bounds, unpack_info
from .lowering import lower_topology
from .records import MCStepResult, Result
from .structure import Molecule


class Optimizer:
    """Optimize target bonds with a Metropolis MC of rigid subunit translations (on CUDA)."""

    def __init__(  # noqa: PLR0913
        self,
        step_size: float,
        target_bond_length: float,
        num_steps: int,
        bond_epsilon: float = 50,
        nonbond_epsilon: float = 20,
        nonbond_sigma: float = 1.2,
        nonbond_mu: float = 3,
        beta: float = 2,
        random_seed: int | None = 1000,
        *,
        precision: str = "fp64",
        device: int | None = None,
    ) -> None:
        """Reference arguments (optimizer.py:27-38) plus two keyword-only device options.

        precision:
            ``"fp64"`` reproduces the reference's accept/reject sequence (pair terms,
            sums and positions in float64); ``"fp32"`` evaluates pair terms and stores
            positions in float32 (potentials within 1e-4 relative of fp64).
        device:
            CUDA device index (default: the current device).
        """
        self._step_size = step_size
        self._target_bond_length = target_bond_length
        self._num_steps = num_steps
        self._bond_epsilon = bond_epsilon
        self._nonbond_epsilon = nonbond_epsilon
        self._nonbond_sigma = nonbond_sigma
        self._nonbond_mu = nonbond_mu
        self._beta = beta
        if precision not in engine.PRECISIONS:
            raise ValueError(f"precision must be 'fp64' or 'fp32', got {precision!r}")
        self._precision = precision
        self._device = device
        self._generator = (
            np.random.default_rng() if random_seed is None else np.random.default_rng(random_seed)
        )

    # ------------------------------------------------------------------ scalar potentials
    def _bond_potential(self, distance: float) -> float:
        """eps_b (d - R_t)^2 -- reference optimizer.py:95-102 (scalar convenience form)."""
        return self._bond_epsilon * (distance - self._target_bond_length) ** 2

    def _nonbond_potential(self, distance: float) -> float:
        """eps_nb (sigma / d)^mu -- reference optimizer.py:104-112 (scalar convenience form)."""
        return self._nonbond_epsilon * ((self._nonbond_sigma / distance) ** self._nonbond_mu)

    def _params(self) -> native.MchParams:
        return engine.make_params(
            self._step_size, self._target_bond_length, self._bond_epsilon,
            self._nonbond_epsilon, self._nonbond_sigma, self._nonbond_mu, self._beta,
        )

    def _torch_device(self):
        torch = native.require_cuda()
        index = torch.cuda.current_device() if self._device is None else self._device
        return torch.device("cuda", index)

    # ------------------------------------------------------------------ potentials on device
    def _potential(self, position_matrix: np.ndarray, bond_pair_ids) -> tuple[float, float, float]:
        torch = native.require_cuda()
        dev = self._torch_device()
        pos = torch.from_numpy(np.ascontiguousarray(position_matrix, dtype=np.float64)).to(dev)
        bonds = np.array([(int(b[0]), int(b[1])) for b in bond_pair_ids], dtype=np.int64).reshape(-1)
        n_atoms = int(np.shape(position_matrix)[0])
        if bonds.size and (bonds.min() < -n_atoms or bonds.max() >= n_atoms):
            # the reference indexes the position matrix with these ids (utilities.py:13-21): IndexError
            raise IndexError("bond_pair_ids refers to an atom id outside the molecule")
        bonds = np.where(bonds < 0, bonds + n_atoms, bonds).astype(np.int32)  # numpy-style negative ids
        bonds_dev = torch.from_numpy(bonds).to(dev) if bonds.size else None
        u, unb, far = engine.potential_on_device(
            pos, bonds_dev, bonds.size // 2, self._params(), self._precision
        )
        return float(u.item()), float(unb.item()), float(far.item())

    def _compute_nonbonded_potential(self, position_matrix: np.ndarray) -> float:
        """Sum of eps (sigma/r)^mu over all atom pairs (reference optimizer.py:114-120)."""
        return self._potential(position_matrix, ())[1]

    def compute_potential(self, mol: Molecule, bond_pair_ids) -> tuple[float, float]:
        """``(system_potential, nonbonded_potential)`` of ``mol`` (reference optimizer.py:122-142)."""
        u, unb, _ = self._potential(mol.get_position_matrix(), bond_pair_ids)
        return u, unb

    # ------------------------------------------------------------------ log text
    def _output_top_lines(self) -> str:
        rule = "=" * 52 + "\n"

        def banner(text: str) -> str:  # 52 columns, odd padding goes to the left
            left = (52 - len(text) + 1) // 2
            return (" " * left + text).ljust(52) + "\n"

        head = "".join(
            banner(t) for t in ("MCHammer optimisation", "-" * 21, "Andrew Tarzia", "-" * 21, "")
        )
        settings = (
            f"{'Settings:':<52}\n"
            f" step size = {self._step_size} \n"
            f" target bond length = {self._target_bond_length} \n"
            f" num. steps = {self._num_steps} \n"
            f" bond epsilon = {self._bond_epsilon} \n"
            f" nonbond epsilon = {self._nonbond_epsilon} \n"
            f" nonbond sigma = {self._nonbond_sigma} \n"
            f" nonbond mu = {self._nonbond_mu} \n"
            f" beta = {self._beta} \n"
        )
        return rule + head + settings + rule + "\n"

    # ------------------------------------------------------------------ one molecule
    def _run_single(self, mol: Molecule, bond_pair_ids, subunits, want_trace: bool):
        torch = native.require_cuda()
        dev = self._torch_device()
        dtopo = engine.cached_topology(
            mol.get_num_atoms(), bond_pair_ids, subunits, dev.index,
            require_bond_subunits=self._num_steps > 1, allow_overlap=True,
        )
        topo = dtopo.host
        pos = torch.from_numpy(mol.get_position_matrix()).to(dev)
        words = rngstate.words_from_bit_generator(self._generator.bit_generator)
        rng_dev = engine.rng_words_to_device(words[None, :], dev)
        out = engine.optimize_on_device(
            dtopo, pos, 1, rng_dev, self._params(), self._num_steps, self._precision,
            want_trace=want_trace, want_trajectory=want_trace,
        )
        # the generator of this Optimizer advances across calls, like the reference's
        rngstate.words_to_bit_generator(
            engine.rng_words_from_device(out.rng_state)[0], self._generator.bit_generator
        )
        return topo, out

    @staticmethod
    def _step_log(step, u, unb, far, bond_row, passed) -> str:
        if passed is None:
            return f"{step} {u} {unb} {far} -- --\n"
        return f"{step} {u} {unb} {far} {bond_row} {'T' if passed else 'F'}\n"

    def get_result(self, mol: Molecule, bond_pair_ids, subunits: dict) -> tuple[Molecule, MCStepResult]:
        """Final molecule and the record of the last step (reference optimizer.py:392-438)."""
        topo, out = self._run_single(mol, bond_pair_ids, subunits, want_trace=False)
        positions = out.positions[0].cpu().numpy().astype(np.float64)
        u = float(out.system_potential.item())
        unb = float(out.nonbonded_potential.item())
        far = float(out.max_bond_distance.item())
        bond, passed, _, _ = unpack_info(out.last_step_info.cpu().numpy())
        last_step = max(int(self._num_steps), 1) - 1
        if passed[0] < 0:
            flag, row = None, None
        else:
            flag, row = bool(passed[0]), np.asarray(bond_pair_ids)[int(bond[0])]
        record = MCStepResult(
            step=last_step, position_matrix=positions, passed=flag, system_potential=u,
            nonbonded_potential=unb, max_bond_distance=far,
            log=self._step_log(last_step, u, unb, far, row, flag),
        )
        return mol.with_position_matrix(positions), record

    def get_trajectory(self, mol: Molecule, bond_pair_ids, subunits: dict) -> tuple[Molecule, Result]:
        """Final molecule and every step's record + log (reference optimizer.py:308-390)."""
        result = Result(start_time=time.time())
        result.update_log(self._output_top_lines())
        result.update_log(f"There are {len(bond_pair_ids)} bonds to optimize.\n")
        result.update_log(
            f"There are {len(subunits)} sub units with N atoms:\n"
            f"{[len(subunits[i]) for i in subunits]}\n"
        )
        rule = "=" * 52 + "\n"
        result.update_log(rule + " " * 17 + "Running optimisation!" + " " * 14 + "\n" + rule + "\n")
        result.update_log("step system_potential nonbond_potential max_dist opt_bbs updated?\n")

        topo, out = self._run_single(mol, bond_pair_ids, subunits, want_trace=True)
        trace = out.trace_potentials[0].cpu().numpy()
        bond, passed, _, _ = unpack_info(out.trace_info[0].cpu().numpy())
        frames = np.asarray(out.trajectory[0].cpu().numpy(), dtype=np.float64)
        # the log prints the chosen bond as numpy prints a row of the bond table: format each row once
        bond_rows = [str(row) for row in np.asarray(bond_pair_ids)]
        trace = trace.tolist()
        for step in range(len(trace)):
            u, unb, far = trace[step]
            if passed[step] < 0:
                flag, row = None, None
            else:
                flag, row = bool(passed[step]), bond_rows[int(bond[step])]
            result.add_step_result(
                MCStepResult(
                    step=step, position_matrix=frames[step], passed=flag, system_potential=u,
                    nonbonded_potential=unb, max_bond_distance=far,
                    log=self._step_log(step, u, unb, far, row, flag),
                )
            )
        num_passed = result.get_number_passed()
        result.update_log(
            "\n============================================\n"
            "Optimisation done:\n"
            f"{num_passed} steps passed: {(num_passed / self._num_steps) * 100}%\n"
            f"Total optimisation time: {round(result.get_timing(time.time()), 4)}s\n"
            "============================================\n"
        )
        final = out.positions[0].cpu().numpy().astype(np.float64)
        return mol.with_position_matrix(final), result

    # ------------------------------------------------------------------ many chains
    def prepare_batch(
        self,
        mol: Molecule,
        bond_pair_ids,
        subunits: dict,
        n_chains: int,
        devices: Sequence[int] | None = None,
        io_dtype=np.float64,
        per_chain_positions: bool = False,
        threads_per_chain: int = 0,
        n_slices: int = 0,
        early_reject: bool | None = None,
    ) -> BatchPlan:
        """Lower the topology once and allocate pinned/device buffers for ``n_chains`` chains.

        ``plan.run(seeds=..., positions=...)`` then costs one H2D copy, one kernel launch per
        device and one D2H copy.  See :class:`mchammer_b200.batch.BatchPlan`.
        """
        native.require_cuda()
        topo = lower_topology(
            mol.get_num_atoms(), bond_pair_ids, subunits,
            require_bond_subunits=self._num_steps > 1, allow_overlap=True,
        )
        if devices is None:
            devices = [self._torch_device().index]
        return BatchPlan(
            topo, mol.get_position_matrix(), self._params(), self._num_steps, self._precision,
            n_chains, list(devices), io_dtype, per_chain_positions, threads_per_chain, n_slices,
            early_reject,
        )

    def get_results_batch(
        self,
        mol: Molecule,
        bond_pair_ids,
        subunits: dict,
        seeds: Sequence[int] | None = None,
        n_chains: int | None = None,
        positions: np.ndarray | None = None,
        devices: Sequence[int] | None = None,
        io_dtype=np.float64,
        threads_per_chain: int = 0,
        early_reject: bool | None = None,
    ) -> BatchResult:
        """Run many independent chains of the same topology (NEW; no reference counterpart).

        Chain ``c`` is what ``Optimizer(..., random_seed=seeds[c]).get_result(...)`` would
        do to molecule ``c``: ``positions[c]`` if ``positions`` ([C,N,3]) is given, else
        ``mol`` itself for every chain.  ``seeds`` default to consecutive integers starting
        at a value drawn from this optimizer's generator.  Chains are split contiguously
        over ``devices`` (default: this optimizer's device); there is no inter-device
        communication -- results are concatenated on the host.  ``early_reject`` (None: the precision
        mode's default) lets the kernel reject proposals on a pair-free bound of the energy change
        without their pair sweep; results are the same bits either way (``BatchResult.pair_evals``
        shows the difference).
        """
        native.require_cuda()
        if positions is not None:
            positions = np.ascontiguousarray(positions, dtype=io_dtype)
            if positions.ndim != 3 or positions.shape[1:] != (mol.get_num_atoms(), 3):
                raise ValueError("positions must have shape [C, n_atoms, 3]")
            if n_chains is None:
                n_chains = positions.shape[0]
        if seeds is not None:
            seeds = [int(s) for s in seeds]
            if n_chains is None:
                n_chains = len(seeds)
        if n_chains is None:
            raise ValueError("give seeds, positions or n_chains")
        if seeds is None:
            base = int(self._generator.integers(0, 2**63))
            seeds = [base + c for c in range(n_chains)]
        if len(seeds) != n_chains or (positions is not None and positions.shape[0] != n_chains):
            raise ValueError("seeds / positions / n_chains disagree on the number of chains")
        plan = self.prepare_batch(
            mol, bond_pair_ids, subunits, n_chains, devices, io_dtype,
            per_chain_positions=positions is not None, threads_per_chain=threads_per_chain,
            early_reject=early_reject,
        )
        return plan.run(seeds=seeds, positions=positions).copy()

    def get_trajectories_batch(
        self,
        mol: Molecule,
        bond_pair_ids,
        subunits: dict,
        seeds: Sequence[int],
        positions: np.ndarray | None = None,
        devices: Sequence[int] | None = None,
        io_dtype=np.float64,
        want_positions: bool = True,
        slice_bytes: int = 1 << 30,
    ) -> BatchTrajectory:
        """Every step of many independent chains (NEW): what ``get_trajectory`` records, for
        chain ``c`` = ``Optimizer(..., random_seed=seeds[c])`` on ``positions[c]`` (or ``mol``).

        Returns per-step potentials, accept flags, bond choices and -- unless
        ``want_positions=False`` -- the frames ``[C, num_steps, N, 3]`` in ``io_dtype``, streamed
        off the device slice by slice (see :meth:`BatchPlan.run_trajectories`).
        ``trajectory_result(traj, c, bond_pair_ids, subunits)`` rebuilds the reference-style
        :class:`Result` (records + log text) of one chain.
        """
        native.require_cuda()
        seeds = [int(s) for s in seeds]
        if positions is not None:
            positions = np.ascontiguousarray(positions, dtype=io_dtype)
            if positions.shape != (len(seeds), mol.get_num_atoms(), 3):
                raise ValueError("positions must have shape [len(seeds), n_atoms, 3]")
        plan = self.prepare_batch(
            mol, bond_pair_ids, subunits, len(seeds), devices, io_dtype,
            per_chain_positions=positions is not None,
        )
        # the plan (and the pinned buffers the arrays live in) is owned by the returned object
        traj = plan.run_trajectories(seeds=seeds, positions=positions, want_positions=want_positions,
                                     slice_bytes=slice_bytes)
        traj.plan = plan
        return traj

    def trajectory_result(self, traj: BatchTrajectory, chain: int, bond_pair_ids, subunits: dict) -> Result:
        """The reference-style ``Result`` (step records and the full log text, optimizer.py:308-390)
        of chain ``chain`` of a :class:`BatchTrajectory`."""
        result = Result(start_time=time.time())
        result.update_log(self._output_top_lines())
        result.update_log(f"There are {len(bond_pair_ids)} bonds to optimize.\n")
        result.update_log(
            f"There are {len(subunits)} sub units with N atoms:\n"
            f"{[len(subunits[i]) for i in subunits]}\n"
        )
        rule = "=" * 52 + "\n"
        result.update_log(rule + " " * 17 + "Running optimisation!" + " " * 14 + "\n" + rule + "\n")
        result.update_log("step system_potential nonbond_potential max_dist opt_bbs updated?\n")
        bond_rows = [str(row) for row in np.asarray(bond_pair_ids)]
        us, unbs = traj.system_potential[chain].tolist(), traj.nonbonded_potential[chain].tolist()
        fars, flags = traj.max_bond_distance[chain].tolist(), traj.passed[chain].tolist()
        picks = traj.bond_index[chain].tolist()
        for step in range(traj.num_steps):
            u, unb, far = us[step], unbs[step], fars[step]
            if flags[step] < 0:
                flag, row = None, None
            else:
                flag, row = bool(flags[step]), bond_rows[picks[step]]
            frame = None if traj.positions is None else np.array(traj.positions[chain, step], dtype=np.float64)
            result.add_step_result(
                MCStepResult(
                    step=step, position_matrix=frame, passed=flag, system_potential=u,
                    nonbonded_potential=unb, max_bond_distance=far,
                    log=self._step_log(step, u, unb, far, row, flag),
                )
            )
        num_passed = result.get_number_passed()
        result.update_log(
            "\n============================================\n"
            "Optimisation done:\n"
            f"{num_passed} steps passed: {(num_passed / self._num_steps) * 100}%\n"
            f"Total optimisation time: {round(traj.seconds, 4)}s\n"
            "============================================\n"
        )
        return result
