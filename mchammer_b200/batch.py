"""Batched Metropolis chains: a prepared plan that owns pinned host staging buffers and the
per-device buffers, so repeated runs pay only H2D copy + kernel + D2H copy.

Independent chains shard contiguously over the devices of one box; there is no collective
and no peer traffic on the data path.  One host thread drives all
devices: copies and launches are asynchronous on each device's current stream, results
are read back device by device.
"""

from __future__ import annotations

import time
from dataclasses import dataclass
from typing import Sequence

import numpy as np

from . import engine, native, rngstate
from .lowering import LoweredTopology


@dataclass
class BatchResult:
    """Outcome of ``C`` independent chains (host arrays)."""

    positions: np.ndarray            # [C,N,3]
    system_potential: np.ndarray     # [C]
    nonbonded_potential: np.ndarray  # [C]
    max_bond_distance: np.ndarray    # [C]
    n_accepted: np.ndarray           # [C] accepted moves
    last_passed: np.ndarray          # [C] 1/0; -1 when no move was made
    last_bond_index: np.ndarray      # [C] index into bond_pair_ids; -1 when no move
    pair_evals: np.ndarray           # [C] pair terms the kernel evaluated
    num_steps: int
    seconds: float                   # wall time of the call (H2D + kernel + D2H)
    h2d_bytes: int = 0
    d2h_bytes: int = 0

    def __len__(self) -> int:
        return int(self.positions.shape[0])

    def molecule(self, index: int, template):
        return template.with_position_matrix(np.array(self.positions[index], dtype=np.float64))

    def copy(self) -> "BatchResult":
        fields = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in self.__dict__.items()}
        return BatchResult(**fields)


@dataclass
class BatchTrajectory:
    """Every step of ``C`` independent chains (host arrays; step 0 is the start geometry)."""

    positions: np.ndarray | None     # [C,steps,N,3] io dtype (None when frames were not requested)
    system_potential: np.ndarray     # [C,steps]
    nonbonded_potential: np.ndarray  # [C,steps]
    max_bond_distance: np.ndarray    # [C,steps]
    passed: np.ndarray               # [C,steps] 1/0; -1 at step 0
    bond_index: np.ndarray           # [C,steps] index into bond_pair_ids; -1 at step 0
    final_positions: np.ndarray      # [C,N,3]
    num_steps: int
    seconds: float
    d2h_bytes: int = 0
    plan: object | None = None       # keeps the pinned host buffers alive when the caller dropped the plan

    def __len__(self) -> int:
        return int(self.system_potential.shape[0])


def unpack_info(info: np.ndarray):
    """Split trace words: bond | passed << 16 | kind << 17 | moved subunit << 18 (-1: no move)."""
    info = np.asarray(info, dtype=np.int64)
    none = info < 0
    bond = np.where(none, -1, info & 0xFFFF)
    passed = np.where(none, -1, (info >> 16) & 1)
    kind = np.where(none, -1, (info >> 17) & 1)
    moved = np.where(none, -1, (info >> 18) & 0x1FFF)
    return bond, passed, kind, moved


def shard_bounds(n_items: int, n_shards: int) -> list[tuple[int, int]]:
    """Contiguous, balanced ``[lo, hi)`` ranges: shard ``g`` gets items ``g*n/G .. (g+1)*n/G``."""
    if n_shards <= 0:
        raise ValueError("n_shards must be positive")
    return [((g * n_items) // n_shards, ((g + 1) * n_items) // n_shards) for g in range(n_shards)]


class _Shard:
    """Buffers of one device."""

    def __init__(self, torch, topo: LoweredTopology, device_index: int, lo: int, hi: int,
                 io_torch_dtype, per_chain_positions: bool):
        self.lo, self.hi, self.count = lo, hi, hi - lo
        self.device = torch.device("cuda", device_index)
        n = topo.n_atoms
        with torch.cuda.device(self.device):
            self.dtopo = engine.upload_topology(topo, device_index)
            rows = self.count if per_chain_positions else 1
            self.pos_in = torch.empty((rows, n, 3), dtype=io_torch_dtype, device=self.device)
            self.rng = torch.empty((self.count, 6), dtype=torch.int64, device=self.device)
            f64 = dict(dtype=torch.float64, device=self.device)
            i32 = dict(dtype=torch.int32, device=self.device)
            self.out = engine.OptimizeOutput(
                positions=torch.empty((self.count, n, 3), dtype=io_torch_dtype, device=self.device),
                system_potential=torch.empty(self.count, **f64),
                nonbonded_potential=torch.empty(self.count, **f64),
                max_bond_distance=torch.empty(self.count, **f64),
                n_accepted=torch.empty(self.count, **i32), last_step_info=torch.empty(self.count, **i32),
                rng_state=self.rng, pair_evals=torch.empty(self.count, dtype=torch.int64, device=self.device),
                trace_potentials=None, trace_info=None, trajectory=None,
            )
            self.streams = []  # side streams of the pipelined run, created on first use

    def out_slice(self, lo: int, hi: int) -> "engine.OptimizeOutput":
        o = self.out
        return engine.OptimizeOutput(
            positions=o.positions[lo:hi], system_potential=o.system_potential[lo:hi],
            nonbonded_potential=o.nonbonded_potential[lo:hi], max_bond_distance=o.max_bond_distance[lo:hi],
            n_accepted=o.n_accepted[lo:hi], last_step_info=o.last_step_info[lo:hi],
            rng_state=self.rng[lo:hi], pair_evals=o.pair_evals[lo:hi],
            trace_potentials=None, trace_info=None, trajectory=None,
        )


class BatchPlan:
    """Prepared batched run of one topology: ``plan.run(seeds, positions)`` -> BatchResult.

    The arrays of the returned BatchResult are views of this plan's pinned host buffers
    and are overwritten by the next ``run``; call ``.copy()`` to keep them.
    """

    def __init__(self, topo: LoweredTopology, base_positions: np.ndarray, params, num_steps: int,
                 precision: str, n_chains: int, devices: Sequence[int], io_dtype=np.float64,
                 per_chain_positions: bool = False, threads_per_chain: int = 0, n_slices: int = 0,
                 early_reject: bool | None = None):
        torch = native.require_cuda()
        native.lib()
        self._torch = torch
        self.topo = topo
        self.params = params
        self.num_steps = max(int(num_steps), 1)
        self.precision = precision
        self.n_chains = int(n_chains)
        self.threads_per_chain = int(threads_per_chain)
        # None: the precision mode's default (engine.resolve_early_reject); never changes a result
        self.early_reject = early_reject
        self.per_chain_positions = bool(per_chain_positions)
        # `run` cuts every device's chains into slices and pipelines H2D / kernel / D2H of
        # consecutive slices on side streams (0 = choose: 4 slices once the copies are worth hiding)
        self.n_slices = int(n_slices)
        self.io_dtype = np.dtype(io_dtype)
        if self.io_dtype not in (np.dtype(np.float64), np.dtype(np.float32)):
            raise TypeError("io_dtype must be float64 or float32")
        tdtype = torch.float64 if self.io_dtype == np.float64 else torch.float32
        n = topo.n_atoms
        c = self.n_chains

        def pinned(shape, dtype):
            return torch.empty(shape, dtype=dtype).pin_memory()

        # pinned host staging: inputs ...
        self.h_pos_in = pinned((c if per_chain_positions else 1, n, 3), tdtype)
        self.h_rng = pinned((c, 6), torch.int64)
        # ... and outputs
        self.h_pos_out = pinned((c, n, 3), tdtype)
        self.h_f64 = pinned((3, c), torch.float64)   # U, U_nb, max bond
        self.h_i32 = pinned((2, c), torch.int32)     # n_accepted, last_step_info
        self.h_pairs = pinned((c,), torch.int64)
        self.h_pos_in[0].copy_(torch.from_numpy(np.ascontiguousarray(base_positions, dtype=self.io_dtype)))
        if per_chain_positions:
            self.h_pos_in[:] = self.h_pos_in[0]
        self.shards = [
            _Shard(torch, topo, dev, lo, hi, tdtype, per_chain_positions)
            for dev, (lo, hi) in zip(devices, shard_bounds(c, len(devices))) if hi > lo
        ]

    # -- the three phases, separately callable so a benchmark can time the device part alone
    def stage_inputs(self, seeds: Sequence[int] | None = None, positions=None,
                     rng_words: np.ndarray | None = None) -> int:
        """Fill the pinned input buffers on the host.  Returns the bytes that will go H2D."""
        torch = self._torch
        if rng_words is None:
            if seeds is None:
                raise ValueError("give seeds or rng_words")
            if len(seeds) != self.n_chains:
                raise ValueError(f"expected {self.n_chains} seeds, got {len(seeds)}")
            rng_words = rngstate.seed_states(seeds)
        self.h_rng.copy_(torch.from_numpy(np.ascontiguousarray(rng_words, dtype=np.uint64).view(np.int64)))
        if positions is not None:
            if not self.per_chain_positions:
                raise ValueError("this plan was prepared without per-chain positions")
            src = positions if torch.is_tensor(positions) else torch.from_numpy(
                np.ascontiguousarray(positions, dtype=self.io_dtype))
            if tuple(src.shape) != tuple(self.h_pos_in.shape):
                raise ValueError(f"positions must be {tuple(self.h_pos_in.shape)}, got {tuple(src.shape)}")
            if src.data_ptr() != self.h_pos_in.data_ptr():
                self.h_pos_in.copy_(src)
        return self.h_rng.numel() * 8 + self.h_pos_in.numel() * self.h_pos_in.element_size()

    def upload(self) -> None:
        torch = self._torch
        for sh in self.shards:
            with torch.cuda.device(sh.device):
                if self.per_chain_positions:
                    sh.pos_in.copy_(self.h_pos_in[sh.lo:sh.hi], non_blocking=True)
                else:
                    sh.pos_in.copy_(self.h_pos_in, non_blocking=True)
                sh.rng.copy_(self.h_rng[sh.lo:sh.hi], non_blocking=True)

    def launch(self) -> None:
        torch = self._torch
        for sh in self.shards:
            with torch.cuda.device(sh.device):
                sh.out = engine.optimize_on_device(
                    sh.dtopo, sh.pos_in, sh.count, sh.rng, self.params, self.num_steps,
                    self.precision, threads_per_chain=self.threads_per_chain, out=sh.out,
                    early_reject=self.early_reject,
                )

    def download(self) -> int:
        """Device -> pinned host (async per device), then wait.  Returns the bytes read back."""
        torch = self._torch
        for sh in self.shards:
            with torch.cuda.device(sh.device):
                o = sh.out
                self.h_pos_out[sh.lo:sh.hi].copy_(o.positions, non_blocking=True)
                self.h_f64[0, sh.lo:sh.hi].copy_(o.system_potential, non_blocking=True)
                self.h_f64[1, sh.lo:sh.hi].copy_(o.nonbonded_potential, non_blocking=True)
                self.h_f64[2, sh.lo:sh.hi].copy_(o.max_bond_distance, non_blocking=True)
                self.h_i32[0, sh.lo:sh.hi].copy_(o.n_accepted, non_blocking=True)
                self.h_i32[1, sh.lo:sh.hi].copy_(o.last_step_info, non_blocking=True)
                self.h_pairs[sh.lo:sh.hi].copy_(o.pair_evals, non_blocking=True)
        for sh in self.shards:
            torch.cuda.current_stream(sh.device).synchronize()
        return (self.h_pos_out.numel() * self.h_pos_out.element_size() + self.h_f64.numel() * 8
                + self.h_i32.numel() * 4 + self.h_pairs.numel() * 8)

    def result(self, seconds: float = 0.0, h2d: int = 0, d2h: int = 0) -> BatchResult:
        bond, passed, _, _ = unpack_info(self.h_i32[1].numpy())
        return BatchResult(
            positions=self.h_pos_out.numpy(),
            system_potential=self.h_f64[0].numpy(), nonbonded_potential=self.h_f64[1].numpy(),
            max_bond_distance=self.h_f64[2].numpy(), n_accepted=self.h_i32[0].numpy().astype(np.int64),
            last_passed=passed, last_bond_index=bond, pair_evals=self.h_pairs.numpy(),
            num_steps=self.num_steps, seconds=seconds, h2d_bytes=h2d, d2h_bytes=d2h,
        )

    def _slices(self, count: int) -> list[tuple[int, int]]:
        # automatic: pipeline in 4 slices once a device's share of the copies is worth hiding (>= 16 MiB)
        per_chain = self.topo.n_atoms * 3 * self.h_pos_out.element_size()
        n = self.n_slices if self.n_slices > 0 else (4 if count >= 8 and count * per_chain >= (16 << 20) else 1)
        n = max(1, min(n, count))
        return [b for b in shard_bounds(count, n) if b[1] > b[0]]

    def _run_pipelined(self) -> int:
        """H2D -> kernel -> D2H per slice on side streams, so the copies of one slice overlap
        the kernel of another.  Returns the bytes read back."""
        torch = self._torch
        for sh in self.shards:
            with torch.cuda.device(sh.device):
                main = torch.cuda.current_stream(sh.device)
                slices = self._slices(sh.count)
                while len(sh.streams) < min(3, len(slices)):
                    sh.streams.append(torch.cuda.Stream(device=sh.device))
                if not self.per_chain_positions:
                    sh.pos_in.copy_(self.h_pos_in, non_blocking=True)
                for st in sh.streams:
                    st.wait_stream(main)
                for k, (lo, hi) in enumerate(slices):
                    with torch.cuda.stream(sh.streams[k % len(sh.streams)]):
                        g_lo, g_hi = sh.lo + lo, sh.lo + hi
                        if self.per_chain_positions:
                            sh.pos_in[lo:hi].copy_(self.h_pos_in[g_lo:g_hi], non_blocking=True)
                            pos = sh.pos_in[lo:hi]
                        else:
                            pos = sh.pos_in
                        sh.rng[lo:hi].copy_(self.h_rng[g_lo:g_hi], non_blocking=True)
                        o = engine.optimize_on_device(
                            sh.dtopo, pos, hi - lo, sh.rng[lo:hi], self.params, self.num_steps,
                            self.precision, threads_per_chain=self.threads_per_chain, out=sh.out_slice(lo, hi),
                            resident_chains_hint=sh.count,  # the slices run concurrently: choices are the batch's
                            early_reject=self.early_reject,
                        )
                        self.h_pos_out[g_lo:g_hi].copy_(o.positions, non_blocking=True)
                        self.h_f64[0, g_lo:g_hi].copy_(o.system_potential, non_blocking=True)
                        self.h_f64[1, g_lo:g_hi].copy_(o.nonbonded_potential, non_blocking=True)
                        self.h_f64[2, g_lo:g_hi].copy_(o.max_bond_distance, non_blocking=True)
                        self.h_i32[0, g_lo:g_hi].copy_(o.n_accepted, non_blocking=True)
                        self.h_i32[1, g_lo:g_hi].copy_(o.last_step_info, non_blocking=True)
                        self.h_pairs[g_lo:g_hi].copy_(o.pair_evals, non_blocking=True)
        for sh in self.shards:
            for st in sh.streams:
                st.synchronize()
                torch.cuda.current_stream(sh.device).wait_stream(st)
        return (self.h_pos_out.numel() * self.h_pos_out.element_size() + self.h_f64.numel() * 8
                + self.h_i32.numel() * 4 + self.h_pairs.numel() * 8)

    def run_trajectories(self, seeds: Sequence[int] | None = None, positions=None,
                         rng_words: np.ndarray | None = None, want_positions: bool = True,
                         slice_bytes: int = 1 << 30, pin_limit_bytes: int = 8 << 30) -> BatchTrajectory:
        """Like ``run`` but keeps every step: per-step potentials / accept flags and, if
        ``want_positions``, the frames ``[C, num_steps, N, 3]`` (reference ``get_trajectory``,
        results.py:149-152, for a whole batch).

        Frames are the only bulky output of the path (``num_steps * N * 3`` values per chain),
        so they are streamed: every device works through its chains in slices of at most
        ``slice_bytes`` of frames, two device frame buffers alternate on two streams, and the
        device->host copy of one slice runs under the kernel of the next.  The host array is
        pinned when it is at most ``pin_limit_bytes`` (asynchronous copies), pageable above; like
        the buffers of ``run`` it belongs to the plan and is overwritten by the next call.
        """
        torch = self._torch
        start = time.perf_counter()
        self.stage_inputs(seeds, positions, rng_words)
        c, n, steps = self.n_chains, self.topo.n_atoms, self.num_steps
        tdtype = self.h_pos_in.dtype
        item = self.h_pos_in.element_size()
        frame_bytes = steps * n * 3 * item

        def host(name, shape, dtype, nbytes):
            # host buffers are kept with the plan (pinning gigabytes costs more than the copy)
            cache = self.__dict__.setdefault("_traj_host", {})
            if name not in cache:
                t = torch.empty(shape, dtype=dtype)
                cache[name] = t.pin_memory() if nbytes <= pin_limit_bytes else t
            return cache[name]

        h_traj = host("frames", (c, steps, n, 3), tdtype, c * frame_bytes) if want_positions else None
        h_trace = host("trace", (c, steps, 3), torch.float64, 0)
        h_info = host("info", (c, steps), torch.int32, 0)
        for sh in self.shards:
            with torch.cuda.device(sh.device):
                main = torch.cuda.current_stream(sh.device)
                per = max(1, min(sh.count, int(slice_bytes // max(frame_bytes, 1)) if want_positions else sh.count))
                while len(sh.streams) < 2:
                    sh.streams.append(torch.cuda.Stream(device=sh.device))
                lanes = sh.streams[:2]
                f64 = dict(dtype=torch.float64, device=sh.device)
                bufs = [dict(
                    traj=torch.empty((per, steps, n, 3), dtype=tdtype, device=sh.device) if want_positions else None,
                    trace=torch.empty((per, steps, 3), **f64),
                    info=torch.empty((per, steps), dtype=torch.int32, device=sh.device),
                ) for _ in lanes]
                if not self.per_chain_positions:
                    sh.pos_in.copy_(self.h_pos_in, non_blocking=True)
                for st in lanes:
                    st.wait_stream(main)
                for k, lo in enumerate(range(0, sh.count, per)):
                    hi = min(lo + per, sh.count)
                    m = hi - lo
                    buf = bufs[k % 2]
                    with torch.cuda.stream(lanes[k % 2]):
                        g_lo, g_hi = sh.lo + lo, sh.lo + hi
                        if self.per_chain_positions:
                            sh.pos_in[lo:hi].copy_(self.h_pos_in[g_lo:g_hi], non_blocking=True)
                            pos = sh.pos_in[lo:hi]
                        else:
                            pos = sh.pos_in
                        sh.rng[lo:hi].copy_(self.h_rng[g_lo:g_hi], non_blocking=True)
                        o = sh.out_slice(lo, hi)
                        o.trace_potentials, o.trace_info = buf["trace"][:m], buf["info"][:m]
                        o.trajectory = buf["traj"][:m] if want_positions else None
                        engine.optimize_on_device(
                            sh.dtopo, pos, m, sh.rng[lo:hi], self.params, steps, self.precision,
                            threads_per_chain=self.threads_per_chain, out=o, resident_chains_hint=sh.count,
                        )
                        if want_positions:
                            h_traj[g_lo:g_hi].copy_(o.trajectory, non_blocking=True)
                        h_trace[g_lo:g_hi].copy_(o.trace_potentials, non_blocking=True)
                        h_info[g_lo:g_hi].copy_(o.trace_info, non_blocking=True)
                        self.h_pos_out[g_lo:g_hi].copy_(o.positions, non_blocking=True)
        for sh in self.shards:
            for st in sh.streams[:2]:
                st.synchronize()
                torch.cuda.current_stream(sh.device).wait_stream(st)
        trace = h_trace.numpy()
        bond, passed, _, _ = unpack_info(h_info.numpy())
        d2h = ((c * frame_bytes if want_positions else 0) + h_trace.numel() * 8 + h_info.numel() * 4
               + self.h_pos_out.numel() * item)
        return BatchTrajectory(
            positions=h_traj.numpy() if want_positions else None,
            system_potential=trace[:, :, 0], nonbonded_potential=trace[:, :, 1],
            max_bond_distance=trace[:, :, 2], passed=passed, bond_index=bond,
            final_positions=self.h_pos_out.numpy().copy(), num_steps=steps,
            seconds=time.perf_counter() - start, d2h_bytes=d2h,
        )

    def run(self, seeds: Sequence[int] | None = None, positions=None,
            rng_words: np.ndarray | None = None) -> BatchResult:
        """Host buffers in, host buffers out: stage -> (H2D -> kernel -> D2H, pipelined by slice)."""
        start = time.perf_counter()
        h2d = self.stage_inputs(seeds, positions, rng_words)
        d2h = self._run_pipelined()
        return self.result(time.perf_counter() - start, h2d, d2h)
