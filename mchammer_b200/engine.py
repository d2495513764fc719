"""Device engine: stages arrays on a B200 and launches the C-ABI kernels.

This is the only module that touches the GPU.  PyTorch provides device memory, pinned
host buffers and streams; the arithmetic is entirely in libmchammer_b200.so.
Everything here raises if CUDA or the library is unavailable -- there is no CPU path.
"""

from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Sequence

import numpy as np

from . import native
from .lowering import LoweredTopology

PRECISIONS = {"fp64": native.MCH_F64, "fp32": native.MCH_F32}


def _dtype_code(torch, dtype) -> int:
    if dtype == torch.float64:
        return native.MCH_F64
    if dtype == torch.float32:
        return native.MCH_F32
    raise TypeError(f"positions must be float64 or float32, got {dtype}")


def _precision_code(precision: str) -> int:
    try:
        return PRECISIONS[precision]
    except KeyError:
        raise ValueError(f"precision must be 'fp64' or 'fp32', got {precision!r}") from None


def make_params(step_size, target_bond_length, bond_epsilon, nonbond_epsilon, nonbond_sigma,
                nonbond_mu, beta) -> native.MchParams:
    return native.MchParams(
        float(step_size), float(target_bond_length), float(bond_epsilon),
        float(nonbond_epsilon), float(nonbond_sigma), float(nonbond_mu), float(beta),
    )


@dataclass
class DeviceTopology:
    """Topology arrays resident on one device (tiny; shared by every chain)."""

    host: LoweredTopology
    device: object
    perm: object
    su_offsets: object
    bonds: object
    bond_su: object
    n_heavy: object
    # explicit moving sets for the split-state kernel (csrc/mch_wide.cuh)
    bond_slots: object = None
    row_offsets: object = None
    row_slots: object = None
    row_weights: object = None
    run_offsets: object = None
    runs: object = None


def upload_topology(topo: LoweredTopology, device=None) -> DeviceTopology:
    """All topology arrays in ONE host buffer and one host-to-device copy (they are tiny: a copy each used to
    cost more than the arrays' bytes); the fields are 16-byte aligned views of that buffer."""
    torch = native.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    slot_of = np.empty(topo.n_atoms, dtype=np.int64)
    slot_of[topo.perm] = np.arange(topo.n_atoms)
    parts = dict(
        perm=topo.perm, su_offsets=topo.su_offsets, bonds=topo.bonds.reshape(-1), bond_su=topo.bond_su.reshape(-1),
        n_heavy=topo.n_heavy,
        bond_slots=slot_of[topo.bonds.reshape(-1)] if topo.n_bonds else np.zeros(0, dtype=np.int32),
        row_offsets=topo.row_offsets, row_slots=topo.row_slots, row_weights=topo.row_weights,
        run_offsets=topo.run_offsets, runs=topo.runs.reshape(-1),
    )
    offsets, total = {}, 0
    for name, a in parts.items():
        offsets[name] = total
        total += (int(np.size(a)) + 3) // 4 * 4
    packed = np.zeros(max(total, 4), dtype=np.int32)
    for name, a in parts.items():
        n = int(np.size(a))
        packed[offsets[name]:offsets[name] + n] = np.asarray(a).reshape(-1)
    buf = torch.from_numpy(packed).to(dev)
    views = {name: buf[offsets[name]:offsets[name] + int(np.size(a))] for name, a in parts.items()}
    return DeviceTopology(host=topo, device=dev, **views)


_TOPOLOGY_CACHE: "dict[tuple, DeviceTopology]" = {}
_TOPOLOGY_CACHE_SIZE = 16


def cached_topology(n_atoms: int, bond_pair_ids, subunits, device_index: int, elements=None,
                    require_bond_subunits: bool = True, allow_overlap: bool = False) -> DeviceTopology:
    """``upload_topology(lower_topology(...))`` with a small LRU cache: a user who optimises many conformers of
    one molecule (same bonds, same subunits) pays the lowering and the upload once.  The key holds the exact
    inputs (bond pairs in the caller's order, subunit keys and member lists in dict order), so a hit returns
    what a fresh lowering would."""
    from .lowering import lower_topology

    try:
        bonds_key = np.asarray(bond_pair_ids, dtype=np.int64).tobytes()
        keys = tuple(subunits)
        lens = tuple(len(subunits[k]) for k in keys)
        members = np.fromiter((a for k in keys for a in subunits[k]), dtype=np.int64, count=sum(lens)).tobytes()
        key = (int(n_atoms), int(device_index), bool(require_bond_subunits), bool(allow_overlap),
               None if elements is None else tuple(elements), bonds_key, keys, lens, members)
        hash(key)
    except (TypeError, ValueError):  # unhashable subunit keys, ragged bond lists, ...: let the lowering report it
        key = None
    if key is not None and key in _TOPOLOGY_CACHE:
        dtopo = _TOPOLOGY_CACHE.pop(key)
        _TOPOLOGY_CACHE[key] = dtopo  # most recently used last
        return dtopo
    topo = lower_topology(n_atoms, bond_pair_ids, subunits, elements=elements,
                          require_bond_subunits=require_bond_subunits, allow_overlap=allow_overlap)
    dtopo = upload_topology(topo, device_index)
    if key is not None:
        _TOPOLOGY_CACHE[key] = dtopo
        while len(_TOPOLOGY_CACHE) > _TOPOLOGY_CACHE_SIZE:
            _TOPOLOGY_CACHE.pop(next(iter(_TOPOLOGY_CACHE)))
    return dtopo


# default of the `early_reject` option per precision mode (see resolve_early_reject)
EARLY_REJECT_DEFAULT = {"fp64": True, "fp32": False}


@dataclass
class OptimizeOutput:
    """Device tensors produced by one ``mch_optimize_batch`` launch."""

    positions: object            # [C,N,3] io dtype
    system_potential: object     # [C] f64
    nonbonded_potential: object  # [C] f64
    max_bond_distance: object    # [C] f64
    n_accepted: object           # [C] i32
    last_step_info: object       # [C] i32
    rng_state: object            # [C,6] i64 view of uint64 words (advanced)
    pair_evals: object           # [C] i64
    trace_potentials: object | None  # [C,steps,3] f64
    trace_info: object | None        # [C,steps] i32
    trajectory: object | None        # [C,steps,N,3] io dtype


def resolve_early_reject(early_reject: bool | None, precision: str) -> bool:
    """Whether a launch asks for early rejection (``MCH_FLAG_EARLY_REJECTION``).

    Early rejection never changes a result -- the library decides a proposal without its pair sweep only when a
    pair-free lower bound of the energy change already rejects it, and ignores the flag for launches it does not
    cover (records per step, clusters, one-warp chains, mu != 3) -- so the choice is one of speed only.  An explicit
    ``early_reject`` wins; then the environment variable ``MCHAMMER_B200_EARLY`` (``1`` / ``0``); the default is on
    in the fp64 mode (measured +24 % on the bench workload) and off in the fp32 mode (+5 % for full batches, -3 % at 1024 chains, slower on small batches).
    """
    if early_reject is not None:
        return bool(early_reject)
    env = os.environ.get("MCHAMMER_B200_EARLY")
    if env in ("0", "1"):
        return env == "1"
    return EARLY_REJECT_DEFAULT[precision]


def optimize_on_device(
    dtopo: DeviceTopology,
    pos_in,                      # torch tensor on dtopo.device: [C,N,3] or [1,N,3] / [N,3]
    n_chains: int,
    rng_state,                   # torch int64 tensor [C,6] on device (uint64 bit patterns); updated in place
    params: native.MchParams,
    num_steps: int,
    precision: str = "fp64",
    inexact_undo: bool | None = None,
    want_trace: bool = False,
    want_trajectory: bool = False,
    threads_per_chain: int = 0,
    out: OptimizeOutput | None = None,
    cluster_size: int = 0,
    resident_chains_hint: int = 0,
    force_split_state: bool = False,
    early_reject: bool | None = None,
) -> OptimizeOutput:
    """Launch the resident Metropolis kernel for ``n_chains`` chains (asynchronous)."""
    torch = native.require_cuda()
    lib = native.lib()
    topo = dtopo.host
    dev = dtopo.device
    n = topo.n_atoms
    if topo.n_bonds == 0:
        raise ValueError("bond_pair_ids is empty: there is no bond to optimize")
    if pos_in.dim() == 2:
        pos_in = pos_in.unsqueeze(0)
    if pos_in.shape[-2:] != (n, 3) or pos_in.shape[0] not in (1, n_chains):
        raise ValueError(f"positions must be [{n_chains} or 1, {n}, 3], got {tuple(pos_in.shape)}")
    if not pos_in.is_contiguous():
        pos_in = pos_in.contiguous()
    io_code = _dtype_code(torch, pos_in.dtype)
    compute_code = _precision_code(precision)
    if inexact_undo is None:
        inexact_undo = compute_code == native.MCH_F64
    steps = max(int(num_steps), 1)

    with torch.cuda.device(dev):
        if out is None:
            f64 = dict(dtype=torch.float64, device=dev)
            out = OptimizeOutput(
                positions=torch.empty((n_chains, n, 3), dtype=pos_in.dtype, device=dev),
                system_potential=torch.empty(n_chains, **f64),
                nonbonded_potential=torch.empty(n_chains, **f64),
                max_bond_distance=torch.empty(n_chains, **f64),
                n_accepted=torch.empty(n_chains, dtype=torch.int32, device=dev),
                last_step_info=torch.empty(n_chains, dtype=torch.int32, device=dev),
                rng_state=rng_state,
                pair_evals=torch.empty(n_chains, dtype=torch.int64, device=dev),
                trace_potentials=torch.empty((n_chains, steps, 3), **f64) if want_trace else None,
                trace_info=torch.empty((n_chains, steps), dtype=torch.int32, device=dev) if want_trace else None,
                trajectory=(torch.empty((n_chains, steps, n, 3), dtype=pos_in.dtype, device=dev)
                            if want_trajectory else None),
            )
        else:
            out.rng_state = rng_state

        def ptr(t):
            return None if t is None else C.c_void_p(t.data_ptr())

        desc = native.OptimizeDesc()
        desc.n_chains, desc.n_atoms = n_chains, n
        desc.n_subunits, desc.n_bonds = topo.n_subunits, topo.n_bonds
        desc.max_subunit_atoms, desc.num_steps = topo.max_subunit_atoms, steps
        desc.io_dtype, desc.compute_dtype = io_code, compute_code
        desc.inexact_undo = 1 if inexact_undo else 0
        desc.threads_per_chain = int(threads_per_chain)
        desc.cluster_size, desc.resident_chains_hint = int(cluster_size), int(resident_chains_hint)
        desc.pos_in = ptr(pos_in)
        desc.pos_in_chain_stride = 0 if pos_in.shape[0] == 1 and n_chains > 1 else n * 3
        desc.pos_out = ptr(out.positions)
        desc.perm, desc.su_offsets = ptr(dtopo.perm), ptr(dtopo.su_offsets)
        desc.bonds, desc.bond_su = ptr(dtopo.bonds), ptr(dtopo.bond_su)
        desc.params = params
        desc.rng_state = ptr(rng_state)
        desc.system_potential = ptr(out.system_potential)
        desc.nonbonded_potential = ptr(out.nonbonded_potential)
        desc.max_bond_distance = ptr(out.max_bond_distance)
        desc.n_accepted, desc.last_step_info = ptr(out.n_accepted), ptr(out.last_step_info)
        desc.trace_potentials, desc.trace_info = ptr(out.trace_potentials), ptr(out.trace_info)
        desc.trajectory, desc.pair_evals = ptr(out.trajectory), ptr(out.pair_evals)
        # explicit moving sets: needed for overlapping subunits (general = 1) and used by the library on
        # its own for molecules that do not fit one CTA; general = 2 forces the split-state kernel
        desc.general = 2 if force_split_state else (1 if topo.general else 0)
        desc.n_row_lists, desc.max_rows = topo.n_user_subunits, topo.max_rows
        desc.flags = native.MCH_FLAG_EARLY_REJECTION if resolve_early_reject(early_reject, precision) else 0
        desc.bond_slots = ptr(dtopo.bond_slots)
        desc.row_offsets, desc.row_slots = ptr(dtopo.row_offsets), ptr(dtopo.row_slots)
        desc.row_weights = ptr(dtopo.row_weights)
        desc.run_offsets, desc.runs = ptr(dtopo.run_offsets), ptr(dtopo.runs)
        stream = torch.cuda.current_stream(dev).cuda_stream
        native.check(lib.mch_optimize_batch(C.byref(desc), C.c_void_p(stream)), "mch_optimize_batch")
    return out


def potential_on_device(pos, bonds_dev, n_bonds: int, params: native.MchParams, precision: str = "fp64"):
    """(U, U_nb, max_bond) device tensors for molecules ``pos`` [M,N,3] on the GPU."""
    torch = native.require_cuda()
    lib = native.lib()
    if pos.dim() == 2:
        pos = pos.unsqueeze(0)
    if not pos.is_contiguous():
        pos = pos.contiguous()
    m, n, _ = pos.shape
    dev = pos.device
    with torch.cuda.device(dev):
        u = torch.empty(m, dtype=torch.float64, device=dev)
        unb = torch.empty(m, dtype=torch.float64, device=dev)
        far = torch.empty(m, dtype=torch.float64, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        native.check(
            lib.mch_potential_batch(
                C.c_void_p(pos.data_ptr()), _dtype_code(torch, pos.dtype), _precision_code(precision),
                m, n, C.c_void_p(bonds_dev.data_ptr()) if n_bonds else None, n_bonds,
                C.byref(params), C.c_void_p(u.data_ptr()), C.c_void_p(unb.data_ptr()),
                C.c_void_p(far.data_ptr()), C.c_void_p(stream),
            ),
            "mch_potential_batch",
        )
    return u, unb, far


@dataclass
class CollapseOutput:
    positions: object
    steps_run: object
    min_distance: object
    max_bond_distance: object
    status: object
    trace_max_bond: object | None
    trajectory: object | None


def collapse_on_device(
    dtopo: DeviceTopology, pos_in, n_mol: int, step_size: float, distance_threshold: float,
    scale_steps: bool, precision: str = "fp64", max_steps: int = 1_000_000,
    trace_capacity: int = 0, want_trajectory: bool = False, threads_per_mol: int = 0,
) -> CollapseOutput:
    torch = native.require_cuda()
    lib = native.lib()
    topo, dev, n = dtopo.host, dtopo.device, dtopo.host.n_atoms
    if pos_in.dim() == 2:
        pos_in = pos_in.unsqueeze(0)
    if pos_in.shape[-2:] != (n, 3) or pos_in.shape[0] not in (1, n_mol):
        raise ValueError(f"positions must be [{n_mol} or 1, {n}, 3], got {tuple(pos_in.shape)}")
    if not pos_in.is_contiguous():
        pos_in = pos_in.contiguous()
    with torch.cuda.device(dev):
        f64 = dict(dtype=torch.float64, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        out = CollapseOutput(
            positions=torch.empty((n_mol, n, 3), dtype=pos_in.dtype, device=dev),
            steps_run=torch.empty(n_mol, **i32), min_distance=torch.empty(n_mol, **f64),
            max_bond_distance=torch.empty(n_mol, **f64), status=torch.empty(n_mol, **i32),
            trace_max_bond=torch.zeros((n_mol, trace_capacity), **f64) if trace_capacity > 0 else None,
            trajectory=(torch.empty((n_mol, trace_capacity, n, 3), dtype=pos_in.dtype, device=dev)
                        if (want_trajectory and trace_capacity > 0) else None),
        )

        def ptr(t):
            return None if t is None else C.c_void_p(t.data_ptr())

        d = native.CollapseDesc()
        d.n_mol, d.n_atoms, d.n_subunits, d.n_bonds = n_mol, n, topo.n_subunits, topo.n_bonds
        d.io_dtype, d.compute_dtype = _dtype_code(torch, pos_in.dtype), _precision_code(precision)
        d.scale_steps, d.max_steps = int(bool(scale_steps)), int(max_steps)
        d.trace_capacity, d.threads_per_mol = int(trace_capacity), int(threads_per_mol)
        # atoms outside every subunit form a trailing static group that the Collapser must leave alone
        d.n_moving_subunits, d.n_heavy_total = topo.n_user_subunits, int(topo.n_heavy.sum())
        d.step_size, d.distance_threshold = float(step_size), float(distance_threshold)
        d.pos_in = ptr(pos_in)
        d.pos_in_mol_stride = 0 if pos_in.shape[0] == 1 and n_mol > 1 else n * 3
        d.pos_out = ptr(out.positions)
        d.perm, d.su_offsets, d.n_heavy = ptr(dtopo.perm), ptr(dtopo.su_offsets), ptr(dtopo.n_heavy)
        d.bonds = ptr(dtopo.bonds) if topo.n_bonds else None
        d.steps_run, d.min_distance = ptr(out.steps_run), ptr(out.min_distance)
        d.max_bond_distance, d.status = ptr(out.max_bond_distance), ptr(out.status)
        d.trace_max_bond, d.trajectory = ptr(out.trace_max_bond), ptr(out.trajectory)
        stream = torch.cuda.current_stream(dev).cuda_stream
        native.check(lib.mch_collapse_batch(C.byref(d), C.c_void_p(stream)), "mch_collapse_batch")
    return out


def host_to_device(array: np.ndarray, device, pin: bool = True):
    """numpy -> device tensor through a pinned staging buffer (async on the current stream)."""
    torch = native.require_cuda()
    t = torch.from_numpy(np.ascontiguousarray(array))
    if pin:
        t = t.pin_memory()
    return t.to(device, non_blocking=True)


def rng_words_to_device(words: np.ndarray, device):
    """uint64 [C,6] numpy -> int64 device tensor with the same bit patterns."""
    return host_to_device(np.ascontiguousarray(words, dtype=np.uint64).view(np.int64), device, pin=False)


def rng_words_from_device(tensor) -> np.ndarray:
    return tensor.cpu().numpy().view(np.uint64)
