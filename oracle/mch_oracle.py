"""CPU ORACLE for the MCHammer Monte-Carlo step loop.  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference algorithm for the one hot path this
repository accelerates.  It is imported only by ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- never by the
product package ``mchammer_b200`` (which fails loudly without its CUDA library).

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the unmodified reference
(imported from /root/reference/src in the build container) and commits its outputs as
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function below
against those vectors and against the literal known answers of the reference's own
tests (reference tests/conftest.py:127-152, :216-299; examples/benzene_opt.out:26-125).

Third-party arithmetic on the reference path (not under /root/reference; versions are
unpinned in its pyproject.toml:10-15; this image has numpy 2.3.5 / scipy 1.18.1):
``scipy.spatial.distance.pdist`` (optimizer.py:119), ``scipy.spatial.distance.euclidean``
(utilities.py:19-21), numpy ufuncs / ``np.sum`` / ``np.exp`` and
``numpy.random.Generator(PCG64)`` (optimizer.py:90-93).  The restatement calls the same
numpy/scipy entry points in the same order so that it is bit-identical to the
reference wherever the reference itself is reproducible.

Every function cites the reference lines it follows (paths relative to
/root/reference/src/mchammer/_internal/).
"""

from __future__ import annotations

from dataclasses import dataclass, field
from itertools import combinations

import numpy as np
from scipy.spatial.distance import euclidean, pdist


# =============================================================================
# Parameters
# =============================================================================
@dataclass
class OptimizerParams:
    """Constructor arguments of the reference Optimizer (optimizer.py:27-38)."""

    step_size: float
    target_bond_length: float
    num_steps: int
    bond_epsilon: float = 50
    nonbond_epsilon: float = 20
    nonbond_sigma: float = 1.2
    nonbond_mu: float = 3
    beta: float = 2


# =============================================================================
# Potentials
# =============================================================================
def bond_potential(distance, p: OptimizerParams):
    """optimizer.py:95-102 -- eps_b * (d - R_t)**2."""
    return p.bond_epsilon * (distance - p.target_bond_length) ** 2


def nonbond_potential(distance, p: OptimizerParams):
    """optimizer.py:104-112 -- eps_nb * (sigma / d)**mu (scalar or array)."""
    return p.nonbond_epsilon * ((p.nonbond_sigma / distance) ** p.nonbond_mu)


def nonbonded_potential(pos: np.ndarray, p: OptimizerParams) -> float:
    """optimizer.py:114-120 -- sum over ALL atom pairs (intra-subunit pairs included)."""
    return np.sum(nonbond_potential(pdist(pos), p))


def atom_distance(pos: np.ndarray, a: int, b: int) -> float:
    """utilities.py:13-21."""
    return float(euclidean(u=pos[a], v=pos[b]))


def compute_potential(pos: np.ndarray, bonds, p: OptimizerParams) -> tuple[float, float]:
    """optimizer.py:122-142 -- returns (system_potential, nonbonded_potential).

    Bond terms are added one by one, in caller order, onto the nonbonded sum.
    """
    u_nb = nonbonded_potential(pos, p)
    u = u_nb
    for b in bonds:
        u += bond_potential(atom_distance(pos, b[0], b[1]), p)
    return u, u_nb


def max_bond_distance(pos: np.ndarray, bonds) -> float:
    """optimizer.py:174-183 / :280-289."""
    return max(atom_distance(pos, b[0], b[1]) for b in bonds)


# =============================================================================
# One Metropolis run
# =============================================================================
@dataclass
class MCTrace:
    """Per-step record arrays; index k is the reference's step k (step 0 = evaluation only)."""

    system_potential: np.ndarray
    nonbonded_potential: np.ndarray
    max_bond_distance: np.ndarray
    bond_index: np.ndarray  # -1 at step 0
    passed: np.ndarray  # -1 at step 0, else 0/1
    moving_subunit: np.ndarray  # position of the moved subunit in dict order, -1 at step 0
    move_kind: np.ndarray  # 0 = along bond, 1 = along centroid vector, -1 at step 0
    positions: np.ndarray | None  # [steps, N, 3] if requested
    final_positions: np.ndarray = field(default=None)  # type: ignore[assignment]


def subunit_of_bond_ends(bonds, subunits: dict) -> list[tuple[int, int]]:
    """optimizer.py:220-221 -- first dict key (in dict order) whose container holds the atom.

    Returned as *positions* in dict order.  Raises StopIteration like the reference when
    an atom is in no subunit.
    """
    keys = list(subunits)
    out = []
    for b in bonds:
        s1 = next(k for k, key in enumerate(keys) if b[0] in subunits[key])
        s2 = next(k for k, key in enumerate(keys) if b[1] in subunits[key])
        out.append((s1, s2))
    return out


def optimizer_run(
    pos0: np.ndarray,
    bonds,
    subunits: dict,
    p: OptimizerParams,
    generator: np.random.Generator,
    keep_positions: bool = False,
) -> MCTrace:
    """optimizer.py:164-306 + loop :422-438 restated on bare arrays.

    ``generator`` is consumed exactly like the reference consumes its own
    (draws at optimizer.py:214, :225, :229, :242 and mc_operations.py:50).
    """
    pos = np.array(pos0, dtype=np.float64)
    n_steps = max(int(p.num_steps), 1)
    bonds_arr = [tuple(int(x) for x in b) for b in bonds]
    keys = list(subunits)
    members = [tuple(i for i in subunits[k]) for k in keys]
    ends = None

    tr = MCTrace(
        system_potential=np.zeros(n_steps),
        nonbonded_potential=np.zeros(n_steps),
        max_bond_distance=np.zeros(n_steps),
        bond_index=np.full(n_steps, -1, dtype=np.int64),
        passed=np.full(n_steps, -1, dtype=np.int64),
        moving_subunit=np.full(n_steps, -1, dtype=np.int64),
        move_kind=np.full(n_steps, -1, dtype=np.int64),
        positions=np.zeros((n_steps,) + pos.shape) if keep_positions else None,
    )

    # step 0 (optimizer.py:164-200)
    u, u_nb = compute_potential(pos, bonds_arr, p)
    tr.system_potential[0] = u
    tr.nonbonded_potential[0] = u_nb
    tr.max_bond_distance[0] = max_bond_distance(pos, bonds_arr)
    if keep_positions:
        tr.positions[0] = pos

    all_ids = list(range(pos.shape[0]))
    for step in range(1, n_steps):
        if ends is None:
            ends = subunit_of_bond_ends(bonds_arr, subunits)
        # :214  choice over the rows of the bond table == integers(0, B)
        k_b = int(generator.choice(len(bonds_arr)))
        b = bonds_arr[k_b]
        bond_vector = pos[b[1]] - pos[b[0]]  # utilities.py:24-31
        # :225  choice of the end whose subunit moves
        side = int(generator.choice(2))
        ms = ends[k_b][side]
        ids = list(members[ms])
        # :229
        rand = (generator.random() - 0.5) * 2
        bond_translation = -bond_vector * p.step_size * rand  # :233
        # :236-238  centroids the way Molecule.get_centroid forms them (molecule.py:204-207)
        pos_t = np.ascontiguousarray(pos.T)
        cent = np.divide(pos_t[:, all_ids].sum(axis=1), len(all_ids))
        su_cent = np.divide(pos_t[:, ids].sum(axis=1), len(ids))
        com_translation = (su_cent - cent) * p.step_size * rand
        # :242-244
        kind = int(generator.choice(2))
        tv = bond_translation if kind == 0 else com_translation
        # :248-252 (mc_operations.py:26: pos - vector)
        new_pos = pos.copy()
        new_pos[ids] = pos[ids] - tv
        new_u, new_u_nb = compute_potential(new_pos, bonds_arr, p)
        # mc_operations.py:39-52
        if new_u < u:
            ok = True
        else:
            ok = bool(np.exp(-p.beta * (new_u - u)) > generator.random())
        if ok:
            pos, u, u_nb = new_pos, new_u, new_u_nb
        else:
            # :273-277  inexact undo: (p - v) - (-v)
            undone = new_pos.copy()
            undone[ids] = new_pos[ids] - (-tv)
            pos = undone
        tr.system_potential[step] = u
        tr.nonbonded_potential[step] = u_nb
        tr.max_bond_distance[step] = max_bond_distance(pos, bonds_arr)
        tr.bond_index[step] = k_b
        tr.passed[step] = int(ok)
        tr.moving_subunit[step] = ms
        tr.move_kind[step] = kind
        if keep_positions:
            tr.positions[step] = pos

    tr.final_positions = pos
    return tr


# =============================================================================
# Collapser
# =============================================================================
def subunit_distances(pos: np.ndarray, is_h, subunits: dict):
    """collapser.py:56-82 -- cross-subunit distances, H atoms skipped, reference order."""
    for su1, su2 in combinations(subunits, 2):
        for a1 in subunits[su1]:
            for a2 in subunits[su2]:
                if not is_h[a1] and not is_h[a2]:
                    yield atom_distance(pos, a1, a2)


def min_subunit_distance(pos: np.ndarray, is_h, subunits: dict) -> float:
    """Vectorised ``min(subunit_distances(...))``; same pairs, same euclidean arithmetic.

    The per-pair distance is sqrt(dot(d, d)) exactly as numpy's norm forms it for
    3-vectors, and min() is order independent, so this equals the generator form.
    """
    keys = list(subunits)
    heavy = [np.array([a for a in subunits[k] if not is_h[a]], dtype=np.int64) for k in keys]
    best = np.inf
    for i, j in combinations(range(len(keys)), 2):
        if len(heavy[i]) == 0 or len(heavy[j]) == 0:
            continue
        d = pos[heavy[i]][:, None, :] - pos[heavy[j]][None, :, :]
        best = min(best, float(np.sqrt(np.einsum("ijk,ijk->ij", d, d).min())))
    return best


def subunit_vectors(pos: np.ndarray, subunits: dict, scale_steps: bool):
    """collapser.py:106-157 -- centroid vectors and their relative magnitudes."""
    pos_t = np.ascontiguousarray(pos.T)
    n = pos.shape[0]
    centroid = np.divide(pos_t[:, list(range(n))].sum(axis=1), n)
    vectors = {}
    for k in subunits:
        ids = subunits[k]
        if not isinstance(ids, (list, tuple)):
            ids = tuple(ids)
        vectors[k] = np.divide(pos_t[:, ids].sum(axis=1), len(ids)) - centroid
    if scale_steps:
        norms = {k: np.linalg.norm(vectors[k]) for k in vectors}
        biggest = max(list(norms.values()))
        scales = {k: float(norms[k] / biggest) for k in norms}
    else:
        scales = {k: 1.0 for k in vectors}
    return vectors, scales


def collapser_step(pos: np.ndarray, subunits: dict, step_size: float, scale_steps: bool):
    """collapser.py:84-104 + :159-180 -- pos[a] - step_size * v_s * scale_s."""
    vectors, scales = subunit_vectors(pos, subunits, scale_steps)
    new_pos = pos.copy()
    for k in subunits:
        for a in subunits[k]:
            new_pos[a] = pos[a] - step_size * vectors[k] * scales[k]
    return new_pos


@dataclass
class CollapseTrace:
    max_bond_distance: np.ndarray  # one entry per executed step (steps are 1-based)
    min_distance: float  # value tested against threshold/2 (collapser.py:264-266)
    steps_run: int
    final_positions: np.ndarray
    positions: np.ndarray | None


def collapser_run(
    pos0: np.ndarray,
    is_h,
    bonds,
    subunits: dict,
    step_size: float,
    distance_threshold: float,
    scale_steps: bool = True,
    keep_positions: bool = False,
    max_steps: int = 1_000_000,
) -> CollapseTrace:
    """collapser.py:252-283 / :317-339 restated on bare arrays."""
    pos = np.array(pos0, dtype=np.float64)
    bonds_arr = [tuple(int(x) for x in b) for b in bonds]
    maxd: list[float] = []
    frames: list[np.ndarray] = []
    step = 0
    while not (min_subunit_distance(pos, is_h, subunits) < distance_threshold):
        if step >= max_steps:
            raise RuntimeError("collapser oracle: max_steps reached")
        step += 1
        pos = collapser_step(pos, subunits, step_size, scale_steps)
        maxd.append(max_bond_distance(pos, bonds_arr))
        if keep_positions:
            frames.append(pos.copy())
    min_dist = min_subunit_distance(pos, is_h, subunits)
    if min_dist < distance_threshold / 2:
        step += 1
        pos = collapser_step(pos, subunits, -(step_size / 2), scale_steps)
        maxd.append(max_bond_distance(pos, bonds_arr))
        if keep_positions:
            frames.append(pos.copy())
    return CollapseTrace(
        max_bond_distance=np.array(maxd),
        min_distance=min_dist,
        steps_run=step,
        final_positions=pos,
        positions=np.array(frames) if keep_positions else None,
    )


# =============================================================================
# numpy PCG64 stream model (integer arithmetic, pure Python -- small cases only)
# =============================================================================
_M64 = (1 << 64) - 1
_M128 = (1 << 128) - 1
_M32 = 0xFFFFFFFF
PCG_MULT = 0x2360ED051FC65DA44385DF649FCCF645  # PCG_DEFAULT_MULTIPLIER_128


def seed_sequence_words(seed: int, n_words32: int) -> list[int]:
    """numpy.random.SeedSequence(seed).generate_state(n, uint32), restated.

    Published algorithm (M.E. O'Neill's seed_seq_fe as adopted by numpy's
    bit_generator.pyx): 4-word pool, hashmix/mix constants below.
    """
    XSHIFT = 16
    INIT_A, MULT_A = 0x43B0D7E5, 0x931E8875
    INIT_B, MULT_B = 0x8B51F9DD, 0x58F38DED
    MIX_L, MIX_R = 0xCA01F9DD, 0x4973F715

    if seed < 0:
        raise ValueError("seed must be non-negative")
    entropy = []
    if seed == 0:
        entropy = [0]
    while seed > 0:
        entropy.append(seed & _M32)
        seed >>= 32

    hash_const = [INIT_A]

    def hashmix(value: int) -> int:
        value = (value ^ hash_const[0]) & _M32
        hash_const[0] = (hash_const[0] * MULT_A) & _M32
        value = (value * hash_const[0]) & _M32
        value ^= value >> XSHIFT
        return value

    def mix(x: int, y: int) -> int:
        r = (MIX_L * x - MIX_R * y) & _M32
        r ^= r >> XSHIFT
        return r

    pool = [0, 0, 0, 0]
    for i in range(4):
        pool[i] = hashmix(entropy[i] if i < len(entropy) else 0)
    for i_src in range(4):
        for i_dst in range(4):
            if i_src != i_dst:
                pool[i_dst] = mix(pool[i_dst], hashmix(pool[i_src]))
    for i_src in range(4, len(entropy)):
        for i_dst in range(4):
            pool[i_dst] = mix(pool[i_dst], hashmix(entropy[i_src]))

    hc = INIT_B
    out = []
    for i_dst in range(n_words32):
        v = pool[i_dst % 4]
        v ^= hc
        hc = (hc * MULT_B) & _M32
        v = (v * hc) & _M32
        v ^= v >> XSHIFT
        out.append(v)
    return out


class PCG64Model:
    """Integer model of numpy's PCG64 bit generator and of the Generator calls the path makes.

    * ``next64``: 128-bit LCG step then XSL-RR output (numpy pcg64.h ``pcg64_next64``).
    * ``next32``: low half first, high half buffered (``pcg64_next32``).
    * ``bounded(n)``: ``Generator.integers(0, n)`` / ``choice(n)`` for n <= 2**32 --
      no draw for n == 1, otherwise Lemire's method on 32-bit halves.
    * ``random()``: ``(next64 >> 11) * 2**-53``; does not touch the buffered half.
    """

    def __init__(self, state: int, inc: int, has_uint32: int = 0, uinteger: int = 0):
        self.state = state & _M128
        self.inc = inc & _M128
        self.has_uint32 = has_uint32
        self.uinteger = uinteger

    @classmethod
    def from_seed(cls, seed: int) -> "PCG64Model":
        w = seed_sequence_words(seed, 8)
        v = [w[2 * i] | (w[2 * i + 1] << 32) for i in range(4)]  # uint64 view, little endian
        initstate = (v[0] << 64) | v[1]
        initseq = (v[2] << 64) | v[3]
        inc = ((initseq << 1) | 1) & _M128
        state = 0
        state = (state * PCG_MULT + inc) & _M128
        state = (state + initstate) & _M128
        state = (state * PCG_MULT + inc) & _M128
        return cls(state, inc)

    def next64(self) -> int:
        self.state = (self.state * PCG_MULT + self.inc) & _M128
        hi, lo = self.state >> 64, self.state & _M64
        x = hi ^ lo
        rot = hi >> 58
        return ((x >> rot) | (x << ((-rot) & 63))) & _M64

    def next32(self) -> int:
        if self.has_uint32:
            self.has_uint32 = 0
            return self.uinteger
        w = self.next64()
        self.has_uint32 = 1
        self.uinteger = w >> 32
        return w & _M32

    def bounded(self, n: int) -> int:
        if n == 1:
            return 0
        m = self.next32() * n
        leftover = m & _M32
        if leftover < n:
            threshold = ((1 << 32) - n) % n
            while leftover < threshold:
                m = self.next32() * n
                leftover = m & _M32
        return m >> 32

    def random(self) -> float:
        return (self.next64() >> 11) * (1.0 / 9007199254740992.0)
