"""CPU simulation of EARLY REJECTION for the Metropolis step (not shipped): a move is rejected iff
dU >= -ln(r)/beta with r the next draw, and dU has a lower bound that grows while the sweep runs (swept columns +
E[ms][t] (1 + |tv|/d_t)^-mu for untouched subunits - old row sum + bond terms).  Prints the fraction of tile calls a
chain would still execute (tile = 128 columns per warp, warps in lock step).

    python profiles/early_reject_sim.py
"""
import sys, numpy as np
sys.path.insert(0, ".")
from mchammer_b200 import synthetic
from scipy.spatial.distance import cdist

def simulate(workload, steps, seed, nwarps, beta=2.0, step_size=0.25, target=1.2, eps_b=50, eps_nb=20, sigma=1.2, mu=3, handicap=1, order="natural"):
    mol, bonds, subunits = synthetic.cube_cage() if workload == "cube" else synthetic.cuboctahedron_cage()
    pos = mol.get_position_matrix().copy()
    keys = list(subunits); members = [np.array(list(subunits[k])) for k in keys]
    S = len(keys); N = pos.shape[0]
    # slot order: subunits contiguous in dict order
    perm = np.concatenate(members); slot_of = np.empty(N, int); slot_of[perm] = np.arange(N)
    suoff = np.concatenate([[0], np.cumsum([len(m) for m in members])])
    owner = {}
    for s, m in enumerate(members):
        for a in m: owner.setdefault(int(a), s)
    scale = eps_nb * sigma ** mu
    def col_sums(ms, P):  # per-column sums over the rows of ms (slot order, zeros inside ms)
        out = np.zeros(N)
        for t in range(S):
            if t == ms: continue
            out[suoff[t]:suoff[t+1]] = (cdist(pos[members[t]], P[members[ms]]) ** -mu).sum(axis=1)
        return out
    def bondU(P):
        return sum(eps_b * (np.linalg.norm(P[a] - P[b]) - target) ** 2 for a, b in bonds)
    E = np.zeros((S, S))
    for s in range(S):
        cs = col_sums(s, pos)
        for t in range(S): E[s, t] = scale * cs[suoff[t]:suoff[t+1]].sum()
    radii = np.array([np.linalg.norm(pos[m] - pos[m].mean(0), axis=1).max() for m in members])
    rng = np.random.default_rng(seed)
    Ub = bondU(pos)
    fracs = []; acc = 0
    for step in range(steps):
        kb = int(rng.choice(len(bonds))); b = bonds[kb]
        side = int(rng.choice(2)); ms = owner[int(b[side])]
        rand = (rng.random() - 0.5) * 2
        bv = pos[b[1]] - pos[b[0]]
        cv = pos[members[ms]].mean(0) - pos.mean(0)
        kind = int(rng.choice(2))
        tv = (-bv if kind == 0 else cv) * step_size * rand
        newP = pos.copy(); newP[members[ms]] -= tv
        cs = col_sums(ms, newP)
        Enew = np.array([scale * cs[suoff[t]:suoff[t+1]].sum() for t in range(S)])
        dU = (Enew.sum() - E[ms].sum()) + (bondU(newP) - Ub)
        if dU < 0: ok = True; r = None
        else: r = rng.random(); ok = np.exp(-beta * dU) > r
        if ok:
            acc += 1; fracs.append(1.0)
            pos = newP; Ub = bondU(pos); E[ms] = Enew; E[:, ms] = Enew
            continue
        thr = -np.log(r) / beta
        # columns in column-index space (gap removed)
        g0, g1 = suoff[ms], suoff[ms+1]
        colslots = np.concatenate([np.arange(0, g0), np.arange(g1, N)])
        ncols = len(colslots); P = (ncols + 63) // 64
        # pairs -> warps (handicap for warp 0), tiles of 2 pairs (+ one single)
        h = handicap if nwarps > 1 else 0
        p0 = max(0, -(-(P + h) // nwarps) - h) if nwarps > 1 else P
        ranges = [(0, p0)]
        rest = P - p0
        if nwarps > 1:
            q, rr = divmod(rest, nwarps - 1); st = p0
            for w in range(nwarps - 1):
                n = q + (1 if w < rr else 0); ranges.append((st, st + n)); st += n
        tiles = []  # per warp: list of (col_lo, col_hi)
        for (a, bnd) in ranges:
            tl = []; p = a
            while p + 2 <= bnd: tl.append((p * 64, min((p + 2) * 64, ncols))); p += 2
            if p < bnd: tl.append((p * 64, min((p + 1) * 64, ncols)))
            tiles.append(tl)
        tvn = np.linalg.norm(tv); cms = pos[members[ms]].mean(0); cnew = newP[members[ms]].mean(0)
        cents = np.array([pos[m].mean(0) for m in members])
        dlow = np.maximum(np.linalg.norm(cents - cms, axis=1) - radii[ms] - radii, 1e-3)
        G = E[ms] * (1.0 + tvn / dlow) ** -mu; G[ms] = 0.0
        sub_of_slot = np.repeat(np.arange(S), np.diff(suoff))
        if order == "near":
            for tl in tiles:
                tl.sort(key=lambda ab: min(np.linalg.norm(cents[sub_of_slot[colslots[c]]] - cnew) for c in (ab[0], ab[1] - 1)))
        swept = np.zeros(N, bool)
        dUb = bondU(newP) - Ub
        maxk = max(len(tl) for tl in tiles); done_tiles = 0; total_tiles = sum(len(tl) for tl in tiles); exited = False
        def bound():
            Ssw = scale * cs[swept].sum()
            un = 0.0
            for t in range(S):
                if t == ms: continue
                if not swept[suoff[t]:suoff[t+1]].any(): un += G[t]
            return Ssw + un - E[ms].sum() + dUb
        if bound() >= thr: fracs.append(0.0); continue
        for k in range(maxk):
            for tl in tiles:
                if k < len(tl):
                    lo, hi = tl[k]; swept[colslots[lo:hi]] = True; done_tiles += 1
            if bound() >= thr: exited = True; break
        fracs.append(done_tiles / total_tiles)
    return acc, fracs

for wl, steps, nw in (("cube", 300, 2), ("cube", 300, 4), ("m12l24", 60, 16)):
    for order in ("natural", "near"):
        acc, fr = simulate(wl, steps, 1000, nw, order=order)
        rej = [f for f in fr]
        print(wl, "warps", nw, order, "accepted %d/%d" % (acc, steps), "mean fraction of tile calls executed %.3f" % np.mean(fr),
              "| rejected only %.3f" % np.mean([f for f, in zip(fr) if f < 1.0] or [1.0]), "| exit before any tile: %d" % sum(1 for f in fr if f == 0.0))
