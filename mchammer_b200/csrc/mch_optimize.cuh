// mch_optimize.cuh -- the resident Metropolis kernel: ONE CTA owns ONE chain for all steps.
//
// Replaces, per chain, the reference loop
//     Optimizer.get_result / get_trajectory   (optimizer.py:308-438)
//       _run_first_step                       (optimizer.py:164-200)
//       _run_step                             (optimizer.py:202-306)
//         translate_atoms_along_vector        (mc_operations.py:15-28)
//         compute_potential                   (optimizer.py:122-142)
//         test_move                           (mc_operations.py:39-52)
//
// Data layout
//   HBM   : positions [C][N][3] (io dtype, caller's atom order), rng_state [C][6] u64,
//           topology arrays shared by all chains (perm, su_offsets, bonds, bond_su).
//   SMEM  : the chain's atoms as SoA X/Y/Z in *slot* order (atoms permuted so that every
//           subunit is a contiguous slot range), the displaced rows Q of the moving
//           subunit, per-column sums, the S x S table E of subunit-pair energies,
//           per-subunit coordinate sums.  Nothing goes back to HBM between steps except
//           the optional trace/trajectory records.
//
// Algorithm per step (exact reference order of random draws)
//   control warp : draws -> translation vector -> encoded rows Q of P[ms] - tv    [barrier A]
//   all warps    : sum_{i in ms, j not in ms} r_ij^-mu  (pair_sweep: n_ms x (N - n_ms) terms;
//                  per-column sums to shared memory, per-warp totals)             [barrier B]
//   control warp : dU_nb = eps sigma^mu * total - sum_t E[ms][t];  Metropolis test;
//                  accept -> P[ms] -= tv;  reject -> P untouched (fp32) or the reference's
//                  inexact undo (p - v) + v (fp64 parity mode, optimizer.py:269-277);
//                  then the next proposal.
//   The half of the control work that the next sweep does not depend on -- rebuilding
//   row/column ms of E from the column sums after an accept, the bond energy of the trial
//   geometry, trace records -- is DEFERRED: warp 0 runs it after barrier A, under the other
//   warps' sweep, reading the previous sweep's column sums from the other of two buffers.
//   Intra-subunit pairs are constant under rigid moves; they are evaluated once at step 0
//   and carried in U_nb (the reference re-sums them every step, optimizer.py:114-120).
//   The control warp's registers (RNG state, potentials, move) are parked in shared memory
//   while it sweeps, so nothing is spilled around the sweep-tile calls.
#pragma once

#include "mch_device.cuh"

#ifdef __CUDACC__
#include <cooperative_groups.h>
#endif

// slot of the double-buffered cluster exchange used at `step`; the negative control of the check build
// (-DMCH_RACECHECK_BREAK=1) drops the double buffering, which the race detector must report
#if defined(MCH_RACECHECK_BREAK) && MCH_RACECHECK_BREAK == 1
#define MCH_RC_PARITY(step) 0
#else
#define MCH_RC_PARITY(step) ((step) & 1)
#endif

// pairs of column units the control warp of a two-warp fp32 chain sweeps less than its share in the EARLY
// instantiation (see the sweep call in the step loop; measured: profiles/r02_ab_early_handicap.txt)
// 1: the deferred control work is shared -- warp 1 forms the bond energy of the trial geometry under the sweep,
// warp 0 keeps the energy-table refresh and the records (see the step loop).  Measured and not taken
// (profiles/r02_ab_split_deferred.txt: fp32 6.61e7 -> 6.47e7, fp64 2.89e7 -> 2.83e7): default 0.
#ifndef MCH_SPLIT_D
#define MCH_SPLIT_D 0
#endif
// 1: the trial geometry's bond energy is summed through shared memory (one store, B broadcast loads) instead of five
// dependent shuffle rounds.  Measured and not taken (profiles/r02_ab_bond_smem.txt: fp32 6.60e7 -> 6.55e7): default 0.
#ifndef MCH_LAZY_CTOT
#define MCH_LAZY_CTOT 1
#endif
#ifndef MCH_BOND_SMEM
#define MCH_BOND_SMEM 0
#endif
#ifndef MCH_EARLY_HANDICAP
#define MCH_EARLY_HANDICAP 1
#endif

namespace mch {

struct OptArgs {
  const void* pos_in;        // [C or 1][N][3] io dtype
  long long pos_in_stride;   // elements between chains (0: every chain starts from chain 0)
  void* pos_out;             // [C][N][3] io dtype
  int C, N, S, B;
  int max_su;                // largest subunit (atoms)
  const int* perm;           // [N] slot -> atom id
  const int* su_off;         // [S+1] slot offsets
  const int* bonds;          // [B][2] atom ids, caller's order inside each pair
  const int* bond_su;        // [B][2] subunit (position in dict order) of each bond end
  Params p;
  int mu_int;
  int num_steps;             // counts step 0, like the reference
  uint64_t* rng;             // [C][6] in/out
  double* U;                 // [C] out
  double* Unb;               // [C] out
  double* maxd;              // [C] out
  int* n_accept;             // [C] out
  int* last_info;            // [C] out: trace word of the last step
  double* trace_f;           // [C][num_steps][3] or null: U, U_nb, max bond distance
  int* trace_i;              // [C][num_steps] or null: bond | passed<<16 | kind<<17 | ms<<18
  void* traj;                // [C][num_steps][N][3] io dtype or null
  unsigned long long* pairs; // [C] or null: pair terms evaluated
  int inexact_undo;          // 1: reproduce (p - v) + v on reject (fp64 parity)
  int defer;                 // 1: deferred control work + second column-sum buffer (see header)
  int cluster;               // CTAs per chain (thread-block cluster; 1 = the chain is one CTA)
  int regs64;                // 1: the host wants the 64-register (1024-thread) fp32 instantiation: two 512-thread chains per SM
  int early;                 // 1: the EARLY instantiation (proposals that a pair-free bound rejects never wake the sweep)
};

struct Ctl {
  int nrows;
  int c0, g0, g1, c1;
  int ms;
  int buf;           // which column-sum buffer this sweep writes
  double cx, cy, cz; // centre of the expanded row encoding (fp32 mode), else unused
  double origin[3];  // frame shift of the fp32 mode (start centroid), zero in fp64 mode
  // Control-warp state parked here while the warp sweeps (the sweep tiles are real function
  // calls; anything kept in registers across them would be spilled to local memory, which
  // at 208 KB of shared memory per SM means L2 latency on every reload).
  unsigned long long rng_hi, rng_lo, rng_inc_hi, rng_inc_lo, pair_count;
  double U, Unb, Ub_trial;
  double tv[3];
  unsigned rng_has32, rng_half;
  int kb, kind, n_acc, info, d_accepted, d_ms, d_buf, d_step, cur_ms;
  int next_step;             // EARLY instantiation: the step whose sweep follows the next barrier A
  double ub_w1;              // bond energy of the trial geometry, formed by warp 1 under the sweep (kSplitD)
};

// Shared-memory carve-up, identical on host (size query) and device.
struct OptLayout {
  int x, y, z, colsum, colsum_stride, rows, E, erow, csum, invn, wsum, xsum, epart, suoff, bslot, bsu, ctl, rc, rad, total;
};

MCH_HD int align_up(int v, int a) { return (v + a - 1) / a * a; }

constexpr int kMaxCluster = 16;  // 16 = the non-portable cluster size (opted in at launch)

MCH_HD OptLayout opt_layout(int N, int S, int B, int max_su, int tsize, int defer, int cluster = 1) {
  OptLayout L;
  const int np = align_up(N, 8);
  int o = 0;
  L.x = o; o += np * tsize;
  L.y = o; o += np * tsize;
  L.z = o; o += np * tsize;
  L.colsum = o; L.colsum_stride = np * (tsize < 4 ? 4 : tsize);
  // double buffered when control work is deferred -- except on a cluster: there the column sums of an accepted
  // sweep are consumed by all warps (coop_row) before the next sweep starts, so one buffer serves
  o += (defer && cluster <= 1 ? 2 : 1) * L.colsum_stride;
  o = align_up(o, 32);
  L.rows = o; o += (max_su + kRowSlack) * row_bytes(tsize);  // + far-away padding row + prefetch overrun
  o = align_up(o, 16);
  L.E = o; o += S * S * 8;
  L.erow = o; o += S * 8;
  L.csum = o; o += (S + 1) * 3 * 8;  // per-subunit coordinate sums + the molecule's
  L.invn = o; o += S * 8;
  L.wsum = o; o += 3 * 32 * 8;
  // cluster mode: per-CTA sweep totals and per-CTA partial rows of E, each double buffered by step parity
  L.xsum = o; o += cluster > 1 ? 2 * kMaxCluster * 8 : 0;
  L.epart = o; o += cluster > 1 ? 2 * kMaxCluster * S * 8 : 0;
  L.suoff = o; o += align_up(S + 1, 4) * 4;
  L.bslot = o; o += align_up(2 * B, 4) * 4;
  L.bsu = o; o += align_up(2 * B, 4) * 4;
  o = align_up(o, 16);
  L.ctl = o; o += (int)sizeof(Ctl);
  L.rc = o;
#ifdef MCH_RACECHECK  // shadows of the locations other CTAs of the cluster touch: xsum, epart, E (8 bytes each)
  o += cluster > 1 ? (2 * kMaxCluster + 2 * kMaxCluster * S + S * S) * 8 : 0;
#endif
  o = align_up(o, 8);
  L.rad = o; o += S * 4;  // subunit radii, fp32 (EARLY instantiation)
  L.total = align_up(o, 16);
  return L;
}

#ifdef __CUDACC__

template <typename Tio> MCH_D double load_io(const void* base, long long idx) {
  return (double)reinterpret_cast<const Tio*>(base)[idx];
}
template <typename Tio> MCH_D void store_io(void* base, long long idx, double v) {
  reinterpret_cast<Tio*>(base)[idx] = (Tio)v;
}

// A chain's start geometry is one contiguous run of 3 N values in HBM: it is staged with 16-byte
// loads (double2 / float4), consecutive lanes taking consecutive vectors -- fully coalesced -- and
// every value is scattered to its SoA slot through `inv` (atom id -> slot, in shared memory).  A
// chain whose run does not start on a 16-byte boundary (odd N and odd chain index) falls back to
// element loads.  fp32 compute works in a frame centred on the start geometry (translation
// invariant path; keeps |coordinates| small): origin = centroid; fp64 keeps the caller's frame bit
// for bit (origin = 0).  A real function, so that the step loop's code does not depend on it.
template <typename T, typename Tio>
__device__ __noinline__ void stage_positions(const Tio* __restrict__ chain_in, int N, const int* __restrict__ inv,
                                             T* __restrict__ X, T* __restrict__ Y, T* __restrict__ Z,
                                             double* __restrict__ wsum, double* __restrict__ origin) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt = blockDim.x, nwarps = nt >> 5;
  constexpr int VEC = 16 / (int)sizeof(Tio);
  const int nvec = (reinterpret_cast<unsigned long long>(chain_in) & 15ULL) == 0ULL ? (3 * N) / VEC : 0;
  auto for_each_value = [&](auto&& visit) {  // visit(element index in [0, 3N), value)
    for (int v = tid; v < nvec; v += nt) {
      if constexpr (VEC == 2) {
        const double2 d = reinterpret_cast<const double2*>(chain_in)[v];
        visit(2 * v, (double)d.x); visit(2 * v + 1, (double)d.y);
      } else {
        const float4 d = reinterpret_cast<const float4*>(chain_in)[v];
        visit(4 * v, (double)d.x); visit(4 * v + 1, (double)d.y); visit(4 * v + 2, (double)d.z); visit(4 * v + 3, (double)d.w);
      }
    }
    for (int e = nvec * VEC + tid; e < 3 * N; e += nt) visit(e, (double)chain_in[e]);
  };
  double ox = 0.0, oy = 0.0, oz = 0.0;
  if (sizeof(T) == 4) {
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for_each_value([&](int e, double v) {
      const int comp = e - 3 * (e / 3);
      sx += comp == 0 ? v : 0.0; sy += comp == 1 ? v : 0.0; sz += comp == 2 ? v : 0.0;
    });
    sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
    __syncthreads();
    if (lane == 0) { wsum[warp] = sx; wsum[32 + warp] = sy; wsum[64 + warp] = sz; }
    __syncthreads();
    for (int w = 0; w < nwarps; ++w) { ox += wsum[w]; oy += wsum[32 + w]; oz += wsum[64 + w]; }
    ox /= N; oy /= N; oz /= N;
  }
  __syncthreads();
  for_each_value([&](int e, double v) {
    const int atom = e / 3, comp = e - 3 * atom, slot = inv[atom];
    if (comp == 0) X[slot] = (T)(v - ox);
    else if (comp == 1) Y[slot] = (T)(v - oy);
    else Z[slot] = (T)(v - oz);
  });
  origin[0] = ox; origin[1] = oy; origin[2] = oz;
}

// Subunits up to this many atoms are re-summed one lane per subunit when the energy table is
// refreshed; larger ones by the whole warp.
constexpr int REGROUP_SMALL = 64;

// NTMAX: largest block an instantiation may be launched with; together with the minimum number
// of resident blocks it fixes the register budget.  Every instantiation but the 1024-thread one
// gets 128 registers per thread (no spills in the control section, rows fetched a whole trip
// ahead in the packed tile):
//   32 threads (10 chains/SM), 64 (8 chains/SM -- the S1 default), 128 (4), 256 (2), 512 (1; also the
//   cluster-capable instantiation).  1024 threads at 64 registers (fp32 only) serves molecules of
//   which two 512-thread chains share an SM (M12L24) and chains that want more than 512 threads.
template <typename T> constexpr int large_block() { return sizeof(T) == 4 ? 1024 : 512; }
template <typename T, int NTMAX> constexpr int min_blocks_per_sm() {
  return NTMAX == 32 ? 10 : NTMAX == 64 ? 8 : NTMAX == 128 ? 4 : NTMAX == 256 ? 2 : 1;
}

// CL: the instantiation can run a chain on a thread-block cluster of a.cluster CTAs (launched with
// that cluster dimension).  Every CTA keeps the whole chain and runs the control section redundantly
// -- same inputs, same arithmetic, same decisions -- and only the sweep's columns are split; what the
// CTAs exchange through distributed shared memory, one cluster barrier per step, is the per-CTA sweep
// total and, after an accept, the per-CTA partial sums of the new energy-table row.
// EARLY: proposals that a pair-free lower bound of U_new - U rejects are decided by the control warp on the spot --
// no barrier, no sweep -- and the chain moves on to its next proposal (see propose_next below).  A separate
// instantiation, so that the default kernel's code is exactly what it was.
template <typename T, typename Tio, int MUK, int NTMAX, bool CL = false, bool EARLY = false>
__global__ void __launch_bounds__(NTMAX, (min_blocks_per_sm<T, NTMAX>()))
mch_optimize_kernel(const __grid_constant__ OptArgs a) {
  static_assert(!(EARLY && CL), "early rejection: one-CTA chains only");
  extern __shared__ __align__(32) unsigned char smem[];
  const int G = CL ? a.cluster : 1;                  // CTAs of this chain
  const int c = CL ? (int)blockIdx.x / G : (int)blockIdx.x;
  const int rank = CL ? (int)blockIdx.x % G : 0;     // == %cluster_ctarank for a (G,1,1) cluster
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nt = blockDim.x, nwarps = nt >> 5;
  const int N = a.N, S = a.S, B = a.B;

  const OptLayout L = opt_layout(N, S, B, a.max_su, (int)sizeof(T), a.defer, G);
  T* X = reinterpret_cast<T*>(smem + L.x);
  T* Y = reinterpret_cast<T*>(smem + L.y);
  T* Z = reinterpret_cast<T*>(smem + L.z);
  T* colsum = reinterpret_cast<T*>(smem + L.colsum);  // buffer 0; buffer 1 follows
  const int colsum_stride = L.colsum_stride / (int)sizeof(T);
  Row4<T>* Q = reinterpret_cast<Row4<T>*>(smem + L.rows);
  double* E = reinterpret_cast<double*>(smem + L.E);
  double* erow = reinterpret_cast<double*>(smem + L.erow);  // erow[s] = sum_t E[s][t]
  double* csum = reinterpret_cast<double*>(smem + L.csum);
  double* ctot = csum + 3 * S;  // coordinate sums of the whole molecule
  double* invn = reinterpret_cast<double*>(smem + L.invn);
  double* wsum = reinterpret_cast<double*>(smem + L.wsum);   // 3 x 32: [0..31] cross, [32..63] intra
  double* xsum = reinterpret_cast<double*>(smem + L.xsum);   // cluster: [parity][rank] sweep totals
  double* epart = reinterpret_cast<double*>(smem + L.epart); // cluster: [parity][rank][S] partial rows of E
  int* suoff = reinterpret_cast<int*>(smem + L.suoff);
  int* bslot = reinterpret_cast<int*>(smem + L.bslot);
  int* bsu = reinterpret_cast<int*>(smem + L.bsu);
  Ctl* ctl = reinterpret_cast<Ctl*>(smem + L.ctl);
#ifdef MCH_RACECHECK
  RcShadow* rc_xsum = reinterpret_cast<RcShadow*>(smem + L.rc);   // [2][kMaxCluster]
  RcShadow* rc_epart = rc_xsum + 2 * kMaxCluster;                 // [2][kMaxCluster][S]
  RcShadow* rc_E = rc_epart + 2 * kMaxCluster * S;                // [S][S]
  unsigned rc_epoch = 1u;                                         // cluster barriers passed + 1
  if (CL && G > 1)
    for (int k = tid; k < 2 * kMaxCluster + 2 * kMaxCluster * S + S * S; k += nt) { rc_xsum[k].w = 0u; rc_xsum[k].r = 0u; }
#endif

  const PairTerm<T, MUK> term(a.p.nonbond_mu, a.mu_int);
  const double scale = a.p.nonbond_epsilon * pow(a.p.nonbond_sigma, a.p.nonbond_mu);

  // ---------------------------------------------------------------- prologue
  // inverse permutation (atom id -> slot) in the colsum scratch, then bonds in slot space
  int* inv = reinterpret_cast<int*>(colsum);
  for (int n = tid; n < N; n += nt) inv[a.perm[n]] = n;
  for (int s = tid; s <= S; s += nt) suoff[s] = a.su_off[s];
  __syncthreads();
  for (int k = tid; k < 2 * B; k += nt) {
    bslot[k] = inv[a.bonds[k]];
    bsu[k] = a.bond_su[k];
  }
  const long long in_base = (long long)c * a.pos_in_stride;
  // fp32 compute works in a frame centred on the start geometry (translation invariant
  // path; keeps |coordinates| small).  fp64 keeps the caller's frame bit for bit.
  double origin[3] = {0.0, 0.0, 0.0};  // prologue only; kept in ctl->origin afterwards
  stage_positions<T, Tio>(reinterpret_cast<const Tio*>(a.pos_in) + in_base, N, inv, X, Y, Z, wsum, origin);
  if (tid == 0) { ctl->origin[0] = origin[0]; ctl->origin[1] = origin[1]; ctl->origin[2] = origin[2]; }
  for (int k = tid; k < S * S; k += nt) E[k] = 0.0;
  for (int k = tid; k < S; k += nt) invn[k] = 1.0 / (double)(suoff[k + 1] - suoff[k]);
  __syncthreads();
  if (CL && G > 1) { cooperative_groups::this_cluster().sync(); MCH_RC(++rc_epoch); }  // every CTA's E is zeroed before any peer writes rows into it

  // ---------------------------------------------------------------- control-warp state
  // (uniform across the 32 lanes of warp 0; other warps never read it)
  constexpr bool kFast = sizeof(T) == 4;  // fp32 mode: cheaper centroid bookkeeping
  // expanded-form squared distances (mch_device.cuh): 4 instead of 6 FP instructions per pair.  Relative
  // rounding noise in r^2 ~ eps (|p_i'| + |p_j'|)^2 / r^2 with coordinates taken from the moving subunit's
  // centroid: ~1e-6 in fp32 (inside the 1e-4 tolerance), ~1e-14 in fp64 (the parity tolerance is 1e-9;
  // MCH_XF64 = 0 restores the plain form there)
  constexpr bool kXF = kFast ? (MCH_XF != 0) : (MCH_XF64 != 0);
  constexpr bool kX2 = kFast && kXF && (MCH_F32X2 != 0);  // packed-instruction tile
  constexpr int kRot3 = !(kX2 && (MCH_X2_ROT3 != 0)) ? 0 : (NTMAX <= 512 ? 2 : 1);  // 2: deep row prefetch (128 registers)
  const bool kDefer = a.defer != 0;  // host policy: off when the second buffer would cost a resident chain
  const bool kTwoBuf = kDefer && !(CL && G > 1);  // two column-sum buffers (see opt_layout)
  // re-sum an accepted move's row of E with all warps (coop_row) where warp 0 alone would take long
  const bool kCoop = nwarps > 1 && ((CL && G > 1) || !kDefer || a.max_su > REGROUP_SMALL);
  // parity mode: the molecule's coordinate sums are re-formed lazily (draw_move), not after every change
  constexpr int kLazyCtot = kFast ? 0 : MCH_LAZY_CTOT;
  // the deferred work is shared between warps 0 and 1 (one-CTA chains of at least two warps)
  const bool kSplitD = (MCH_SPLIT_D != 0) && kDefer && nwarps > 1 && !(CL && G > 1);
  Pcg64 rng;  // loaded after step 0 (nothing of the control state is live across the step-0 sweeps)
  rng.hi = rng.lo = rng.inc_hi = rng.inc_lo = 0ULL; rng.has32 = rng.half = 0u;
  double U = 0.0, Unb = 0.0, Ub_trial = 0.0;
  T tvx = (T)0, tvy = (T)0, tvz = (T)0;
  int ms = 0, kb = 0, kind = 0, n_acc = 0, info = -1;
  unsigned long long pair_count = 0ULL;

  bool d_accepted = false;  // the last decided move was accepted: row/column d_ms of E is stale
  int d_ms = 0, d_buf = 0, d_step = 0;
  auto park = [&]() {  // registers -> shared memory (lane 0 writes; every lane holds the same values)
    if (lane == 0) {
      ctl->rng_hi = rng.hi; ctl->rng_lo = rng.lo; ctl->rng_inc_hi = rng.inc_hi; ctl->rng_inc_lo = rng.inc_lo;
      ctl->rng_has32 = rng.has32; ctl->rng_half = rng.half;
      ctl->U = U; ctl->Unb = Unb; ctl->Ub_trial = Ub_trial; ctl->pair_count = pair_count;
      ctl->tv[0] = (double)tvx; ctl->tv[1] = (double)tvy; ctl->tv[2] = (double)tvz;
      ctl->kb = kb; ctl->kind = kind; ctl->n_acc = n_acc; ctl->info = info; ctl->cur_ms = ms;
      ctl->d_accepted = d_accepted ? 1 : 0; ctl->d_ms = d_ms; ctl->d_buf = d_buf; ctl->d_step = d_step;
    }
    __syncwarp();
  };
  auto unpark = [&]() {  // shared memory -> registers of all 32 lanes
    rng.hi = ctl->rng_hi; rng.lo = ctl->rng_lo; rng.inc_hi = ctl->rng_inc_hi; rng.inc_lo = ctl->rng_inc_lo;
    rng.has32 = ctl->rng_has32; rng.half = ctl->rng_half;
    U = ctl->U; Unb = ctl->Unb; Ub_trial = ctl->Ub_trial; pair_count = ctl->pair_count;
    tvx = (T)ctl->tv[0]; tvy = (T)ctl->tv[1]; tvz = (T)ctl->tv[2];
    kb = ctl->kb; kind = ctl->kind; n_acc = ctl->n_acc; info = ctl->info; ms = ctl->cur_ms;
    d_accepted = ctl->d_accepted != 0; d_ms = ctl->d_ms; d_buf = ctl->d_buf; d_step = ctl->d_step;
  };
  // Q = encoded rows of P[s] - d, padded with vanishing rows to an even count (see sweep_tile);
  // warp 0 only.  Expanded encoding: relative to the centroid of the displaced subunit, which
  // is also published for the column side.  Returns the padded count.
  auto fill_rows = [&](int s, T dx, T dy, T dz) {
    const int r0 = suoff[s], n = suoff[s + 1] - r0, np = kRot3 ? padded_rows3(n) : padded_rows(n);
    Centre<T> ctr{(T)0, (T)0, (T)0};
    if (kXF) {
      ctr.x = (T)(csum[3 * s] * invn[s]) - dx;
      ctr.y = (T)(csum[3 * s + 1] * invn[s]) - dy;
      ctr.z = (T)(csum[3 * s + 2] * invn[s]) - dz;
      if (lane == 0) { ctl->cx = (double)ctr.x; ctl->cy = (double)ctr.y; ctl->cz = (double)ctr.z; }
    }
    for (int i = lane; i < np; i += 32) {
      Q[i] = i < n ? encode_row<T, kXF>(X[r0 + i] - dx, Y[r0 + i] - dy, Z[r0 + i] - dz, ctr)
                   : padding_row<T, kXF>();
    }
    return np;
  };
  auto refresh_csum = [&](int s) {  // coordinate sums of subunit s from current positions
    const int r0 = suoff[s], r1 = suoff[s + 1];
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int i = r0 + lane; i < r1; i += 32) { sx += (double)X[i]; sy += (double)Y[i]; sz += (double)Z[i]; }
    warp_sum3(sx, sy, sz, lane);  // totals in lanes 0 / 16 / 8, the bits of three warp_sums
    if (lane == 0) csum[3 * s] = sx;
    if (lane == 16) csum[3 * s + 1] = sy;
    if (lane == 8) csum[3 * s + 2] = sz;
  };
  auto refresh_ctot = [&]() {  // molecule sums = sum of the subunit sums (fixed order)
    double tx = 0.0, ty = 0.0, tz = 0.0;
    for (int s = lane; s < S; s += 32) { tx += csum[3 * s]; ty += csum[3 * s + 1]; tz += csum[3 * s + 2]; }
    warp_sum3(tx, ty, tz, lane);
    if (lane == 0) ctot[0] = tx;
    if (lane == 16) ctot[1] = ty;
    if (lane == 8) ctot[2] = tz;
  };
  // bond energy sum and longest bond; ends inside subunit `moved` are displaced by -tv
  // (a warp reduction whose result nobody reads is not dead code to the compiler: say which ones are wanted)
  auto bond_terms = [&](int moved, T dx, T dy, T dz, double& longest, bool want_sum = true, bool want_longest = true) {
    const int m0 = moved >= 0 ? suoff[moved] : -1, m1 = moved >= 0 ? suoff[moved + 1] : -1;
    double e = 0.0, far = 0.0;
    for (int b = lane; b < B; b += 32) {
      const int sa = bslot[2 * b], sb = bslot[2 * b + 1];
      const bool ma = sa >= m0 && sa < m1, mb = sb >= m0 && sb < m1;
      const T ax = ma ? X[sa] - dx : X[sa], ay = ma ? Y[sa] - dy : Y[sa], az = ma ? Z[sa] - dz : Z[sa];
      const T bx = mb ? X[sb] - dx : X[sb], by = mb ? Y[sb] - dy : Y[sb], bz = mb ? Z[sb] - dz : Z[sb];
      const double ex = (double)ax - (double)bx, ey = (double)ay - (double)by, ez = (double)az - (double)bz;
      const double r = sqrt(ex * ex + ey * ey + ez * ez);
      const double d = r - a.p.target_bond_length;
      e += a.p.bond_epsilon * (d * d);
      far = fmax(far, r);
    }
    if (want_longest) longest = warp_max(far);
#if MCH_BOND_SMEM
    // the sum alone, at most one bond per lane: through shared memory (one store, B broadcast loads that pipeline)
    // instead of five dependent shuffle rounds -- fixed order b = 0 .. B - 1
    if (want_sum && !want_longest && B <= 32) {
      wsum[64 + lane] = e;
      __syncwarp();
      double s0 = 0.0, s1 = 0.0;
      for (int b = 0; b < B; b += 2) { s0 += wsum[64 + b]; s1 += wsum[64 + b + 1]; }  // lanes >= B hold 0
      __syncwarp();
      return s0 + s1;
    }
#endif
    return want_sum ? warp_sum(e) : 0.0;
  };
  auto write_trace = [&](int step, double longest) {
    if (lane == 0 && rank == 0) {
      if (a.trace_f) {
        double* t = a.trace_f + ((long long)c * a.num_steps + step) * 3;
        t[0] = U; t[1] = Unb; t[2] = longest;
      }
      if (a.trace_i) a.trace_i[(long long)c * a.num_steps + step] = info;
    }
  };
  // Row/column `row` of the energy table from the per-column sums `cs` of a finished sweep:
  // E[row][t] = E[t][row] = scale * sum_{j in t} cs[j] for t in [t_first, S), t != row.
  // Small subunits: one lane sums one subunit (32 subunits per pass, no shuffles);
  // large subunits: the warp strides over the subunit and butterfly-reduces.
  auto regroup = [&](const T* cs, int row, int t_first) {
    for (int t0 = t_first; t0 < S; t0 += 32) {
      const int t = t0 + lane;
      const bool mine = t < S && t != row;
      const int j0 = mine ? suoff[t] : 0;
      const int len = mine ? suoff[t + 1] - j0 : 0;
      const bool small = len <= REGROUP_SMALL;
      const int span = small ? len : 0;
      const int longest_span = __reduce_max_sync(0xffffffffu, span);
      double v0 = 0.0, v1 = 0.0;
      for (int k = 0; k < longest_span; k += 2) {
        if (k < span) v0 += (double)cs[j0 + k];
        if (k + 1 < span) v1 += (double)cs[j0 + k + 1];
      }
      if (mine && small) { const double v = (v0 + v1) * scale; E[row * S + t] = v; E[t * S + row] = v; }
      unsigned big = __ballot_sync(0xffffffffu, mine && !small);
      while (big) {
        const int tb = t0 + (__ffs(big) - 1);
        big &= big - 1;
        double v = 0.0;
        for (int j = suoff[tb] + lane; j < suoff[tb + 1]; j += 32) v += (double)cs[j];
        v = warp_sum(v) * scale;
        if (lane == 0) { E[row * S + tb] = v; E[tb * S + row] = v; }
      }
    }
  };
  auto refresh_erow = [&]() {  // erow[s] = sum_t E[s][t], one lane per row, fixed order
    __syncwarp();
    for (int s = lane; s < S; s += 32) {
      double v = 0.0;
      for (int t = 0; t < S; ++t) v += (t == s) ? 0.0 : E[s * S + t];
      erow[s] = v;
    }
    __syncwarp();
  };
  // The new row `row` of E after an accepted move, re-summed from the column sums `cs` of the accepted
  // sweep by ALL warps of the CTA (subunit t goes to warp t mod nwarps) -- used where one warp would
  // take long over it: large subunits, chains that do not defer, chains on a cluster.
  //   one CTA : E[row][t] = E[t][row] = scale * sum_{j in t} cs[j]
  //   cluster : this CTA swept only the column range [o_lo, o_hi) of the sweep that excluded `row`;
  //             its share of every sum goes into slot [parity][rank] of EVERY CTA's epart through
  //             distributed shared memory (summed in rank order by install_partial_rows).
  auto coop_row = [&](const T* cs, int row, int parity) {
    const int g0 = suoff[row], g1 = suoff[row + 1];
    int o_lo = 0, o_hi = N;
    if (CL && G > 1) sweep_part_columns(N - (g1 - g0), rank, G, o_lo, o_hi);
    for (int t = warp; t < S; t += nwarps) {
      double v = 0.0;
      if (t != row) {
        const int shift = (CL && G > 1 && t > row) ? (g1 - g0) : 0;  // slot = column + shift (columns skip the gap)
        const int j_lo = (CL && G > 1) ? max(suoff[t], o_lo + shift) : suoff[t];
        const int j_hi = (CL && G > 1) ? min(suoff[t + 1], o_hi + shift) : suoff[t + 1];
        for (int j = j_lo + lane; j < j_hi; j += 32) v += (double)cs[j];
        v = warp_sum(v) * scale;
      }
      if (CL && G > 1) {
        if (lane < G) {
          MCH_RC(rc_write(cooperative_groups::this_cluster().map_shared_rank(rc_epart, lane) + (parity * kMaxCluster + rank) * S + t, rc_epoch, (unsigned)rank, "epart"));
          cooperative_groups::this_cluster().map_shared_rank(epart, lane)[(parity * kMaxCluster + rank) * S + t] = v;
        }
      } else if (lane == 0 && t != row) {
        E[row * S + t] = v; E[t * S + row] = v;
      }
    }
  };
  auto install_partial_rows = [&](int row, int parity) {  // E[row][t] = sum over CTAs, rank order
    for (int t = lane; t < S; t += 32) {
      if (t == row) continue;
      double v = 0.0;
      for (int r = 0; r < G; ++r) {
        MCH_RC(rc_read(rc_epart + (parity * kMaxCluster + r) * S + t, rc_epoch, (unsigned)rank, "epart"));
        v += epart[(parity * kMaxCluster + r) * S + t];
      }
      E[row * S + t] = v; E[t * S + row] = v;
    }
  };
  // CRITICAL-PATH half of a proposal: the draws of optimizer.py:214-244, the translation
  // vector, the displaced rows Q and the sweep descriptor.  The other warps are waiting at
  // barrier A while this runs; everything that is not needed to START the sweep (bond
  // energies of the trial geometry, energy-table refresh, trace) is done in `deferred`.
  auto draw_move = [&]() {
    kb = (int)pcg_bounded(rng, (uint32_t)B);                       // :214
    const int sa = bslot[2 * kb], sb = bslot[2 * kb + 1];
    const double bvx = (double)X[sb] - (double)X[sa];              // :215-217 (P[b1] - P[b0])
    const double bvy = (double)Y[sb] - (double)Y[sa];
    const double bvz = (double)Z[sb] - (double)Z[sa];
    const int side = (int)pcg_bounded(rng, 2u);                    // :225
    ms = bsu[2 * kb + side];
    const double rnd = (pcg_double(rng) - 0.5) * 2.0;              // :229
    kind = (int)pcg_bounded(rng, 2u);                              // :242-244
    const double ss = a.p.step_size;
    double vx, vy, vz;
    if (kind == 0) {                                               // :233
      vx = ((-bvx) * ss) * rnd; vy = ((-bvy) * ss) * rnd; vz = ((-bvz) * ss) * rnd;
    } else {
      // subunit centroid minus molecule centroid (:236-238) -- formed only when the draw asks for it
      // (warp-uniform branch; in fp64 these are six divisions)
      double cvx, cvy, cvz;
      if (kFast) {
        const double in = invn[ms], iN = 1.0 / (double)N;
        cvx = csum[3 * ms] * in - ctot[0] * iN;
        cvy = csum[3 * ms + 1] * in - ctot[1] * iN;
        cvz = csum[3 * ms + 2] * in - ctot[2] * iN;
      } else {
        // the molecule's sums are a pure function of the subunit sums (fixed order): formed here, when a centroid
        // move needs them, instead of after every change of a subunit sum -- the same bits, half as often
        if (kLazyCtot != 0) { __syncwarp(); refresh_ctot(); __syncwarp(); }
        const double nms = (double)(suoff[ms + 1] - suoff[ms]);
        cvx = csum[3 * ms] / nms - ctot[0] / N;
        cvy = csum[3 * ms + 1] / nms - ctot[1] / N;
        cvz = csum[3 * ms + 2] / nms - ctot[2] / N;
      }
      vx = (cvx * ss) * rnd; vy = (cvy * ss) * rnd; vz = (cvz * ss) * rnd;
    }
    tvx = (T)vx; tvy = (T)vy; tvz = (T)vz;
    MCH_CHECK(kb >= 0 && kb < B && ms >= 0 && ms < S, "bond / subunit index of the proposal");
    MCH_CHECK(sa >= 0 && sa < N && sb >= 0 && sb < N, "slot of a bond end");
  };
  auto publish_move = [&](int buf) {
    const int n = fill_rows(ms, tvx, tvy, tvz);                    // :248-252  (pos - vector)
    MCH_CHECK(n + 3 <= a.max_su + kRowSlack, "row buffer capacity (padded rows + prefetch overrun)");
    MCH_CHECK(suoff[ms] >= 0 && suoff[ms] <= suoff[ms + 1] && suoff[ms + 1] <= N, "slot range of the moving subunit");
    if (lane == 0) {
      ctl->nrows = n; ctl->c0 = 0; ctl->g0 = suoff[ms]; ctl->g1 = suoff[ms + 1]; ctl->c1 = N; ctl->ms = ms;
      ctl->buf = kTwoBuf ? buf : 0;
      if (kSplitD) { ctl->tv[0] = (double)tvx; ctl->tv[1] = (double)tvy; ctl->tv[2] = (double)tvz; }  // for warp 1
    }
  };
  auto propose = [&](int buf) { draw_move(); publish_move(buf); };
  // ---------------------------------------------------------------- early rejection (EARLY instantiation)
  // A move is rejected iff U_new >= U and exp(-beta (U_new - U)) <= r, r the next double of the chain's random
  // stream (mc_operations.py:46-52) -- i.e. iff  dU = U_new - U >= -ln(r) / beta.  r can be read ahead (the
  // generator is advanced past it only once the draw is certain to happen), and dU has a lower bound that costs
  // no pair evaluation:
  //   dU >= - sum_t E[ms][t] (1 - f_t)  +  (bond energy of the trial geometry - current bond energy),
  //   f_t = (1 + |tv| / d_t)^-mu,   d_t <= every current distance between an atom of ms and an atom of subunit t
  // (distance of the centroids minus the two subunit radii; rigid subunits: the radii are constants): a
  // translation by tv lengthens no distance by more than |tv|, so r'^-mu >= r^-mu (1 + |tv| / r)^-mu >= r^-mu f_t
  // for every pair.  When the bound, less an allowance for every rounding on the way (below), reaches the threshold
  // (taken from above), the move IS rejected whatever its sweep would add up to.  The control warp then consumes the
  // draw, books the rejection (fp64: the reference's inexact undo) and proposes the next move at once: the other
  // warps stay at barrier A, no rows are encoded, no sweep is started.  Decisions and results are those of the full
  // evaluation, bit for bit (a move that is accepted always gets its full sweep).  After an accepted move of subunit
  // d_ms the table's new row is still to be rebuilt (under the NEXT sweep, deferred): a proposal for another subunit
  // takes its one stale entry from the accepted sweep's column sums (bound_rejects), a second move of d_ms itself is
  // swept.
  // The bound is a sufficient condition only, so it is formed in fp32 in both precision modes (it sits on the chain's
  // critical path).  Allowance: 1e-4 of sum_t E[ms][t] (the fp32 sweep reproduces the carried table entries to ~1e-6;
  // the fp32 arithmetic here adds ~3e-6 per term), 1e-5-scaled first-order error terms of the bond part (lengths
  // to ~4e-7 relative), 1e-9 of the bond energy and 1e-12 of U for the fp64 roundings of the carried potentials
  // (fp64 mode: the sub-ulp drift of inexactly undone atoms), 1e-5 / 1e-4 A on every radius.
  float* radf = reinterpret_cast<float*>(smem + L.rad);  // radius of every subunit about its centroid, from above
  const float inv_beta_up = EARLY ? __double2float_ru(1.0 / a.p.beta) * (1.0f + 1.0e-6f) : 0.0f;
  auto subunit_radius = [&](int s2) {  // warp 0
    const double cx = csum[3 * s2] * invn[s2], cy = csum[3 * s2 + 1] * invn[s2], cz = csum[3 * s2 + 2] * invn[s2];
    float m = 0.0f;
    for (int i = suoff[s2] + lane; i < suoff[s2 + 1]; i += 32) {
      const float dx = (float)((double)X[i] - cx), dy = (float)((double)Y[i] - cy), dz = (float)((double)Z[i] - cz);
      m = fmaxf(m, dx * dx + dy * dy + dz * dz);
    }
    m = warp_max_f(m);
    if (lane == 0) radf[s2] = sqrtf(m) * (1.0f + 1.0e-5f) + 1.0e-4f;
  };
  auto bound_rejects = [&]() -> bool {  // warp 0; the drawn move (ms, tv), energy table current
    // Everything a lane contributes -- its bonds' energy change less their error terms, less what its table entries
    // can lose -- goes into ONE lane-local value and one warp reduction: what this section costs on the chain's
    // critical path is the depth of its dependent shared-memory / shuffle / special-function operations (they queue
    // behind the sweeping warps' traffic), not its instruction count (profiles/r02_phase_clocks.txt).
    const int m0 = suoff[ms], m1 = suoff[ms + 1];
    float v = 0.0f;
    // bond part: only bonds with exactly one end in ms change; dE = eps (r1 - r0)(r1 + r0 - 2 target), the trial end
    // formed exactly as bond_terms forms it; first-order error terms at 1e-5 (lengths are good to ~4e-7 relative)
    const float be = (float)a.p.bond_epsilon, tl = (float)a.p.target_bond_length;
    for (int b = lane; b < B; b += 32) {
      const int sa = bslot[2 * b], sb = bslot[2 * b + 1];
      const bool ma = sa >= m0 && sa < m1, mb = sb >= m0 && sb < m1;
      if (ma != mb) {
        const T ax = X[sa], ay = Y[sa], az = Z[sa], bx = X[sb], by = Y[sb], bz = Z[sb];
        const float ex = (float)(ax - bx), ey = (float)(ay - by), ez = (float)(az - bz);
        const T axt = ma ? ax - tvx : ax, ayt = ma ? ay - tvy : ay, azt = ma ? az - tvz : az;
        const T bxt = mb ? bx - tvx : bx, byt = mb ? by - tvy : by, bzt = mb ? bz - tvz : bz;
        const float fx = (float)(axt - bxt), fy = (float)(ayt - byt), fz = (float)(azt - bzt);
        const float r0 = sqrtf(ex * ex + ey * ey + ez * ez), r1 = sqrtf(fx * fx + fy * fy + fz * fz);
        const float dr = r1 - r0, sm = r1 + r0 - 2.0f * tl;
        v = fmaf(be * dr, sm, v);
        v = fmaf(-1.0e-5f * fabsf(be) * (fmaxf(r0, r1) + fabsf(tl)), fabsf(sm) + fabsf(dr), v);
      }
    }
    // -ln(r) / beta from above: r rounded down, the fast logarithm's error (2^-21 absolute near 1, 3 ulp elsewhere)
    Pcg64 peek = rng;
    const double r = pcg_double(peek);
    const float thr = ((-__logf(__double2float_rd(r))) * (1.0f + 1.0e-5f) + 2.0e-6f) * inv_beta_up;
    // parity mode (four warps per chain, long sweeps, a control section that re-sums coordinates after every
    // rejection anyway): staged -- the bound is at most the bond part, so a move whose bond part stays below the
    // threshold is swept without looking at the table (measured: 2.97e7 against 2.94e7 for the one-reduction form;
    // fast mode the other way round, 6.87e7 against 6.98e7)
    float bond_part = 0.0f;
    if (!kFast) {
      bond_part = warp_sum_f(v);
      v = 0.0f;
      if (!(bond_part >= thr)) return false;
    }
    // non-bonded part: what the table entries of ms can lose at most
    const float tx = (float)tvx, ty = (float)tvy, tz = (float)tvz;
    const float tvn = sqrtf(tx * tx + ty * ty + tz * tz) * (1.0f + 1.0e-6f);
    const double cmx = csum[3 * ms] * invn[ms], cmy = csum[3 * ms + 1] * invn[ms], cmz = csum[3 * ms + 2] * invn[ms];
    const float rms = radf[ms];
    auto may_go = [&](int t) -> float {  // upper bound of 1 - f_t
      const float dx = (float)(csum[3 * t] * invn[t] - cmx), dy = (float)(csum[3 * t + 1] * invn[t] - cmy),
                  dz = (float)(csum[3 * t + 2] * invn[t] - cmz);
      const float d = sqrtf(dx * dx + dy * dy + dz * dz) * (1.0f - 1.0e-6f) - rms - radf[t];
      float gone = 1.0f;  // touching or overlapping spheres bound nothing: the whole entry may go
      if (d > 0.0f) { const float q = __fdividef(d, d + tvn); gone = 1.0f - q * q * q; }
      return gone;
    };
    // after an accepted move of d_ms (its row of the table is rebuilt under the NEXT sweep) the entry between ms and
    // d_ms is stale: its new value is scale * (the accepted sweep's column sums over the slots of ms), taken from
    // above, lane by lane; 1e-4 of it joins the allowance
    if (d_accepted) {
      const T* cs = colsum + d_buf * colsum_stride;
      float part = 0.0f;
      for (int j = m0 + lane; j < m1; j += 32) part += (float)cs[j];
      v = fmaf(-(may_go(d_ms) * (1.0f + 1.0e-5f) + 1.0e-4f) * ((float)scale * (1.0f + 1.0e-6f)), part, v);
    }
    for (int t = lane; t < S; t += 32) {
      if (t == ms || (d_accepted && t == d_ms)) continue;
      // an infinite or NaN entry (coincident atoms) poisons the bound: no early decision
      v = fmaf(-(float)E[ms * S + t] * (1.0f + 1.0e-7f), may_go(t), v);
    }
    v = warp_sum_f(v) + bond_part;
    const float allowance = 1.0e-4f * fabsf((float)erow[ms]) +
                            (float)(1.0e-9 * fabs(U - Unb) + 1.0e-12 * (fabs(U) + fabs(Unb)));
    return v - allowance >= thr;  // false for NaN
  };
  // exp(x) > r for x <= 0, r in [0, 1): decided by a fast fp32 exponential whenever it is
  // clear of r by more than its own error, by the fp64 exponential otherwise -- the outcome
  // always equals the fp64 comparison (reference mc_operations.py:49-52).
  auto boltzmann_beats = [&](double x, double r) {
    const float ef = __expf((float)x);
    const float rf = (float)r;
    if (fabsf(ef - rf) > 2.0e-3f * fmaxf(ef, rf)) return ef > rf;
    return exp(x) > r;
  };

  // EARLY: draw the move of step `decided + 1`; while the bound rejects it, book the rejection and draw again.  The
  // first move the bound does not reject is published for the sweep.  Returns the last step decided this way (its
  // successor is the step whose sweep comes next; a.num_steps - 1 if the chain has run out of steps).  `nbuf`: the
  // column-sum buffer the next sweep writes.
  auto propose_next = [&](int decided, int nbuf) -> int {
    while (decided + 1 < a.num_steps) {
      draw_move();
      // while the energy table waits for an accepted move's row, a second move of the same subunit is swept
      if ((d_accepted && ms == d_ms) || !bound_rejects()) { publish_move(nbuf); break; }
      ++decided;                      // step `decided` is over: rejected without a single pair term
      (void)pcg_double(rng);          // U_new >= U is certain, so the accept test's draw is consumed (mc_operations.py:50)
      if (a.inexact_undo) {           // reference optimizer.py:273-277, as after a swept rejection
        const int r0 = suoff[ms], n = suoff[ms + 1] - r0;
        for (int i = lane; i < n; i += 32) {
          X[r0 + i] = (X[r0 + i] - tvx) - (-tvx); Y[r0 + i] = (Y[r0 + i] - tvy) - (-tvy); Z[r0 + i] = (Z[r0 + i] - tvz) - (-tvz);
        }
        __syncwarp();
        refresh_csum(ms);
        __syncwarp();
        if (kLazyCtot == 0) { refresh_ctot(); __syncwarp(); }
      }
      info = kb | (kind << 17) | (ms << 18);
    }
    return decided;
  };

  // ---------------------------------------------------------------- step 0: full evaluation
  // For every subunit s: cross terms against all later subunits (each cross pair once)
  // and the strict lower triangle inside s.
  if (warp == 0) {
    for (int s = 0; s < S; ++s) refresh_csum(s);
    __syncwarp();
    refresh_ctot();
    if (lane == 0) ctl->U = 0.0;  // accumulates the intra-subunit sum during step 0
    __syncwarp();
  }
  // On a cluster the subunits are dealt to the CTAs (boustrophedon, so that the shrinking column
  // ranges balance): a CTA computes complete rows E[s][t > s] and the intra-subunit sums of ITS
  // subunits, then pushes them to every peer -- one cluster barrier for the whole of step 0.
  auto step0_owner = [&](int s) { const int q = s / G, m = s - q * G; return (q & 1) ? G - 1 - m : m; };
  for (int s = 0; s < S; ++s) {
    if (CL && G > 1 && step0_owner(s) != rank) continue;
    if (warp == 0) {
      const int n = fill_rows(s, (T)0, (T)0, (T)0);
      if (lane == 0) { ctl->nrows = n; ctl->c0 = suoff[s + 1]; ctl->g0 = suoff[s + 1]; ctl->g1 = suoff[s + 1]; ctl->c1 = N; ctl->ms = s; ctl->buf = 0; }
    }
    __syncthreads();
    {
      Cols cross{ctl->c0, ctl->g0, ctl->g1, ctl->c1};
      const Centre<T> ctr{(T)ctl->cx, (T)ctl->cy, (T)ctl->cz};
      (void)pair_sweep<T, MUK, false, true, kXF, kX2, kRot3>(Q, ctl->nrows, X, Y, Z, colsum, cross, term, ctr);
      const int r0 = suoff[s], nrows = suoff[s + 1] - r0;
      Cols own{r0, r0 + nrows, r0 + nrows, r0 + nrows};
      double tri = pair_sweep<T, MUK, true, false, kXF, kX2>(Q, nrows, X, Y, Z, (T*)nullptr, own, term, ctr);
      tri = warp_sum(tri);
      if (lane == 0) wsum[32 + warp] = tri;
    }
    __syncthreads();
    if (warp == 0) {
      double intra = 0.0;
      for (int w = 0; w < nwarps; ++w) intra += wsum[32 + w];
      if (lane == 0) ctl->U += intra;
      regroup(colsum, s, s + 1);
    }
  }
  if (CL && G > 1) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __syncthreads();  // this CTA's rows and its intra-subunit sum are complete
    for (int r = 0; r < G; ++r) {
      if (r == rank) continue;
      double* peer = cluster.map_shared_rank(E, r);
      for (int own = 0; own < S; ++own) {
        if (step0_owner(own) != rank) continue;
        for (int t = own + 1 + tid; t < S; t += nt) {
          const double v = E[own * S + t];
          MCH_RC(rc_write(cluster.map_shared_rank(rc_E, r) + own * S + t, rc_epoch, (unsigned)rank, "energy table"));
          MCH_RC(rc_write(cluster.map_shared_rank(rc_E, r) + t * S + own, rc_epoch, (unsigned)rank, "energy table"));
          peer[own * S + t] = v; peer[t * S + own] = v;
        }
      }
    }
    if (tid < G) {
      MCH_RC(rc_write(cluster.map_shared_rank(rc_xsum, tid) + rank, rc_epoch, (unsigned)rank, "sweep totals"));
      cluster.map_shared_rank(xsum, tid)[rank] = ctl->U;  // parity-0 slots (the step loop starts with parity 1)
    }
    cluster.sync();
    MCH_RC(++rc_epoch);
    if (tid == 0) {
      double u = 0.0;
      for (int r = 0; r < G; ++r) { MCH_RC(rc_read(rc_xsum + r, rc_epoch, (unsigned)rank, "sweep totals")); u += xsum[r]; }
      ctl->U = u;
    }
    __syncthreads();
  }
  if (warp == 0) {
    __syncwarp();
    rng = pcg_load(a.rng + 6LL * c);
    pair_count = (unsigned long long)N * (unsigned long long)(N - 1) / 2ULL;  // step 0: every pair once
    const double e_intra = ctl->U * scale;
    double cross = 0.0;
    for (int k = lane; k < S * S; k += 32) {
      const int s = k / S, t = k - s * S;
      if (CL && G > 1 && t != s) { MCH_RC(rc_read(rc_E + k, rc_epoch, (unsigned)rank, "energy table")); }
      if (t > s) cross += E[k];
    }
    Unb = warp_sum(cross) + e_intra;
    refresh_erow();
    double longest;
    U = Unb + bond_terms(-1, (T)0, (T)0, (T)0, longest);
    info = -1;
    if (lane == 0) { ctl->d_accepted = 0; ctl->d_ms = 0; }
    write_trace(0, longest);
    if (a.num_steps <= 1 && lane == 0 && rank == 0) a.maxd[c] = longest;
    __syncwarp();
    if (a.num_steps > 1) {
      if constexpr (EARLY) {
        for (int s2 = 0; s2 < S; ++s2) subunit_radius(s2);
        __syncwarp();
        const int decided = propose_next(0, 1);
        if (lane == 0) ctl->next_step = decided + 1;
      } else {
        propose(1);
        if (!kDefer) { double unused; Ub_trial = bond_terms(ms, tvx, tvy, tvz, unused, true, false); }
      }
    }
  }

#ifdef MCH_TIMING  // tuning builds only (profiles/phase_clocks.py): where a swept step's cycles go, per warp
  long long tm_a = 0, tm_d = 0, tm_s = 0, tm_b = 0;               // stamps of the current step
  long long tm_D = 0, tm_S = 0, tm_W = 0, tm_C = 0, tm_n = 0;     // sums: deferred, own sweep, wait at B, control, swept steps
#endif
  // ---------------------------------------------------------------- Monte-Carlo steps
  for (int step = 1; step < a.num_steps; ++step) {
#ifdef MCH_TIMING
    if (warp == 0 && tm_b != 0) tm_C += clock64() - tm_b;  // barrier B release -> arrival at A
#endif
    __syncthreads();  // A: move published, positions of step-1 final
#ifdef MCH_TIMING
    tm_a = clock64();
#endif
    if constexpr (EARLY) {  // the control warp may have decided several steps since the last sweep
      step = ctl->next_step;
      if (step >= a.num_steps) break;
    }
    if (a.traj && rank == 0) {
      const long long frame = ((long long)c * a.num_steps + (step - 1)) * N * 3;
      for (int n = tid; n < N; n += nt) {
        const long long dst = frame + 3LL * a.perm[n];
        store_io<Tio>(a.traj, dst + 0, (double)X[n] + ctl->origin[0]);
        store_io<Tio>(a.traj, dst + 1, (double)Y[n] + ctl->origin[1]);
        store_io<Tio>(a.traj, dst + 2, (double)Z[n] + ctl->origin[2]);
      }
    }
    if (kCoop && ctl->d_accepted != 0) {  // uniform over the CTA: published by warp 0 before barrier A
      // the accepted sweep's buffer: the previous step's -- EARLY: steps may have been decided since, the buffer is published
      coop_row(colsum + (kTwoBuf ? (EARLY ? ctl->d_buf : ((step - 1) & 1)) : 0) * colsum_stride, ctl->d_ms, step & 1);
      __syncthreads();
    }
    if (warp == 0) {
      // DEFERRED half of the control work, overlapped with the other warps' sweep:
      // (1) the previous move was accepted -> rebuild row/column d_ms of the energy table
      //     from the column sums of the PREVIOUS sweep (the other buffer),
      // (2) trace record of the previous step, (3) bond energy of the trial geometry.
      if (kDefer) {
        if (d_accepted) {
          if (CL && G > 1) {
            // coop_row pushed this CTA's share to every CTA; installed after this step's cluster barrier
          } else {
            if (!kCoop) regroup(colsum + d_buf * colsum_stride, d_ms, 0);
            refresh_erow();
            if constexpr (EARLY) subunit_radius(d_ms);
            d_accepted = false;
          }
        }
        if (d_step > 0 && (a.trace_f || a.trace_i)) {
          double longest = 0.0;
          if (a.trace_f) (void)bond_terms(-1, (T)0, (T)0, (T)0, longest, false, true);
          write_trace(d_step, longest);
        }
        if (!kSplitD) {
          double unused;
          Ub_trial = bond_terms(ms, tvx, tvy, tvz, unused, true, false);
        }
      } else if (kCoop && d_accepted) {
        refresh_erow();
        d_accepted = false;
      }
      park();
    } else if (kSplitD && warp == 1) {
      // warp 1's share of the deferred work: the bond energy of the published trial geometry (read after barrier B)
      double unused;
      const double ub = bond_terms(ctl->ms, (T)ctl->tv[0], (T)ctl->tv[1], (T)ctl->tv[2], unused, true, false);
      if (lane == 0) ctl->ub_w1 = ub;
    }
#ifdef MCH_TIMING
    tm_d = clock64();
#endif
    {
      Cols others{ctl->c0, ctl->g0, ctl->g1, ctl->c1};
      const Centre<T> ctr{(T)ctl->cx, (T)ctl->cy, (T)ctl->cz};
      // warp 0 spends roughly the time of 100 rows x 32 columns on the deferred control work:
      // one pair of column units less for subunits of up to ~50 atoms, nothing for large ones
      constexpr int kHandicap = (EARLY && sizeof(T) == 4) ? MCH_EARLY_HANDICAP : 1;
      const int handicap = kDefer ? min(kHandicap, kHandicap * 50 / ctl->nrows) : 0;
      double part = pair_sweep<T, MUK, false, true, kXF, kX2, kRot3>(Q, ctl->nrows, X, Y, Z, colsum + ctl->buf * colsum_stride, others, term, ctr, handicap, rank, G);
      part = warp_sum(part);
      if (lane == 0) wsum[warp] = part;
    }
#ifdef MCH_TIMING
    tm_s = clock64();
#endif
    __syncthreads();  // B: column sums and warp totals published
#ifdef MCH_TIMING
    tm_b = clock64();
    tm_D += tm_d - tm_a; tm_S += tm_s - tm_d; tm_W += tm_b - tm_s; ++tm_n;
#endif
    if (CL && G > 1) {
      // every CTA of the chain learns every CTA's sweep total (slots double buffered by step parity:
      // a CTA that runs ahead cannot overwrite a slot a slower one still has to read)
      namespace cg = cooperative_groups;
      cg::cluster_group cluster = cg::this_cluster();
      if (warp == 0) {
        double mine = 0.0;
        for (int w = 0; w < nwarps; ++w) mine += wsum[w];
        if (lane < G) {
          MCH_RC(rc_write(cluster.map_shared_rank(rc_xsum, lane) + MCH_RC_PARITY(step) * kMaxCluster + rank, rc_epoch, (unsigned)rank, "sweep totals"));
          cluster.map_shared_rank(xsum, lane)[MCH_RC_PARITY(step) * kMaxCluster + rank] = mine;
        }
      }
      cluster.sync();
      MCH_RC(++rc_epoch);
    }
    if (warp == 0) {
      unpark();
      if (kSplitD) Ub_trial = ctl->ub_w1;
      double fresh = 0.0;
      if (CL && G > 1) {
        for (int r = 0; r < G; ++r) {
          MCH_RC(rc_read(rc_xsum + MCH_RC_PARITY(step) * kMaxCluster + r, rc_epoch, (unsigned)rank, "sweep totals"));
          fresh += xsum[MCH_RC_PARITY(step) * kMaxCluster + r];
        }
        if (d_accepted) {  // the row pushed under this step's sweep: needed from here on (erow below)
          install_partial_rows(d_ms, step & 1);
          refresh_erow();
          d_accepted = false;
        }
      } else {
        for (int w = 0; w < nwarps; ++w) fresh += wsum[w];
      }
      fresh *= scale;
      const double Unb_new = Unb + (fresh - erow[ms]);
      const double U_new = Unb_new + Ub_trial;
      bool ok;
      if (U_new < U) ok = true;                                                       // mc_operations.py:46-47
      else ok = boltzmann_beats(-a.p.beta * (U_new - U), pcg_double(rng));            // mc_operations.py:49-52
      const int r0 = suoff[ms], n = suoff[ms + 1] - r0;
      pair_count += (unsigned long long)n * (unsigned long long)(N - n);
      if (ok) {
        for (int i = lane; i < n; i += 32) { X[r0 + i] -= tvx; Y[r0 + i] -= tvy; Z[r0 + i] -= tvz; }  // = the swept rows
        U = U_new; Unb = Unb_new; ++n_acc;
        __syncwarp();
        if (kFast) {
          // rigid translation: every sum moves by n * tv (no re-summation of the subunit)
          if (lane < 3) {
            const double d = (double)n * (double)(lane == 0 ? tvx : (lane == 1 ? tvy : tvz));
            csum[3 * ms + lane] -= d;
            ctot[lane] -= d;
          }
        } else {
          refresh_csum(ms);
          __syncwarp();
          if (kLazyCtot == 0) refresh_ctot();
        }
      } else if (a.inexact_undo) {
        // reference optimizer.py:273-277: translate the moved atoms by -tv again
        for (int i = lane; i < n; i += 32) {
          X[r0 + i] = (X[r0 + i] - tvx) - (-tvx); Y[r0 + i] = (Y[r0 + i] - tvy) - (-tvy); Z[r0 + i] = (Z[r0 + i] - tvz) - (-tvz);
        }
        __syncwarp();
        refresh_csum(ms);
        __syncwarp();
        if (kLazyCtot == 0) refresh_ctot();
      }
      __syncwarp();
      info = kb | ((ok ? 1 : 0) << 16) | (kind << 17) | (ms << 18);
      d_accepted = ok; d_ms = ms; d_buf = kTwoBuf ? (EARLY ? ctl->buf : (step & 1)) : 0; d_step = step;
      if (kCoop && lane == 0) { ctl->d_accepted = ok ? 1 : 0; ctl->d_ms = ms; ctl->d_buf = d_buf; }  // for coop_row after barrier A
      if (!kDefer) {
        if (!kCoop) {
          if (ok) { regroup(colsum + d_buf * colsum_stride, ms, 0); refresh_erow(); }
          d_accepted = false;
        }
        if (step + 1 < a.num_steps && (a.trace_f || a.trace_i)) {
          double longest = 0.0;
          if (a.trace_f) (void)bond_terms(-1, (T)0, (T)0, (T)0, longest, false, true);
          write_trace(step, longest);
        }
      }
      if constexpr (EARLY) {
        const int decided = propose_next(step, (kTwoBuf ? ctl->buf : 0) ^ 1);
        if (lane == 0) ctl->next_step = decided + 1;
      } else if (step + 1 < a.num_steps) {
        propose((step + 1) & 1);
        if (!kDefer) { double unused; Ub_trial = bond_terms(ms, tvx, tvy, tvz, unused, true, false); }
      }
    }
  }
  if (warp == 0 && a.num_steps > 1) {  // the last step has no following sweep to hide behind
    double longest;
    (void)bond_terms(-1, (T)0, (T)0, (T)0, longest, false, true);
    write_trace(a.num_steps - 1, longest);
    if (lane == 0 && rank == 0) a.maxd[c] = longest;
  }
  __syncthreads();
  if (CL && G > 1) {
    cooperative_groups::this_cluster().sync();  // no CTA leaves while another may still write into it
    if (rank != 0) return;                      // CTA 0 reports the chain
  }

  // ---------------------------------------------------------------- epilogue
  for (int n = tid; n < N; n += nt) {
    const long long at = 3LL * a.perm[n];
    const double px = (double)X[n] + ctl->origin[0], py = (double)Y[n] + ctl->origin[1], pz = (double)Z[n] + ctl->origin[2];
    const long long dst = (long long)c * N * 3 + at;
    store_io<Tio>(a.pos_out, dst + 0, px); store_io<Tio>(a.pos_out, dst + 1, py); store_io<Tio>(a.pos_out, dst + 2, pz);
    if (a.traj) {
      const long long fr = ((long long)c * a.num_steps + (a.num_steps > 0 ? a.num_steps - 1 : 0)) * N * 3 + at;
      store_io<Tio>(a.traj, fr + 0, px); store_io<Tio>(a.traj, fr + 1, py); store_io<Tio>(a.traj, fr + 2, pz);
    }
  }
#ifdef MCH_TIMING  // the first 12 position values of the chain are overwritten with the cycle sums (warp 0, warp 1)
  __syncthreads();
  if (lane == 0 && warp < 2 && rank == 0 && sizeof(Tio) == 8) {
    double* o = reinterpret_cast<double*>(a.pos_out) + (long long)c * N * 3 + 6 * warp;
    o[0] = (double)tm_D; o[1] = (double)tm_S; o[2] = (double)tm_W; o[3] = (double)tm_C; o[4] = (double)tm_n; o[5] = 0.0;
  }
#endif
  if (warp == 0 && lane == 0) {
    a.U[c] = U; a.Unb[c] = Unb; a.n_accept[c] = n_acc; a.last_info[c] = info;
    if (a.pairs) a.pairs[c] = pair_count;
    pcg_store(rng, a.rng + 6LL * c);
  }
}

#endif  // __CUDACC__

}  // namespace mch
