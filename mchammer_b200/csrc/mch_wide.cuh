// mch_wide.cuh -- the Metropolis chain with its STATE SPLIT over a thread-block cluster, and with
// explicit moving sets: the general form of mch_optimize_kernel for inputs that kernel refuses.
//
//   * molecules that do not fit one CTA's shared memory (the reference has no size limit,
//     optimizer.py:114-120): CTA `rank` of the chain's cluster of G CTAs owns the slots
//     [rank * chunk, (rank + 1) * chunk) -- positions live ONLY there;
//   * any number of subunits (no S x S energy table);
//   * overlapping subunits with the reference's semantics (optimizer.py:220-226, :248-252): the
//     chosen subunit's ROW LIST (every atom it lists, each once) moves, whoever "owns" the atoms;
//     its centroid is the mean over the listed ids (repeats counted, molecule.py:204-207).
//
// Per step (same draws in the same order as the reference; every CTA of the cluster runs the
// control section redundantly -- same inputs, same arithmetic, same decisions):
//   1. draws -> moving subunit ms; every CTA PULLS the current positions of ms's atoms from their
//      owners through distributed shared memory (rows R0), forms the translation vector and the
//      encoded rows of the old and of the displaced geometry;
//   2. sweeps rows x (own columns minus ms's atoms) twice -- displaced and undisplaced rows --
//      through the same pair_sweep tiles as the one-CTA kernel; dU_nb = scale (new - old).
//      (No energy table: its bookkeeping presumes a partition.  Two sweeps per step is the price.)
//   3. CTA totals are exchanged through DSMEM (cluster barrier 1), every CTA decides, owners commit
//      (or redo the reference's inexact undo), re-sum their coordinate sums, publish them
//      (cluster barrier 2: positions are final before the next step's pull).
// Replicated per CTA: random stream, potentials, positions of the bond ends (updated with the same
// arithmetic as the owners' copies, so they stay bit-identical), the molecule's coordinate sums.
#pragma once

#include "mch_device.cuh"
#include "mch_host.h"
#include "mch_optimize.cuh"

#ifdef __CUDACC__
#include <cooperative_groups.h>
#endif

namespace mch {

constexpr int kWideMaxCluster = 16;
constexpr int kWideBlockRows = 64;  // step 0 pulls rows in blocks of this many atoms

struct WideArgs {
  const void* pos_in;
  long long pos_in_stride;
  void* pos_out;
  int C, N, S, B;        // S = number of row lists (the caller's subunits)
  int G;                 // CTAs per chain (cluster size)
  int chunk;             // slots per CTA
  int max_rows;          // longest row list
  const int* perm;       // [N] slot -> atom id
  const int* bond_slot;  // [B][2] slot of each bond end, caller's order inside the pair
  const int* bond_su;    // [B][2] first-owner subunit of each end
  const int* row_off;    // [S+1]
  const int* row_slot;   // [R] ascending per list
  const int* row_w;      // [R] multiplicity in the caller's list
  const int* run_off;    // [S+1]
  const int* runs;       // [K][2] contiguous slot ranges of each row list
  Params p;
  int mu_int;
  int num_steps;
  uint64_t* rng;
  double* U;
  double* Unb;
  double* maxd;
  int* n_accept;
  int* last_info;
  double* trace_f;
  int* trace_i;
  void* traj;
  unsigned long long* pairs;
  int inexact_undo;
};

struct WideCtl {
  int ms, n, np, ok, touched;
  double tv[3];
  double ctr[3];     // centre of the expanded row encoding (old geometry)
  double origin[3];
};

struct WideLayout { int x, y, z, qn, qo, r0, rw, bpos, bslot, bsu, xs, x0, cs, wsum, ctl, rc, total; };

MCH_HD WideLayout wide_layout(int chunk, int B, int max_rows, int tsize) {
  WideLayout L;
  const int np = align_up(chunk, 8);
  const int rows_cap = (max_rows > kWideBlockRows ? max_rows : kWideBlockRows) + kRowSlack;
  int o = 0;
  L.x = o; o += np * tsize;
  L.y = o; o += np * tsize;
  L.z = o; o += np * tsize;
  o = align_up(o, 32);
  L.qn = o; o += rows_cap * row_bytes(tsize);
  L.qo = o; o += rows_cap * row_bytes(tsize);
  L.r0 = o; o += 3 * rows_cap * tsize;
  L.rw = o; o += align_up(rows_cap, 4) * 4;
  o = align_up(o, 16);
  L.bpos = o; o += 3 * align_up(2 * B, 2) * tsize;
  L.bslot = o; o += align_up(2 * B, 4) * 4;
  L.bsu = o; o += align_up(2 * B, 4) * 4;
  o = align_up(o, 16);
  L.xs = o; o += kWideMaxCluster * 2 * 8;
  L.x0 = o; o += kWideMaxCluster * 8;  // step-0 totals: slots of their own (a fast CTA writes its step-1 totals into
                                       // xs while a slow one may still be reading step 0's -- found by the check build)
  L.cs = o; o += kWideMaxCluster * 3 * 8;
  L.wsum = o; o += 3 * 32 * 8;
  L.ctl = o; o += (int)sizeof(WideCtl);
  L.rc = o;
#ifdef MCH_RACECHECK  // shadows: one per own slot (x, y, z together), per sweep-total entry, per coordinate-sum triple
  o += (np + kWideMaxCluster * 2 + kWideMaxCluster + kWideMaxCluster) * 8;
#endif
  L.total = align_up(o, 16);
  return L;
}

#ifdef __CUDACC__

template <typename T, typename Tio, int MUK>
__global__ void __launch_bounds__(512, 1)
mch_optimize_wide_kernel(const __grid_constant__ WideArgs a) {
  namespace cg = cooperative_groups;
  extern __shared__ __align__(32) unsigned char smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int G = a.G;
  const int c = (int)blockIdx.x / G, rank = (int)blockIdx.x % G;  // == %cluster_ctarank for a (G,1,1) cluster
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nt = blockDim.x, nwarps = nt >> 5;
  const int N = a.N, B = a.B, chunk = a.chunk;
  const int lo = rank * chunk < N ? rank * chunk : N;
  const int hi = lo + chunk < N ? lo + chunk : N;
  const int cnt = hi - lo;  // own slots [lo, hi)
  const int rows_cap = (a.max_rows > kWideBlockRows ? a.max_rows : kWideBlockRows) + kRowSlack;

  const WideLayout L = wide_layout(chunk, B, a.max_rows, (int)sizeof(T));
  T* X = reinterpret_cast<T*>(smem + L.x);
  T* Y = reinterpret_cast<T*>(smem + L.y);
  T* Z = reinterpret_cast<T*>(smem + L.z);
  Row4<T>* Qn = reinterpret_cast<Row4<T>*>(smem + L.qn);
  Row4<T>* Qo = reinterpret_cast<Row4<T>*>(smem + L.qo);
  T* R0 = reinterpret_cast<T*>(smem + L.r0);  // [3][rows_cap]: current positions of the moving atoms
  int* Rw = reinterpret_cast<int*>(smem + L.rw);
  T* bpos = reinterpret_cast<T*>(smem + L.bpos);  // [3][2B]: positions of the bond ends (replica)
  const int bstride = align_up(2 * B, 2);
  int* bslot = reinterpret_cast<int*>(smem + L.bslot);
  int* bsu = reinterpret_cast<int*>(smem + L.bsu);
  double* xs = reinterpret_cast<double*>(smem + L.xs);  // [G][2] per-CTA sweep totals (new, old)
  double* x0 = reinterpret_cast<double*>(smem + L.x0);  // [G] per-CTA totals of step 0
  double* cs = reinterpret_cast<double*>(smem + L.cs);  // [G][3] per-CTA coordinate sums
  double* wsum = reinterpret_cast<double*>(smem + L.wsum);
  WideCtl* ctl = reinterpret_cast<WideCtl*>(smem + L.ctl);
#ifdef MCH_RACECHECK
  const int rc_np = align_up(chunk, 8);
  RcShadow* rc_pos = reinterpret_cast<RcShadow*>(smem + L.rc);  // [chunk]
  RcShadow* rc_xs = rc_pos + rc_np;                             // [G][2]
  RcShadow* rc_cs = rc_xs + kWideMaxCluster * 2;                // [G]
  RcShadow* rc_x0 = rc_cs + kWideMaxCluster;                    // [G]
  unsigned rc_epoch = 1u;                                       // cluster barriers passed + 1
  for (int k = tid; k < rc_np + kWideMaxCluster * 4; k += nt) { rc_pos[k].w = 0u; rc_pos[k].r = 0u; }
  __syncthreads();
#endif

  constexpr bool kFast = sizeof(T) == 4;
  constexpr bool kXF = kFast ? (MCH_XF != 0) : (MCH_XF64 != 0);
  constexpr bool kX2 = kFast && kXF && (MCH_F32X2 != 0);
  constexpr int kRot3 = kX2 && (MCH_X2_ROT3 != 0) ? 2 : 0;
  const PairTerm<T, MUK> term(a.p.nonbond_mu, a.mu_int);
  const double scale = a.p.nonbond_epsilon * pow(a.p.nonbond_sigma, a.p.nonbond_mu);
  const long long in_base = (long long)c * a.pos_in_stride;

  // sum over the block of one double per thread, returned to every thread (uses wsum[0..31], two barriers)
  auto block_sum3 = [&](double& sx, double& sy, double& sz) {
    sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
    __syncthreads();
    if (lane == 0) { wsum[warp] = sx; wsum[32 + warp] = sy; wsum[64 + warp] = sz; }
    __syncthreads();
    sx = sy = sz = 0.0;
    for (int w = 0; w < nwarps; ++w) { sx += wsum[w]; sy += wsum[32 + w]; sz += wsum[64 + w]; }
  };
  // own coordinate sums -> every CTA's cs[rank] (all threads call; ends with the values published)
  auto publish_own_sums = [&]() {
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int l = tid; l < cnt; l += nt) { sx += (double)X[l]; sy += (double)Y[l]; sz += (double)Z[l]; }
    block_sum3(sx, sy, sz);
    if (warp == 0 && lane < G) {
      double* peer = cluster.map_shared_rank(cs, lane);
      MCH_RC(rc_write(cluster.map_shared_rank(rc_cs, lane) + rank, rc_epoch, (unsigned)rank, "coordinate sums"));
      peer[3 * rank] = sx; peer[3 * rank + 1] = sy; peer[3 * rank + 2] = sz;
    }
  };
  // positions of `count` atoms given by slot -> dst[3][rows_cap-strided], read from their owners (DSMEM)
  auto pull = [&](int slot, T& x, T& y, T& z) {
    const int q = slot / chunk, l = slot - q * chunk;
    MCH_CHECK(slot >= 0 && slot < N && q < G, "slot of a pulled atom");
    MCH_RC(rc_read(cluster.map_shared_rank(rc_pos, q) + l, rc_epoch, (unsigned)rank, "positions"));
    x = cluster.map_shared_rank(X, q)[l];
    y = cluster.map_shared_rank(Y, q)[l];
    z = cluster.map_shared_rank(Z, q)[l];
  };
  auto in_rows_of = [&](int s, int slot) {  // is `slot` one of the atoms subunit s lists?
    for (int k = a.run_off[s]; k < a.run_off[s + 1]; ++k)
      if (slot >= a.runs[2 * k] && slot < a.runs[2 * k + 1]) return true;
    return false;
  };
  // bond energy and longest bond from the replica; ends listed by subunit `moved` are displaced by -d
  auto bond_terms = [&](int moved, T dx, T dy, T dz, double& longest) {
    double e = 0.0, far = 0.0;
    for (int b = lane; b < B; b += 32) {
      const int ka = 2 * b, kb2 = 2 * b + 1;
      const bool ma = moved >= 0 && in_rows_of(moved, bslot[ka]), mb = moved >= 0 && in_rows_of(moved, bslot[kb2]);
      const T ax = ma ? bpos[ka] - dx : bpos[ka], ay = ma ? bpos[bstride + ka] - dy : bpos[bstride + ka];
      const T az = ma ? bpos[2 * bstride + ka] - dz : bpos[2 * bstride + ka];
      const T bx = mb ? bpos[kb2] - dx : bpos[kb2], by = mb ? bpos[bstride + kb2] - dy : bpos[bstride + kb2];
      const T bz = mb ? bpos[2 * bstride + kb2] - dz : bpos[2 * bstride + kb2];
      const double ex = (double)ax - (double)bx, ey = (double)ay - (double)by, ez = (double)az - (double)bz;
      const double r = sqrt(ex * ex + ey * ey + ez * ez);
      const double d = r - a.p.target_bond_length;
      e += a.p.bond_epsilon * (d * d);
      far = fmax(far, r);
    }
    longest = warp_max(far);
    return warp_sum(e);
  };
  auto boltzmann_beats = [&](double x, double r) {  // exp(x) > r, decided in fp32 when that is safe
    const float ef = __expf((float)x);
    const float rf = (float)r;
    if (fabsf(ef - rf) > 2.0e-3f * fmaxf(ef, rf)) return ef > rf;
    return exp(x) > r;
  };
  // Sweep `nrows_p` encoded rows against the own columns [0, cnt) minus the atoms subunit `skip` lists
  // (skip < 0: all own columns); the gap-free stretches between its runs are handed to pair_sweep two
  // at a time (one gap per call).  Returns this thread's partial sum.
  auto sweep_own = [&](const Row4<T>* rows, int nrows_p, int skip, const Centre<T>& ctr) {
    double part = 0.0;
    int have = 0, s0b = 0, s0e = 0;  // a pending first stretch
    int cur = 0;
    auto emit = [&](int b, int e) {  // next gap-free stretch [b, e) in local slots (may be empty)
      if (!have) { s0b = b; s0e = e; have = 1; return; }
      const Cols cols{s0b, s0e, b, e};
      part += pair_sweep<T, MUK, false, false, kXF, kX2, kRot3>(rows, nrows_p, X, Y, Z, (T*)nullptr, cols, term, ctr);
      have = 0;
    };
    if (skip >= 0) {
      for (int k = a.run_off[skip]; k < a.run_off[skip + 1]; ++k) {
        int b = a.runs[2 * k] - lo, e = a.runs[2 * k + 1] - lo;
        b = b < 0 ? 0 : b; e = e > cnt ? cnt : e;
        if (b >= e) continue;  // the run lies outside this CTA's slots
        emit(cur, b);
        cur = e;
      }
    }
    emit(cur, cnt);
    if (have) {
      const Cols cols{s0b, s0e, s0e, s0e};
      part += pair_sweep<T, MUK, false, false, kXF, kX2, kRot3>(rows, nrows_p, X, Y, Z, (T*)nullptr, cols, term, ctr);
    }
    return part;
  };
  auto padded = [&](int n) { return kRot3 ? padded_rows3(n) : padded_rows(n); };

  // ---------------------------------------------------------------- prologue
  double origin[3] = {0.0, 0.0, 0.0};
  if (kFast) {  // frame centred on the start geometry (every CTA sums the whole molecule: the same bits everywhere)
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int n = tid; n < N; n += nt) {
      sx += load_io<Tio>(a.pos_in, in_base + 3LL * n + 0);
      sy += load_io<Tio>(a.pos_in, in_base + 3LL * n + 1);
      sz += load_io<Tio>(a.pos_in, in_base + 3LL * n + 2);
    }
    block_sum3(sx, sy, sz);
    origin[0] = sx / N; origin[1] = sy / N; origin[2] = sz / N;
  }
  for (int l = tid; l < cnt; l += nt) {
    const long long src = in_base + 3LL * a.perm[lo + l];
    MCH_RC(rc_write(rc_pos + l, rc_epoch, (unsigned)rank, "positions"));
    X[l] = (T)(load_io<Tio>(a.pos_in, src + 0) - origin[0]);
    Y[l] = (T)(load_io<Tio>(a.pos_in, src + 1) - origin[1]);
    Z[l] = (T)(load_io<Tio>(a.pos_in, src + 2) - origin[2]);
  }
  for (int k = tid; k < 2 * B; k += nt) { bslot[k] = a.bond_slot[k]; bsu[k] = a.bond_su[k]; }
  if (tid == 0) { ctl->origin[0] = origin[0]; ctl->origin[1] = origin[1]; ctl->origin[2] = origin[2]; }
  __syncthreads();
  cluster.sync();  // every CTA's slots are loaded
  MCH_RC(++rc_epoch);
  for (int k = tid; k < 2 * B; k += nt) {
    T x, y, z;
    pull(bslot[k], x, y, z);
    bpos[k] = x; bpos[bstride + k] = y; bpos[2 * bstride + k] = z;
  }
  publish_own_sums();

  // ---------------------------------------------------------------- step 0: every pair once
  // Pairs inside the own slots: row blocks of 64 against the later own columns + the block's triangle.
  // Pairs between two CTAs' slots: computed by the CTA that is at most G/2 ranks "behind" (half ring),
  // rows pulled from the other CTA in blocks of 64, all own columns.
  double part0 = 0.0;
  for (int d = 0; d <= G / 2; ++d) {
    if (d > 0 && 2 * d == G && rank >= G / 2) break;  // the pair {r, r + G/2} belongs to the lower rank
    const int src = (rank + d) % G;
    const int s_lo = src * chunk < N ? src * chunk : N, s_hi = s_lo + chunk < N ? s_lo + chunk : N;
    if (d > 0 && G == 1) break;
    for (int b0 = s_lo; b0 < s_hi; b0 += kWideBlockRows) {
      const int nrows = (s_hi - b0) < kWideBlockRows ? (s_hi - b0) : kWideBlockRows;
      __syncthreads();  // the previous block's sweeps are done with Qn / R0
      for (int i = tid; i < nrows; i += nt) {
        T x, y, z;
        pull(b0 + i, x, y, z);
        R0[i] = x; R0[rows_cap + i] = y; R0[2 * rows_cap + i] = z;
      }
      __syncthreads();
      if (warp == 0) {  // centre of the block (expanded encoding)
        double sx = 0.0, sy = 0.0, sz = 0.0;
        for (int i = lane; i < nrows; i += 32) { sx += (double)R0[i]; sy += (double)R0[rows_cap + i]; sz += (double)R0[2 * rows_cap + i]; }
        sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
        if (lane == 0) { ctl->ctr[0] = sx / nrows; ctl->ctr[1] = sy / nrows; ctl->ctr[2] = sz / nrows; }
      }
      __syncthreads();
      const Centre<T> ctr{(T)ctl->ctr[0], (T)ctl->ctr[1], (T)ctl->ctr[2]};
      const int np = padded(nrows);
      for (int i = tid; i < np; i += nt)
        Qn[i] = i < nrows ? encode_row<T, kXF>(R0[i], R0[rows_cap + i], R0[2 * rows_cap + i], ctr) : padding_row<T, kXF>();
      __syncthreads();
      if (d == 0) {
        const int l0 = b0 - lo, l1 = l0 + nrows;
        const Cols later{l1, l1, l1, cnt};
        part0 += pair_sweep<T, MUK, false, false, kXF, kX2, kRot3>(Qn, np, X, Y, Z, (T*)nullptr, later, term, ctr);
        const Cols own{l0, l1, l1, l1};
        part0 += pair_sweep<T, MUK, true, false, kXF, kX2>(Qn, nrows, X, Y, Z, (T*)nullptr, own, term, ctr);
      } else {
        const Cols all{0, cnt, cnt, cnt};
        part0 += pair_sweep<T, MUK, false, false, kXF, kX2, kRot3>(Qn, np, X, Y, Z, (T*)nullptr, all, term, ctr);
      }
    }
  }
  {
    double unused1 = 0.0, unused2 = 0.0;
    block_sum3(part0, unused1, unused2);
    if (warp == 0 && lane < G) {
      MCH_RC(rc_write(cluster.map_shared_rank(rc_x0, lane) + rank, rc_epoch, (unsigned)rank, "step-0 totals"));
      cluster.map_shared_rank(x0, lane)[rank] = part0;
    }
  }
  cluster.sync();  // step-0 totals and coordinate sums of every CTA are visible
  MCH_RC(++rc_epoch);

  // ---------------------------------------------------------------- control state (warp 0, uniform)
  Pcg64 rng = pcg_load(a.rng + 6LL * c);
  double U = 0.0, Unb = 0.0, Ub_trial = 0.0;
  int kb = 0, kind = 0, n_acc = 0, info = -1, ms = 0;
  double rnd = 0.0;
  unsigned long long pair_count = (unsigned long long)N * (unsigned long long)(N - 1) / 2ULL;
  if (warp == 0) {
    double tot = 0.0;
    for (int r = 0; r < G; ++r) { MCH_RC(rc_read(rc_x0 + r, rc_epoch, (unsigned)rank, "step-0 totals")); tot += x0[r]; }
    Unb = tot * scale;
    double longest;
    U = Unb + bond_terms(-1, (T)0, (T)0, (T)0, longest);
    if (lane == 0 && rank == 0) {
      if (a.trace_f) { double* t = a.trace_f + (long long)c * a.num_steps * 3; t[0] = U; t[1] = Unb; t[2] = longest; }
      if (a.trace_i) a.trace_i[(long long)c * a.num_steps] = -1;
      if (a.num_steps <= 1) a.maxd[c] = longest;
    }
  }
  auto write_frame = [&](int step) {  // own atoms of frame `step`
    if (!a.traj) return;
    const long long frame = ((long long)c * a.num_steps + step) * N * 3;
    for (int l = tid; l < cnt; l += nt) {
      const long long dst = frame + 3LL * a.perm[lo + l];
      store_io<Tio>(a.traj, dst + 0, (double)X[l] + ctl->origin[0]);
      store_io<Tio>(a.traj, dst + 1, (double)Y[l] + ctl->origin[1]);
      store_io<Tio>(a.traj, dst + 2, (double)Z[l] + ctl->origin[2]);
    }
  };
  write_frame(0);

  // ---------------------------------------------------------------- Monte-Carlo steps
  for (int step = 1; step < a.num_steps; ++step) {
    if (warp == 0) {
      kb = (int)pcg_bounded(rng, (uint32_t)B);                       // optimizer.py:214
      const int side = (int)pcg_bounded(rng, 2u);                    // :225
      ms = bsu[2 * kb + side];
      rnd = (pcg_double(rng) - 0.5) * 2.0;                           // :229
      kind = (int)pcg_bounded(rng, 2u);                              // :242-244
      if (lane == 0) { ctl->ms = ms; ctl->n = a.row_off[ms + 1] - a.row_off[ms]; ctl->np = padded(ctl->n); }
    }
    __syncthreads();
    const int cur = ctl->ms, n = ctl->n, np = ctl->np, r_off = a.row_off[cur];
    MCH_CHECK(cur >= 0 && cur < a.S && np + 3 <= rows_cap, "moving subunit / row buffer capacity");
    for (int i = tid; i < n; i += nt) {  // current positions of the moving atoms, from their owners
      T x, y, z;
      pull(a.row_slot[r_off + i], x, y, z);
      R0[i] = x; R0[rows_cap + i] = y; R0[2 * rows_cap + i] = z;
      Rw[i] = a.row_w[r_off + i];
    }
    __syncthreads();
    if (warp == 0) {
      // subunit centroid = mean over the LISTED ids (repeats counted); molecule centroid from the CTAs' sums
      double sx = 0.0, sy = 0.0, sz = 0.0, sw = 0.0;
      for (int i = lane; i < n; i += 32) {
        const double w = (double)Rw[i];
        sx += w * (double)R0[i]; sy += w * (double)R0[rows_cap + i]; sz += w * (double)R0[2 * rows_cap + i]; sw += w;
      }
      sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz); sw = warp_sum(sw);
      double tx = 0.0, ty = 0.0, tz = 0.0;
      for (int r = 0; r < G; ++r) {
        MCH_RC(rc_read(rc_cs + r, rc_epoch, (unsigned)rank, "coordinate sums"));
        tx += cs[3 * r]; ty += cs[3 * r + 1]; tz += cs[3 * r + 2];
      }
      const int sa = 2 * kb, sb = 2 * kb + 1;
      const double bvx = (double)bpos[sb] - (double)bpos[sa];        // :215-217 (P[b1] - P[b0])
      const double bvy = (double)bpos[bstride + sb] - (double)bpos[bstride + sa];
      const double bvz = (double)bpos[2 * bstride + sb] - (double)bpos[2 * bstride + sa];
      const double cvx = sx / sw - tx / N, cvy = sy / sw - ty / N, cvz = sz / sw - tz / N;   // :236-237
      const double ss = a.p.step_size;
      double vx, vy, vz;
      if (kind == 0) { vx = ((-bvx) * ss) * rnd; vy = ((-bvy) * ss) * rnd; vz = ((-bvz) * ss) * rnd; }   // :233
      else           { vx = (cvx * ss) * rnd;    vy = (cvy * ss) * rnd;    vz = (cvz * ss) * rnd; }      // :238
      if (lane == 0) {
        ctl->tv[0] = (double)(T)vx; ctl->tv[1] = (double)(T)vy; ctl->tv[2] = (double)(T)vz;
        ctl->ctr[0] = sx / sw; ctl->ctr[1] = sy / sw; ctl->ctr[2] = sz / sw;
      }
      double unused;
      Ub_trial = bond_terms(ms, (T)vx, (T)vy, (T)vz, unused);
    }
    __syncthreads();
    const T tvx = (T)ctl->tv[0], tvy = (T)ctl->tv[1], tvz = (T)ctl->tv[2];
    const Centre<T> ctr_old{(T)ctl->ctr[0], (T)ctl->ctr[1], (T)ctl->ctr[2]};
    const Centre<T> ctr_new{ctr_old.x - tvx, ctr_old.y - tvy, ctr_old.z - tvz};
    for (int i = tid; i < np; i += nt) {
      if (i < n) {
        const T x = R0[i], y = R0[rows_cap + i], z = R0[2 * rows_cap + i];
        Qo[i] = encode_row<T, kXF>(x, y, z, ctr_old);
        Qn[i] = encode_row<T, kXF>(x - tvx, y - tvy, z - tvz, ctr_new);   // :248-252 (pos - vector)
      } else {
        Qo[i] = padding_row<T, kXF>();
        Qn[i] = padding_row<T, kXF>();
      }
    }
    __syncthreads();
    {
      double p_new = sweep_own(Qn, np, cur, ctr_new);
      double p_old = sweep_own(Qo, np, cur, ctr_old);
      double unused = 0.0;
      block_sum3(p_new, p_old, unused);
      if (warp == 0 && lane < G) {
        double* peer = cluster.map_shared_rank(xs, lane);
        MCH_RC(rc_write(cluster.map_shared_rank(rc_xs, lane) + 2 * rank, rc_epoch, (unsigned)rank, "sweep totals"));
        MCH_RC(rc_write(cluster.map_shared_rank(rc_xs, lane) + 2 * rank + 1, rc_epoch, (unsigned)rank, "sweep totals"));
        peer[2 * rank] = p_new; peer[2 * rank + 1] = p_old;
      }
    }
    cluster.sync();  // barrier 1: every CTA's sweep totals are here
    MCH_RC(++rc_epoch);
    if (warp == 0) {
      double fresh = 0.0, stale = 0.0;
      for (int r = 0; r < G; ++r) {
        MCH_RC(rc_read(rc_xs + 2 * r, rc_epoch, (unsigned)rank, "sweep totals"));
        MCH_RC(rc_read(rc_xs + 2 * r + 1, rc_epoch, (unsigned)rank, "sweep totals"));
        fresh += xs[2 * r]; stale += xs[2 * r + 1];
      }
      const double Unb_new = Unb + (fresh - stale) * scale;
      const double U_new = Unb_new + Ub_trial;
      bool ok;
      if (U_new < U) ok = true;                                                       // mc_operations.py:46-47
      else ok = boltzmann_beats(-a.p.beta * (U_new - U), pcg_double(rng));            // mc_operations.py:49-52
      pair_count += 2ULL * (unsigned long long)n * (unsigned long long)(N - n);
      if (ok) { U = U_new; Unb = Unb_new; ++n_acc; }
      // the replica of the bond ends follows the owners' arithmetic exactly
      if (ok || a.inexact_undo) {
        for (int k = lane; k < 2 * B; k += 32) {
          if (!in_rows_of(ms, bslot[k])) continue;
          T x = bpos[k] - tvx, y = bpos[bstride + k] - tvy, z = bpos[2 * bstride + k] - tvz;
          if (!ok) { x = x - (-tvx); y = y - (-tvy); z = z - (-tvz); }                // optimizer.py:273-277
          bpos[k] = x; bpos[bstride + k] = y; bpos[2 * bstride + k] = z;
        }
      }
      __syncwarp();
      info = kb | ((ok ? 1 : 0) << 16) | (kind << 17) | (ms << 18);
      if (lane == 0) ctl->ok = ok ? 1 : 0;
      if (a.trace_f || a.trace_i || step + 1 == a.num_steps) {
        double longest;
        (void)bond_terms(-1, (T)0, (T)0, (T)0, longest);
        if (lane == 0 && rank == 0) {
          if (a.trace_f) { double* t = a.trace_f + ((long long)c * a.num_steps + step) * 3; t[0] = U; t[1] = Unb; t[2] = longest; }
          if (a.trace_i) a.trace_i[(long long)c * a.num_steps + step] = info;
          if (step + 1 == a.num_steps) a.maxd[c] = longest;
        }
      }
    }
    __syncthreads();
    {
      const bool ok = ctl->ok != 0;
      bool touched = false;
      if (ok || a.inexact_undo) {
        for (int k = a.run_off[cur]; k < a.run_off[cur + 1]; ++k) {
          int b = a.runs[2 * k] - lo, e = a.runs[2 * k + 1] - lo;
          b = b < 0 ? 0 : b; e = e > cnt ? cnt : e;
          if (b >= e) continue;
          touched = true;
          for (int l = b + tid; l < e; l += nt) {
            T x = X[l] - tvx, y = Y[l] - tvy, z = Z[l] - tvz;
            if (!ok) { x = x - (-tvx); y = y - (-tvy); z = z - (-tvz); }
            MCH_RC(rc_write(rc_pos + l, rc_epoch, (unsigned)rank, "positions"));
            X[l] = x; Y[l] = y; Z[l] = z;
          }
        }
      }
      if (touched) {  // uniform over the CTA
        __syncthreads();
        publish_own_sums();
      }
      if (a.traj) { __syncthreads(); write_frame(step); }
    }
#if !(defined(MCH_RACECHECK_BREAK) && MCH_RACECHECK_BREAK == 2)  // negative control of the check build: no barrier 2
    cluster.sync();  // barrier 2: positions and coordinate sums are final before the next pull
    MCH_RC(++rc_epoch);
#endif
  }

  // ---------------------------------------------------------------- epilogue
  for (int l = tid; l < cnt; l += nt) {
    const long long dst = (long long)c * N * 3 + 3LL * a.perm[lo + l];
    store_io<Tio>(a.pos_out, dst + 0, (double)X[l] + ctl->origin[0]);
    store_io<Tio>(a.pos_out, dst + 1, (double)Y[l] + ctl->origin[1]);
    store_io<Tio>(a.pos_out, dst + 2, (double)Z[l] + ctl->origin[2]);
  }
  if (rank == 0 && warp == 0 && lane == 0) {
    a.U[c] = U; a.Unb[c] = Unb; a.n_accept[c] = n_acc; a.last_info[c] = info;
    if (a.pairs) a.pairs[c] = pair_count;
    pcg_store(rng, a.rng + 6LL * c);
  }
  cluster.sync();  // no CTA leaves while a peer may still read its shared memory
}

template <typename T, typename Tio>
int launch_wide_io(int mk, const WideArgs& a, int threads, int smem, cudaStream_t s) {
  const int blocks = a.C * a.G;
  const char* what = "mch_optimize_batch (split-state kernel)";
#ifdef MCH_QUICK
  if (mk != MU_3) return fail(-3, "MCH_QUICK build: mu = 3 only");
  return launch_clustered_wide(mch_optimize_wide_kernel<T, Tio, MU_3>, a, blocks, threads, smem, a.G, s, what);
#else
  switch (mk) {
    case MU_3: return launch_clustered_wide(mch_optimize_wide_kernel<T, Tio, MU_3>, a, blocks, threads, smem, a.G, s, what);
    case MU_INT: return launch_clustered_wide(mch_optimize_wide_kernel<T, Tio, MU_INT>, a, blocks, threads, smem, a.G, s, what);
    default: return launch_clustered_wide(mch_optimize_wide_kernel<T, Tio, MU_REAL>, a, blocks, threads, smem, a.G, s, what);
  }
#endif
}

#endif  // __CUDACC__

}  // namespace mch
