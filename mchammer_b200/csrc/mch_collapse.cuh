// mch_collapse.cuh -- the Collapser loop, one CTA per molecule, resident for all steps.
//
// Replaces Collapser.get_result / get_trajectory (reference collapser.py:200-340):
//   while min over cross-subunit non-H pairs of |r_i - r_j| >= threshold:   (:49-82, :253)
//       v_s = centroid(s) - centroid(all); scale_s = |v_s| / max_t |v_t| (or 1)   (:106-157)
//       pos[a] -= (step_size * v_s) * scale_s   for a in s                        (:84-104)
//   if min distance < threshold / 2: one more step with step_size = -(step_size/2) (:264-276)
//
// Caller's slot order (the ABI): subunits contiguous, inside each subunit the non-H atoms
// first (n_heavy[s] of them).  Inside the kernel the atoms are re-ordered once more: the
// heavy atoms of ALL subunits first (grouped by subunit), then all H atoms, so that the
// contact test works on one dense array q in [0, nh) in which the partners of column q --
// the heavy atoms of all EARLIER subunits, each cross pair once -- are simply the rows
// [0, hoff[subunit(q)]).  The triangle of (column, row) pairs is cut into blocks of
// 128 columns x 32 rows that are dealt round-robin to the warps (equal cost per block, so
// the triangle's imbalance disappears); a block entirely below every column's limit runs
// without the per-pair "row belongs to an earlier subunit" predicate.  Rows are broadcast LDS.
//
// The contact test needs, per step, only the BOOLEAN "min distance < threshold"; the exact minimum
// matters once, at the end (threshold/2 test, log).  Every pass runs in a FAST form -- expanded-form
// squared distances (4 FP instructions per pair instead of 6: rows are broadcast as one LDS.128 of
// (-2x', -2y', -2z', |p'|^2)), in fp32 on packed FFMA2/FADD2, in fp64 with the running minimum kept on
// the HIGH WORD of r^2 only (one integer min per pair instead of DSETP + two selects).  A (row block,
// column) whose fast minimum lies under threshold^2 + the fast form's error bound is a candidate and is
// re-evaluated on the spot in the plain form at full precision (the arithmetic of round 1); every pair
// under the threshold is a candidate, so the decision and the reported minimum are those of the plain
// form -- and a pass without a contact has no candidates: there is no separate exact pass any more.
#pragma once

#include "mch_device.cuh"
#include "mch_optimize.cuh"

#ifndef MCH_COL_RB
#define MCH_COL_RB 64  // rows per task of the contact pass
#endif

namespace mch {

struct ColArgs {
  const void* pos_in;       // [C][N][3]
  long long pos_in_stride;  // 0 -> broadcast chain 0
  void* pos_out;            // [C][N][3]
  int C, N, S, B;
  int S_moving;        // the first S_moving subunits are the caller's; a trailing group of atoms that are in no
                       // subunit (if any) never moves, is not contact-tested and does not enter the scales
                       // (reference collapser.py:56-82, :106-157 loop over the caller's dict only)
  const int* perm;     // [N] slot -> atom id
  const int* su_off;   // [S+1]
  const int* n_heavy;  // [S] non-H atoms per subunit (leading slots of the subunit)
  const int* bonds;    // [B][2] atom ids
  double step_size, threshold;
  int scale_steps;
  int nh_total;         // non-H atoms of the molecule (sizes the row buffer of the fast contact pass)
  int max_steps;        // safety cap (the reference loops forever if nothing ever touches)
  int* steps_run;       // [C]
  double* min_dist;     // [C] value compared with threshold/2
  double* maxd_last;    // [C] longest listed bond after the last step (0 if no step ran)
  int* status;          // [C] 0 ok, 1 = max_steps reached
  double* trace_maxd;   // [C][trace_cap] or null; entry k = after step k+1
  void* traj;           // [C][trace_cap][N][3] or null
  int trace_cap;
};

struct ColLayout { int x, y, z, rows, suof, qatom, hoff, loff, bslot, csum, vec, wsum, total; };

MCH_HD ColLayout col_layout(int N, int S, int B, int tsize, int n_heavy_total) {
  ColLayout L;
  const int np = align_up(N, 8);
  int o = 0;
  L.x = o; o += np * tsize;
  L.y = o; o += np * tsize;
  L.z = o; o += np * tsize;
  o = align_up(o, 32);
  L.rows = o; o += align_up(n_heavy_total + 4, 4) * row_bytes(tsize);  // encoded heavy atoms (fast contact pass)
  L.suof = o; o += np * 4;    // internal slot -> subunit
  L.qatom = o; o += np * 4;   // internal slot -> atom id
  L.hoff = o; o += align_up(S + 1, 4) * 4;  // heavy atoms before subunit s (internal slots of its heavy atoms)
  L.loff = o; o += align_up(S + 1, 4) * 4;  // nh + H atoms before subunit s (internal slots of its H atoms)
  L.bslot = o; o += align_up(2 * B + 2, 4) * 4;
  o = align_up(o, 16);
  L.csum = o; o += S * 3 * 8;
  L.vec = o; o += S * 3 * 8;
  L.wsum = o; o += 3 * 32 * 8;
  L.total = align_up(o, 16);
  return L;
}

#ifdef __CUDACC__

// Register budget: fp32 molecules are small in shared memory (33 KB for a 1000-atom cage), so six
// 128-thread blocks share an SM at 80 registers (no spills); fp64 fits four per SM by shared memory
// anyway and keeps 128 registers.
template <typename T, typename Tio>
__global__ void __launch_bounds__(256, (sizeof(T) == 4 ? 3 : 2))
mch_collapse_kernel(const __grid_constant__ ColArgs a) {
  extern __shared__ __align__(32) unsigned char smem[];
  const int c = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nt = blockDim.x, nwarps = nt >> 5;
  const int N = a.N, S = a.S, B = a.B;
  const ColLayout L = col_layout(N, S, B, (int)sizeof(T), a.nh_total);
  T* X = reinterpret_cast<T*>(smem + L.x);
  T* Y = reinterpret_cast<T*>(smem + L.y);
  T* Z = reinterpret_cast<T*>(smem + L.z);
  int* suof = reinterpret_cast<int*>(smem + L.suof);
  int* qatom = reinterpret_cast<int*>(smem + L.qatom);
  int* hoff = reinterpret_cast<int*>(smem + L.hoff);
  int* loff = reinterpret_cast<int*>(smem + L.loff);
  int* bslot = reinterpret_cast<int*>(smem + L.bslot);
  Row4<T>* Qh = reinterpret_cast<Row4<T>*>(smem + L.rows);
  double* csum = reinterpret_cast<double*>(smem + L.csum);
  double* vec = reinterpret_cast<double*>(smem + L.vec);
  double* wsum = reinterpret_cast<double*>(smem + L.wsum);

  // ------------------------------------------------------------ prologue
  if (tid == 0) {  // prefix sums over subunits (once per molecule)
    int h = 0, l = 0;
    for (int s = 0; s < S; ++s) {
      const int n = a.su_off[s + 1] - a.su_off[s], nh_s = a.n_heavy[s];
      hoff[s] = h; loff[s] = l;
      h += nh_s; l += n - nh_s;
    }
    hoff[S] = h; loff[S] = l;
    for (int s = 0; s <= S; ++s) loff[s] += h;
  }
  __syncthreads();
  const int Sm = a.S_moving;
  const int nh = hoff[Sm];  // heavy atoms that take part in the contact test (the caller's subunits only)
  int* inv = reinterpret_cast<int*>(X);  // atom id -> internal slot, N ints: borrows X (positions are loaded later)
  for (int sidx = warp; sidx < S; sidx += nwarps) {
    const int r0 = a.su_off[sidx], n = a.su_off[sidx + 1] - r0, nh_s = a.n_heavy[sidx];
    for (int k = lane; k < n; k += 32) {
      const int q = k < nh_s ? hoff[sidx] + k : loff[sidx] + (k - nh_s);
      const int atom = a.perm[r0 + k];
      qatom[q] = atom; suof[q] = sidx; inv[atom] = q;
    }
  }
  __syncthreads();
  for (int k = tid; k < 2 * B; k += nt) bslot[k] = inv[a.bonds[k]];
  __syncthreads();
  const long long in_base = (long long)c * a.pos_in_stride;
  double origin[3] = {0.0, 0.0, 0.0};
  if (sizeof(T) == 4) {
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int n = tid; n < N; n += nt) {
      sx += load_io<Tio>(a.pos_in, in_base + 3LL * n + 0);
      sy += load_io<Tio>(a.pos_in, in_base + 3LL * n + 1);
      sz += load_io<Tio>(a.pos_in, in_base + 3LL * n + 2);
    }
    sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
    if (lane == 0) { wsum[warp] = sx; wsum[32 + warp] = sy; wsum[64 + warp] = sz; }
    __syncthreads();
    for (int w = 0; w < nwarps; ++w) { origin[0] += wsum[w]; origin[1] += wsum[32 + w]; origin[2] += wsum[64 + w]; }
    origin[0] /= N; origin[1] /= N; origin[2] /= N;
    __syncthreads();
  }
  for (int q = tid; q < N; q += nt) {
    const long long src = in_base + 3LL * qatom[q];
    X[q] = (T)(load_io<Tio>(a.pos_in, src + 0) - origin[0]);
    Y[q] = (T)(load_io<Tio>(a.pos_in, src + 1) - origin[1]);
    Z[q] = (T)(load_io<Tio>(a.pos_in, src + 2) - origin[2]);
  }
  __syncthreads();

  constexpr int CT = 128, RB = MCH_COL_RB;  // block = 4 column units x RB rows
  constexpr bool kFast = sizeof(T) == 4;
  // centre of the expanded encoding in fp64 mode: the start centroid of the contact-tested atoms
  double cen_x = 0.0, cen_y = 0.0, cen_z = 0.0;
  if (!kFast) {
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int q = tid; q < nh; q += nt) { sx += (double)X[q]; sy += (double)Y[q]; sz += (double)Z[q]; }
    sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
    if (lane == 0) { wsum[warp] = sx; wsum[32 + warp] = sy; wsum[64 + warp] = sz; }
    __syncthreads();
    for (int w = 0; w < nwarps; ++w) { cen_x += wsum[w]; cen_y += wsum[32 + w]; cen_z += wsum[64 + w]; }
    const double inv_nh = nh > 0 ? 1.0 / nh : 0.0;
    cen_x *= inv_nh; cen_y *= inv_nh; cen_z *= inv_nh;
    __syncthreads();
  }
  // min over cross-subunit heavy pairs of r^2, as a distance (sqrt is monotone, so
  // min(dist) == sqrt(min r2) exactly)
  auto min_cross_distance = [&]() -> double {
    T best = far_away<T>();
    const int n_ct = (nh + CT - 1) / CT;
    int task = 0;
    for (int ct = 0; ct < n_ct; ++ct) {
      const int q0 = ct * CT;
      const int q_last = min(q0 + CT, nh) - 1;
      const int lim_lo = hoff[suof[q0]], lim_hi = hoff[suof[q_last]];  // rows of the first / last column
      const int n_rb = (lim_hi + RB - 1) / RB;
      // this warp's blocks of the column tile: task, task + nwarps, ...
      int rb = (warp - task % nwarps + nwarps) % nwarps;
      task += n_rb;
      if (rb >= n_rb) continue;
      T xj[4], yj[4], zj[4], m[4];
      int lim[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int q = q0 + 32 * k + lane;
        const bool live = q < nh;
        lim[k] = live ? hoff[suof[q]] : 0;
        xj[k] = live ? X[q] : far_away<T>();
        yj[k] = live ? Y[q] : far_away<T>();
        zj[k] = live ? Z[q] : far_away<T>();
        m[k] = far_away<T>();
      }
      for (; rb < n_rb; rb += nwarps) {
        const int i0 = rb * RB, i1 = min(i0 + RB, lim_hi);
        if (i1 <= lim_lo) {  // every row is a partner of every column of the tile
#pragma unroll 4
          for (int i = i0; i < i1; ++i) {
            const T xi = X[i], yi = Y[i], zi = Z[i];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const T dx = xj[k] - xi, dy = yj[k] - yi, dz = zj[k] - zi;
              m[k] = min(m[k], fma(dz, dz, fma(dy, dy, dx * dx)));
            }
          }
        } else {
          for (int i = i0; i < i1; ++i) {
            const T xi = X[i], yi = Y[i], zi = Z[i];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const T dx = xj[k] - xi, dy = yj[k] - yi, dz = zj[k] - zi;
              const T r2 = fma(dz, dz, fma(dy, dy, dx * dx));
              m[k] = (i < lim[k] && r2 < m[k]) ? r2 : m[k];
            }
          }
        }
      }
      best = min(best, min(min(m[0], m[1]), min(m[2], m[3])));
    }
    best = warp_min(best);
    __syncthreads();
    if (lane == 0) wsum[warp] = (double)best;
    __syncthreads();
    double r2 = wsum[0];
    for (int w = 1; w < nwarps; ++w) r2 = fmin(r2, wsum[w]);
    return sqrt(r2);
  };

  // ---- contact pass ------------------------------------------------------------------------------
  // Encoded rows Qh[q] of the contact-tested heavy atoms relative to `cen` (fp32: the frame is centred
  // already; fp64: the start centroid).  Rebuilt by encode_heavy() after every change of positions.
  const Centre<T> cen{(T)(kFast ? 0.0 : cen_x), (T)(kFast ? 0.0 : cen_y), (T)(kFast ? 0.0 : cen_z)};
  // rmax2 = largest |p'|^2: the rounding error of an expanded-form r^2 is bounded by a few ulps of
  // (|p_i'| + |p_j'|)^2 <= 4 rmax2, which sets the width of the candidate band below.
  // The per-pair work of the pass is r^2 - |p_j'|^2 = fma(z_j', qz, fma(y_j', qy, fma(x_j', qx, qw))) -- the
  // column's own |p_j'|^2 is the same for every row, so it is taken out of the minimum (3 FP instructions per
  // pair; it comes back in the per-column candidate threshold).  fp64 keeps its minimum on the HIGH WORD,
  // which orders positive doubles only: there the rows carry qw = |p_i'|^2 + bias with bias = rmax2 >= every
  // |p_j'|^2, so every value stays positive.
  double rmax2 = 0.0;
  auto encode_heavy = [&]() {
    T big = (T)0;
    for (int q = tid; q < nh; q += nt) {
      const Row4<T> r = encode_row<T, true>(X[q], Y[q], Z[q], cen);
      Qh[q] = r;
      big = r.w > big ? r.w : big;
    }
    const double bw = warp_max((double)big);
    if (lane == 0) wsum[32 + warp] = bw;  // (slots [32, 64): last read before the previous barrier pair)
    __syncthreads();
    double b = wsum[32];
    for (int w = 1; w < nwarps; ++w) b = fmax(b, wsum[32 + w]);
    rmax2 = b;
    if (!kFast) {  // every thread biases the rows it wrote itself
      for (int q = tid; q < nh; q += nt) Qh[q].w += (T)b;
      __syncthreads();
    }
  };
  // One pass over all cross-subunit heavy pairs in the fast form.  Every (task, column) whose fast minimum
  // lies under `cut` = threshold^2 + the fast form's error bound is a CANDIDATE: its rows are evaluated once
  // more in the plain form (round 1's arithmetic, full precision) and folded into the exact minimum `em`.
  // A pair whose exact r^2 is under threshold^2 is always a candidate, so
  //   * "min distance < threshold"  ==  sqrt(em) < threshold, the decision of the plain-form pass, and
  //   * when that holds, em IS the exact global minimum (reported, tested against threshold / 2).
  // Without a contact there are no candidates (or a handful inside the error band) and the pass costs
  // nothing beyond the fast form: no second pass, no running global minimum.
  auto contact_pass = [&]() -> double {
    const double t2 = a.threshold * a.threshold;
    // fp32 expanded form: |error| <= 8 roundings of magnitude <= 4 rmax2 (2^-24 each) -> 2e-6 rmax2;
    // fp64: the same bound with 2^-53 roundings, plus the high-word granularity (2^-20 relative)
    const double cut = kFast ? t2 * (1.0 + 2.0e-6) + 2.0e-6 * rmax2 : t2 * (1.0 + 2.0e-6) + 1.0e-14 * rmax2;
    const double bias = kFast ? 0.0 : rmax2;  // see encode_heavy
    T em = far_away<T>() * far_away<T>();
    const int n_ct = (nh + CT - 1) / CT;
    int task0 = 0;
    for (int ct = 0; ct < n_ct; ++ct) {
      // Partners of the tile's columns: rows [0, lim_lo) for every column (blocks of RB rows, no
      // predicate), and the heavy atoms of the subunits s_lo .. s_hi-1, each of which is a partner of
      // exactly the columns that belong to a LATER subunit -- one task per such subunit: its rows are
      // swept without a predicate; which columns they count for is decided per task afterwards.
      const int q0 = ct * CT;
      const int q_last = min(q0 + CT, nh) - 1;
      const int s_lo = suof[q0], s_hi = suof[q_last];
      const int lim_lo = hoff[s_lo];
      const int n_rb = (lim_lo + RB - 1) / RB;
      const int n_task = n_rb + (s_hi - s_lo);
      int task = (warp - task0 % nwarps + nwarps) % nwarps;  // this warp's tasks: task, task + nwarps, ...
      task0 += n_task;
      if (task >= n_task) continue;
      int su[4];
      // the plain-form re-evaluation of one candidate (task rows [i0, i1), this lane's column k)
      auto refine = [&](int k, int i0, int i1) {
        const int q = q0 + 32 * k + lane;
        const T xq = X[q], yq = Y[q], zq = Z[q];
        for (int i = i0; i < i1; ++i) {
          const T dx = xq - X[i], dy = yq - Y[i], dz = zq - Z[i];
          em = min(em, fma(dz, dz, fma(dy, dy, dx * dx)));
        }
      };
      if constexpr (kFast) {
        f32x2 xj[2], yj[2], zj[2];
        float cutk[4];
#pragma unroll
        for (int kp = 0; kp < 2; ++kp) {
          float x[2], y[2], z[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int q = q0 + 32 * (2 * kp + h) + lane;
            const bool live = q < nh;
            su[2 * kp + h] = live ? suof[q] : -1;
            x[h] = live ? X[q] : 1.0e18f; y[h] = live ? Y[q] : 1.0e18f; z[h] = live ? Z[q] : 1.0e18f;
            // candidate threshold of this column: r^2 - |p_j'|^2 <= cut - |p_j'|^2 (rounded up)
            cutk[2 * kp + h] = ((float)cut - fmaf(z[h], z[h], fmaf(y[h], y[h], x[h] * x[h]))) + 1.0e-6f * (float)cut;
          }
          xj[kp] = pack2(x[0], x[1]); yj[kp] = pack2(y[0], y[1]); zj[kp] = pack2(z[0], z[1]);
        }
        for (; task < n_task; task += nwarps) {
          const bool edge = task >= n_rb;
          const int u = s_lo + (task - n_rb);
          const int i0 = edge ? hoff[u] : task * RB;
          const int i1 = edge ? hoff[u + 1] : min(i0 + RB, lim_lo);
          float mt[4] = {3.0e38f, 3.0e38f, 3.0e38f, 3.0e38f};
#pragma unroll 4
          for (int i = i0; i < i1; ++i) {
            const Row4<float> q = Qh[i];
#pragma unroll
            for (int kp = 0; kp < 2; ++kp) {
              const f32x2 r2 = fma2(zj[kp], pack2(q.z, q.z), fma2(yj[kp], pack2(q.y, q.y),
                               fma2(xj[kp], pack2(q.x, q.x), pack2(q.w, q.w))));
              float a0, a1;
              unpack2(r2, a0, a1);
              mt[2 * kp] = fminf(mt[2 * kp], a0);
              mt[2 * kp + 1] = fminf(mt[2 * kp + 1], a1);
            }
          }
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (mt[k] <= cutk[k] && (edge ? su[k] > u : su[k] >= 0)) refine(k, i0, i1);
        }
      } else {
        double xj[4], yj[4], zj[4];
        int cutk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int q = q0 + 32 * k + lane;
          const bool live = q < nh;
          su[k] = live ? suof[q] : -1;
          xj[k] = live ? (double)X[q] - (double)cen.x : 0.0;  // (dead columns: never candidates, su = -1)
          yj[k] = live ? (double)Y[q] - (double)cen.y : 0.0;
          zj[k] = live ? (double)Z[q] - (double)cen.z : 0.0;
          // candidate threshold of this column, on the high word: r^2 - |p_j'|^2 + bias <= cut - |p_j'|^2 + bias
          cutk[k] = __double2hiint((cut - fma(zj[k], zj[k], fma(yj[k], yj[k], xj[k] * xj[k]))) + bias);
        }
        for (; task < n_task; task += nwarps) {
          const bool edge = task >= n_rb;
          const int u = s_lo + (task - n_rb);
          const int i0 = edge ? hoff[u] : task * RB;
          const int i1 = edge ? hoff[u + 1] : min(i0 + RB, lim_lo);
          constexpr int kHuge = 0x7fe00000;  // high word of a huge finite double
          int mt[4] = {kHuge, kHuge, kHuge, kHuge};
#pragma unroll 2
          for (int i = i0; i < i1; ++i) {
            const Row4<double> q = reinterpret_cast<const Row4<double>*>(Qh)[i];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const double v = fma(zj[k], q.z, fma(yj[k], q.y, fma(xj[k], q.x, q.w)));  // = r^2 - |p_j'|^2 + bias > 0
              mt[k] = min(mt[k], __double2hiint(v));  // positive doubles order like their high words
            }
          }
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (mt[k] <= cutk[k] && (edge ? su[k] > u : su[k] >= 0)) refine(k, i0, i1);
        }
      }
    }
    em = warp_min(em);
    if (lane == 0) wsum[64 + warp] = (double)em;  // (slots [64, 96): last read before the barriers of the previous step)
    __syncthreads();
    double v = wsum[64];
    for (int w = 1; w < nwarps; ++w) v = fmin(v, wsum[64 + w]);
    return v;  // exact minimum r^2 over the candidates (huge if there is none)
  };
  // "min distance < threshold"; when true, dmin is the exact minimum distance (dmin < 0: not computed)
  auto contact = [&](double& dmin) -> bool {
    const double d = sqrt(contact_pass());
    const bool touching = d < a.threshold;
    dmin = touching ? d : -1.0;
    return touching;
  };

  // one collapse step (collapser.py:159-180) with the given signed step size
  auto do_step = [&](double step) {
    for (int s = warp; s < S; s += nwarps) {
      double sx = 0.0, sy = 0.0, sz = 0.0;
      for (int n = hoff[s] + lane; n < hoff[s + 1]; n += 32) { sx += (double)X[n]; sy += (double)Y[n]; sz += (double)Z[n]; }
      for (int n = loff[s] + lane; n < loff[s + 1]; n += 32) { sx += (double)X[n]; sy += (double)Y[n]; sz += (double)Z[n]; }
      sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
      if (lane == 0) { csum[3 * s] = sx; csum[3 * s + 1] = sy; csum[3 * s + 2] = sz; }
    }
    __syncthreads();
    {
      // Every warp forms the step vectors itself -- the same inputs, the same arithmetic, the same bits -- and
      // writes them all (identical values from every warp: no warp has to wait for another one here).
      double tx = 0.0, ty = 0.0, tz = 0.0;
      for (int s = lane; s < S; s += 32) { tx += csum[3 * s]; ty += csum[3 * s + 1]; tz += csum[3 * s + 2]; }
      tx = warp_sum(tx) / N; ty = warp_sum(ty) / N; tz = warp_sum(tz) / N;
      auto centroid_vector = [&](int s, double& vx, double& vy, double& vz) {
        const double n = (double)((hoff[s + 1] - hoff[s]) + (loff[s + 1] - loff[s]));
        vx = csum[3 * s] / n - tx; vy = csum[3 * s + 1] / n - ty; vz = csum[3 * s + 2] / n - tz;
      };
      double big = 0.0;
      for (int s = lane; s < Sm; s += 32) {
        double vx, vy, vz;
        centroid_vector(s, vx, vy, vz);
        big = fmax(big, sqrt(vx * vx + vy * vy + vz * vz));
      }
      big = warp_max(big);
      for (int s = lane; s < S; s += 32) {
        double vx = 0.0, vy = 0.0, vz = 0.0, sc = 1.0;
        if (s < Sm) {  // the static group (atoms in no subunit) keeps a zero vector
          centroid_vector(s, vx, vy, vz);
          if (a.scale_steps) sc = sqrt(vx * vx + vy * vy + vz * vz) / big;
        }
        vec[3 * s] = (step * vx) * sc; vec[3 * s + 1] = (step * vy) * sc; vec[3 * s + 2] = (step * vz) * sc;
      }
      __syncwarp();
    }
    for (int n = tid; n < N; n += nt) {
      const int s = suof[n];
      if (s >= Sm) continue;  // atoms in no subunit keep their positions bit for bit
      X[n] = (T)((double)X[n] - vec[3 * s]);
      Y[n] = (T)((double)Y[n] - vec[3 * s + 1]);
      Z[n] = (T)((double)Z[n] - vec[3 * s + 2]);
    }
    __syncthreads();
    encode_heavy();
  };

  auto record = [&](int step) {  // step is 1-based; longest listed bond + optional frame
    double far = 0.0;
    if (warp == 0) {
      for (int b = lane; b < B; b += 32) {
        const int sa = bslot[2 * b], sb = bslot[2 * b + 1];
        const double ex = (double)X[sa] - (double)X[sb], ey = (double)Y[sa] - (double)Y[sb], ez = (double)Z[sa] - (double)Z[sb];
        far = fmax(far, sqrt(ex * ex + ey * ey + ez * ez));
      }
      far = warp_max(far);
      if (lane == 0) {
        a.maxd_last[c] = far;
        if (a.trace_maxd && step <= a.trace_cap) a.trace_maxd[(long long)c * a.trace_cap + (step - 1)] = far;
      }
    }
    if (a.traj && step <= a.trace_cap) {
      const long long frame = ((long long)c * a.trace_cap + (step - 1)) * N * 3;
      for (int n = tid; n < N; n += nt) {
        const long long dst = frame + 3LL * qatom[n];
        store_io<Tio>(a.traj, dst + 0, (double)X[n] + origin[0]);
        store_io<Tio>(a.traj, dst + 1, (double)Y[n] + origin[1]);
        store_io<Tio>(a.traj, dst + 2, (double)Z[n] + origin[2]);
      }
    }
  };

  // ------------------------------------------------------------ the loop
  int step = 0, status = 0;
  if (tid == 0) a.maxd_last[c] = 0.0;
  encode_heavy();
  double dmin;
  bool touching = contact(dmin);
  while (!touching) {
    if (step >= a.max_steps) { status = 1; break; }
    ++step;
    do_step(a.step_size);
    record(step);
    touching = contact(dmin);
  }
  if (dmin < 0.0) dmin = min_cross_distance();  // the exact minimum is reported and tested against threshold / 2
  if (status == 0 && dmin < a.threshold / 2) {
    ++step;
    do_step(-(a.step_size / 2));
    record(step);
  }
  __syncthreads();
  for (int n = tid; n < N; n += nt) {
    const long long dst = (long long)c * N * 3 + 3LL * qatom[n];
    store_io<Tio>(a.pos_out, dst + 0, (double)X[n] + origin[0]);
    store_io<Tio>(a.pos_out, dst + 1, (double)Y[n] + origin[1]);
    store_io<Tio>(a.pos_out, dst + 2, (double)Z[n] + origin[2]);
  }
  if (tid == 0) { a.steps_run[c] = step; a.min_dist[c] = dmin; a.status[c] = status; }
}

#endif  // __CUDACC__

}  // namespace mch
