// mch_potential.cuh -- standalone potential of a batch of molecules (no subunit knowledge).
//
// Replaces Optimizer.compute_potential / _compute_nonbonded_potential
// (reference optimizer.py:114-142): U_nb = sum over ALL atom pairs of eps (sigma/r)^mu,
// U = U_nb + sum_b eps_b (r_b - R_t)^2, plus the longest listed bond
// (optimizer.py:174-183).  One CTA per molecule; atoms staged once into shared memory;
// the N(N-1)/2 pairs are covered by row blocks of ROWB atoms swept against all later
// atoms (cross part) and against themselves (strict lower triangle).  Every WARP takes whole row
// blocks (its own row buffer, blocks dealt boustrophedon so the shrinking column ranges balance)
// and sweeps them alone: no block-wide barrier until the final reduction (the first version
// gave every block to the whole CTA and spent 36 % of its warp time at the two barriers per block).
// Both parts go through
// pair_sweep.  The long-bond harmonic term is fused into the same launch.
#pragma once

#include "mch_device.cuh"
#include "mch_optimize.cuh"  // align_up, load_io

#ifndef MCH_POT_X2
#define MCH_POT_X2 1  // fp32: the cross sweeps on the packed plain-form tile
#endif

namespace mch {

struct PotArgs {
  const void* pos;  // [C][N][3] io dtype
  int C, N, B;
  const int* bonds;  // [B][2] atom ids
  Params p;
  int mu_int;
  double* U;     // [C]
  double* Unb;   // [C]
  double* maxd;  // [C] (0 when B == 0)
};

constexpr int POT_ROWB = 64;

struct PotLayout { int x, y, z, rows, wsum, total; };

MCH_HD PotLayout pot_layout(int N, int tsize) {
  PotLayout L;
  const int np = align_up(N, 8);
  int o = 0;
  L.x = o; o += np * tsize;
  L.y = o; o += np * tsize;
  L.z = o; o += np * tsize;
  o = align_up(o, 32);
  L.rows = o; o += 8 * (POT_ROWB + 2 + kRowSlack) * 4 * tsize;  // one buffer per warp (+ padding rows + prefetch overrun)
  L.wsum = o; o += 3 * 32 * 8;
  L.total = align_up(o, 16);
  return L;
}

#ifdef __CUDACC__

template <typename T, typename Tio, int MUK>
__global__ void __launch_bounds__(256)
mch_potential_kernel(const __grid_constant__ PotArgs a) {
  extern __shared__ __align__(32) unsigned char smem[];
  const int c = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nt = blockDim.x, nwarps = nt >> 5;
  const int N = a.N;
  const PotLayout L = pot_layout(N, (int)sizeof(T));
  T* X = reinterpret_cast<T*>(smem + L.x);
  T* Y = reinterpret_cast<T*>(smem + L.y);
  T* Z = reinterpret_cast<T*>(smem + L.z);
  Row4<T>* Q = reinterpret_cast<Row4<T>*>(smem + L.rows);
  double* wsum = reinterpret_cast<double*>(smem + L.wsum);

  const PairTerm<T, MUK> term(a.p.nonbond_mu, a.mu_int);
  const double scale = a.p.nonbond_epsilon * pow(a.p.nonbond_sigma, a.p.nonbond_mu);
  const long long base = (long long)c * N * 3;

  double origin[3] = {0.0, 0.0, 0.0};
  if (sizeof(T) == 4) {  // centred frame for fp32 (pair distances are translation invariant)
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int n = tid; n < N; n += nt) {
      sx += load_io<Tio>(a.pos, base + 3LL * n + 0);
      sy += load_io<Tio>(a.pos, base + 3LL * n + 1);
      sz += load_io<Tio>(a.pos, base + 3LL * n + 2);
    }
    sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
    if (lane == 0) { wsum[warp] = sx; wsum[32 + warp] = sy; wsum[64 + warp] = sz; }
    __syncthreads();
    for (int w = 0; w < nwarps; ++w) { origin[0] += wsum[w]; origin[1] += wsum[32 + w]; origin[2] += wsum[64 + w]; }
    origin[0] /= N; origin[1] /= N; origin[2] /= N;
    __syncthreads();
  }
  for (int n = tid; n < N; n += nt) {
    X[n] = (T)(load_io<Tio>(a.pos, base + 3LL * n + 0) - origin[0]);
    Y[n] = (T)(load_io<Tio>(a.pos, base + 3LL * n + 1) - origin[1]);
    Z[n] = (T)(load_io<Tio>(a.pos, base + 3LL * n + 2) - origin[2]);
  }
  __syncthreads();

  // Expanded form in fp64 only: without subunit knowledge a row block is 64 arbitrary atoms, so its
  // centroid can be a molecule radius away from both atoms of a close pair; the rounding noise
  // eps (|p_i'| + |p_j'|)^2 / r^2 is harmless in fp64 (~1e-13) but not in fp32 (~1e-4 on the closest pairs).
  constexpr bool kFast = sizeof(T) == 4;
  constexpr bool kXF = !kFast && (MCH_XF64 != 0);
  constexpr bool kX2 = kFast && (MCH_F32X2 != 0) && (MCH_POT_X2 != 0);  // packed instructions, plain form (sweep_tile_x2<PLAIN>)
  constexpr int kRot3 = kX2 ? 2 : 0;
  double mine = 0.0;
  Row4<T>* Qw = Q + warp * (POT_ROWB + 2 + kRowSlack);  // this warp's row buffer
  const int nblocks = (N + POT_ROWB - 1) / POT_ROWB;
  for (int round = 0; round * nwarps < nblocks; ++round) {
    // boustrophedon: round 0 deals blocks 0..W-1 to warps 0..W-1, round 1 deals W..2W-1 to warps W-1..0, ...
    const int blk = round * nwarps + ((round & 1) ? nwarps - 1 - warp : warp);
    if (blk >= nblocks) continue;
    const int r0 = blk * POT_ROWB;
    const int nrows = (N - r0) < POT_ROWB ? (N - r0) : POT_ROWB;
    const int nrows_padded = kRot3 ? padded_rows3(nrows) : padded_rows(nrows);
    Centre<T> ctr{(T)0, (T)0, (T)0};
    if (kXF) {  // centre of the expanded encoding: the block's centroid
      double sx = 0.0, sy = 0.0, sz = 0.0;
      for (int i = lane; i < nrows; i += 32) { sx += (double)X[r0 + i]; sy += (double)Y[r0 + i]; sz += (double)Z[r0 + i]; }
      ctr.x = (T)(warp_sum(sx) / nrows); ctr.y = (T)(warp_sum(sy) / nrows); ctr.z = (T)(warp_sum(sz) / nrows);
    }
    __syncwarp();  // the previous block's sweeps of this warp are done with Qw
    for (int i = lane; i < nrows_padded; i += 32)
      Qw[i] = i < nrows ? encode_row<T, kXF>(X[r0 + i], Y[r0 + i], Z[r0 + i], ctr) : padding_row<T, kXF>();
    __syncwarp();
    const int r1 = r0 + nrows;
    Cols later{r1, r1, r1, N};
    mine += pair_sweep<T, MUK, false, false, kXF, kX2, kRot3>(Qw, nrows_padded, X, Y, Z, (T*)nullptr, later, term, ctr, 0, 0, 1, true);
    Cols own{r0, r1, r1, r1};
    mine += pair_sweep<T, MUK, true, false, kXF, kX2>(Qw, nrows, X, Y, Z, (T*)nullptr, own, term, ctr, 0, 0, 1, true);
  }
  mine = warp_sum(mine);
  if (lane == 0) wsum[warp] = mine;
  __syncthreads();
  if (warp == 0) {
    double unb = 0.0;
    for (int w = 0; w < nwarps; ++w) unb += wsum[w];
    unb *= scale;
    double e = 0.0, far = 0.0;
    for (int b = lane; b < a.B; b += 32) {
      const int sa = a.bonds[2 * b], sb = a.bonds[2 * b + 1];
      const double ex = (double)X[sa] - (double)X[sb], ey = (double)Y[sa] - (double)Y[sb], ez = (double)Z[sa] - (double)Z[sb];
      const double r = sqrt(ex * ex + ey * ey + ez * ez);
      const double d = r - a.p.target_bond_length;
      e += a.p.bond_epsilon * (d * d);
      far = fmax(far, r);
    }
    e = warp_sum(e); far = warp_max(far);
    if (lane == 0) { a.Unb[c] = unb; a.U[c] = unb + e; a.maxd[c] = far; }
  }
}

#endif  // __CUDACC__

}  // namespace mch
