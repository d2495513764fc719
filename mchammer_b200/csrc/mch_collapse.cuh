// mch_collapse.cuh -- the Collapser loop, one CTA per molecule, resident for all steps.
//
// Replaces Collapser.get_result / get_trajectory (reference collapser.py:200-340):
//   while min over cross-subunit non-H pairs of |r_i - r_j| >= threshold:   (:49-82, :253)
//       v_s = centroid(s) - centroid(all); scale_s = |v_s| / max_t |v_t| (or 1)   (:106-157)
//       pos[a] -= (step_size * v_s) * scale_s   for a in s                        (:84-104)
//   if min distance < threshold / 2: one more step with step_size = -(step_size/2) (:264-276)
//
// Slot order: subunits contiguous, and inside each subunit the non-H atoms first, so the
// heavy atoms of subunit s are the slot range [su_off[s], su_off[s] + n_heavy[s]).
// The min-distance sweep gives every thread a heavy column atom j and walks the heavy
// atoms of all EARLIER subunits (each cross pair once); rows are broadcast LDS.
#pragma once

#include "mch_device.cuh"
#include "mch_optimize.cuh"

namespace mch {

struct ColArgs {
  const void* pos_in;       // [C][N][3]
  long long pos_in_stride;  // 0 -> broadcast chain 0
  void* pos_out;            // [C][N][3]
  int C, N, S, B;
  const int* perm;     // [N] slot -> atom id
  const int* su_off;   // [S+1]
  const int* n_heavy;  // [S] non-H atoms per subunit (leading slots of the subunit)
  const int* bonds;    // [B][2] atom ids
  double step_size, threshold;
  int scale_steps;
  int max_steps;        // safety cap (the reference loops forever if nothing ever touches)
  int* steps_run;       // [C]
  double* min_dist;     // [C] value compared with threshold/2
  double* maxd_last;    // [C] longest listed bond after the last step (0 if no step ran)
  int* status;          // [C] 0 ok, 1 = max_steps reached
  double* trace_maxd;   // [C][trace_cap] or null; entry k = after step k+1
  void* traj;           // [C][trace_cap][N][3] or null
  int trace_cap;
};

struct ColLayout { int x, y, z, suof, suoff, nheavy, bslot, csum, vec, wsum, total; };

MCH_HD ColLayout col_layout(int N, int S, int B, int tsize) {
  ColLayout L;
  const int np = align_up(N, 8);
  int o = 0;
  L.x = o; o += np * tsize;
  L.y = o; o += np * tsize;
  L.z = o; o += np * tsize;
  L.suof = o; o += np * 4;
  L.suoff = o; o += align_up(S + 1, 4) * 4;
  L.nheavy = o; o += align_up(S, 4) * 4;
  L.bslot = o; o += align_up(2 * B + 2, 4) * 4;
  o = align_up(o, 16);
  L.csum = o; o += S * 3 * 8;
  L.vec = o; o += S * 3 * 8;
  L.wsum = o; o += 3 * 32 * 8;
  L.total = align_up(o, 16);
  return L;
}

#ifdef __CUDACC__

template <typename T, typename Tio>
__global__ void __launch_bounds__(256)
mch_collapse_kernel(const __grid_constant__ ColArgs a) {
  extern __shared__ __align__(32) unsigned char smem[];
  const int c = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nt = blockDim.x, nwarps = nt >> 5;
  const int N = a.N, S = a.S, B = a.B;
  const ColLayout L = col_layout(N, S, B, (int)sizeof(T));
  T* X = reinterpret_cast<T*>(smem + L.x);
  T* Y = reinterpret_cast<T*>(smem + L.y);
  T* Z = reinterpret_cast<T*>(smem + L.z);
  int* suof = reinterpret_cast<int*>(smem + L.suof);
  int* suoff = reinterpret_cast<int*>(smem + L.suoff);
  int* nheavy = reinterpret_cast<int*>(smem + L.nheavy);
  int* bslot = reinterpret_cast<int*>(smem + L.bslot);
  double* csum = reinterpret_cast<double*>(smem + L.csum);
  double* vec = reinterpret_cast<double*>(smem + L.vec);
  double* wsum = reinterpret_cast<double*>(smem + L.wsum);

  // ------------------------------------------------------------ prologue
  for (int n = tid; n < N; n += nt) suof[a.perm[n]] = n;  // inverse permutation, temporarily
  for (int s = tid; s <= S; s += nt) suoff[s] = a.su_off[s];
  for (int s = tid; s < S; s += nt) nheavy[s] = a.n_heavy[s];
  __syncthreads();
  for (int k = tid; k < 2 * B; k += nt) bslot[k] = suof[a.bonds[k]];
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
  for (int n = tid; n < N; n += nt) {
    const long long src = in_base + 3LL * a.perm[n];
    X[n] = (T)(load_io<Tio>(a.pos_in, src + 0) - origin[0]);
    Y[n] = (T)(load_io<Tio>(a.pos_in, src + 1) - origin[1]);
    Z[n] = (T)(load_io<Tio>(a.pos_in, src + 2) - origin[2]);
  }
  for (int s = warp; s < S; s += nwarps)
    for (int n = suoff[s] + lane; n < suoff[s + 1]; n += 32) suof[n] = s;
  __syncthreads();

  // min over cross-subunit heavy pairs of r^2, as a distance (sqrt is monotone, so
  // min(dist) == sqrt(min r2) exactly)
  auto min_cross_distance = [&]() -> double {
    T best = far_away<T>();
    for (int base = 0; base < N; base += 4 * nt) {
      T xj[4], yj[4], zj[4], m[4];
      int lim[4];
      int warp_lim = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int j = base + k * nt + tid;
        bool live = j < N;
        int s = live ? suof[j] : 0;
        live = live && (j - suoff[s]) < nheavy[s];
        lim[k] = live ? suoff[s] : 0;  // rows: all slots of earlier subunits
        xj[k] = live ? X[j] : far_away<T>();
        yj[k] = live ? Y[j] : far_away<T>();
        zj[k] = live ? Z[j] : far_away<T>();
        m[k] = far_away<T>();
        warp_lim = lim[k] > warp_lim ? lim[k] : warp_lim;
      }
      warp_lim = __reduce_max_sync(0xffffffffu, warp_lim);
      for (int s = 0; s < S; ++s) {
        const int r0 = suoff[s];
        if (r0 >= warp_lim) break;
        const int r1 = r0 + nheavy[s];
        for (int i = r0; i < r1; ++i) {
          const T xi = X[i], yi = Y[i], zi = Z[i];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const T dx = xj[k] - xi, dy = yj[k] - yi, dz = zj[k] - zi;
            const T r2 = fma(dz, dz, fma(dy, dy, dx * dx));
            m[k] = (i < lim[k] && r2 < m[k]) ? r2 : m[k];
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

  // one collapse step (collapser.py:159-180) with the given signed step size
  auto do_step = [&](double step) {
    for (int s = warp; s < S; s += nwarps) {
      double sx = 0.0, sy = 0.0, sz = 0.0;
      for (int n = suoff[s] + lane; n < suoff[s + 1]; n += 32) { sx += (double)X[n]; sy += (double)Y[n]; sz += (double)Z[n]; }
      sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
      if (lane == 0) { csum[3 * s] = sx; csum[3 * s + 1] = sy; csum[3 * s + 2] = sz; }
    }
    __syncthreads();
    if (warp == 0) {
      double tx = 0.0, ty = 0.0, tz = 0.0;
      for (int s = lane; s < S; s += 32) { tx += csum[3 * s]; ty += csum[3 * s + 1]; tz += csum[3 * s + 2]; }
      tx = warp_sum(tx) / N; ty = warp_sum(ty) / N; tz = warp_sum(tz) / N;
      double big = 0.0;
      for (int s = lane; s < S; s += 32) {
        const double n = (double)(suoff[s + 1] - suoff[s]);
        const double vx = csum[3 * s] / n - tx, vy = csum[3 * s + 1] / n - ty, vz = csum[3 * s + 2] / n - tz;
        vec[3 * s] = vx; vec[3 * s + 1] = vy; vec[3 * s + 2] = vz;
        big = fmax(big, sqrt(vx * vx + vy * vy + vz * vz));
      }
      big = warp_max(big);
      __syncwarp();
      for (int s = lane; s < S; s += 32) {
        const double vx = vec[3 * s], vy = vec[3 * s + 1], vz = vec[3 * s + 2];
        const double sc = a.scale_steps ? sqrt(vx * vx + vy * vy + vz * vz) / big : 1.0;
        vec[3 * s] = (step * vx) * sc; vec[3 * s + 1] = (step * vy) * sc; vec[3 * s + 2] = (step * vz) * sc;
      }
    }
    __syncthreads();
    for (int n = tid; n < N; n += nt) {
      const int s = suof[n];
      X[n] = (T)((double)X[n] - vec[3 * s]);
      Y[n] = (T)((double)Y[n] - vec[3 * s + 1]);
      Z[n] = (T)((double)Z[n] - vec[3 * s + 2]);
    }
    __syncthreads();
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
        const long long dst = frame + 3LL * a.perm[n];
        store_io<Tio>(a.traj, dst + 0, (double)X[n] + origin[0]);
        store_io<Tio>(a.traj, dst + 1, (double)Y[n] + origin[1]);
        store_io<Tio>(a.traj, dst + 2, (double)Z[n] + origin[2]);
      }
    }
  };

  // ------------------------------------------------------------ the loop
  int step = 0, status = 0;
  if (tid == 0) a.maxd_last[c] = 0.0;
  double dmin = min_cross_distance();
  while (!(dmin < a.threshold)) {
    if (step >= a.max_steps) { status = 1; break; }
    ++step;
    do_step(a.step_size);
    record(step);
    dmin = min_cross_distance();
  }
  if (status == 0 && dmin < a.threshold / 2) {
    ++step;
    do_step(-(a.step_size / 2));
    record(step);
  }
  __syncthreads();
  for (int n = tid; n < N; n += nt) {
    const long long dst = (long long)c * N * 3 + 3LL * a.perm[n];
    store_io<Tio>(a.pos_out, dst + 0, (double)X[n] + origin[0]);
    store_io<Tio>(a.pos_out, dst + 1, (double)Y[n] + origin[1]);
    store_io<Tio>(a.pos_out, dst + 2, (double)Z[n] + origin[2]);
  }
  if (tid == 0) { a.steps_run[c] = step; a.min_dist[c] = dmin; a.status[c] = status; }
}

#endif  // __CUDACC__

}  // namespace mch
