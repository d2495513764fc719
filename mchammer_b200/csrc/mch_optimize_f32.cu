// fp32-compute instantiations of the resident Metropolis kernel (fast mode).
#include "mch_host.h"

namespace mch {
int launch_optimize_f32(bool io_f64, int mk, const OptArgs& a, int threads, int smem, cudaStream_t s) {
  return io_f64 ? launch_optimize_io<float, double>(mk, a, threads, smem, s)
                : launch_optimize_io<float, float>(mk, a, threads, smem, s);
}

// ---------------------------------------------------------------------------------------
// Ceiling probe: the pair sweep alone (no control section) on a synthetic chain, `iters`
// sweeps of 50 rows x 950 columns per CTA, optionally separated by __syncthreads().
// ---------------------------------------------------------------------------------------
template <int NTMAX>
__global__ void __launch_bounds__(NTMAX, (NTMAX == 64 ? 8 : 1024 / NTMAX))
probe_sweep_kernel(int iters, int with_barrier, int nrows_real, int ncols, float* sink, int store) {
  extern __shared__ __align__(32) unsigned char smem[];
  const int N = 1000;
  float* X = reinterpret_cast<float*>(smem);
  float* Y = X + 1008;
  float* Z = Y + 1008;
  float* colsum = Z + 1008;  // only present when store != 0
  Row4<float>* Q = reinterpret_cast<Row4<float>*>(colsum + (store ? 2016 : 0));
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    X[n] = 0.013f * n + 0.5f * blockIdx.x; Y[n] = 7.0f - 0.009f * n; Z[n] = 0.003f * n * (n & 7);
  }
  constexpr bool kXF = (MCH_XF != 0);
  constexpr bool kX2 = kXF && (MCH_F32X2 != 0);
  constexpr int kRot3 = !(kX2 && (MCH_X2_ROT3 != 0)) ? 0 : (NTMAX <= 64 ? 2 : 1);
  const Centre<float> ctr{-5.7f, 21.7f, -7.8f};
  const int np = kRot3 ? padded_rows3(nrows_real) : padded_rows(nrows_real);
  for (int i = threadIdx.x; i < np + 3; i += blockDim.x) {
    Q[i] = i < nrows_real ? encode_row<float, kXF>(-3.0f - 0.11f * i, 20.0f + 0.07f * i, -9.0f + 0.05f * i, ctr)
                          : padding_row<float, kXF>();
  }
  __syncthreads();
  const PairTerm<float, MU_3> term(3.0, 3);
  double total = 0.0;
  for (int it = 0; it < iters; ++it) {
    Cols cols{0, 0, 50, 50 + ncols};
    if (store) total += pair_sweep<float, MU_3, false, true, kXF, kX2, kRot3>(Q, np, X, Y, Z, colsum + (it & 1) * 1008, cols, term, ctr);
    else total += pair_sweep<float, MU_3, false, false, kXF, kX2, kRot3>(Q, np, X, Y, Z, (float*)nullptr, cols, term, ctr);
    if (with_barrier) __syncthreads();
  }
  if (total == -1.2345) *sink = (float)total;
}

int launch_probe_sweep(int blocks, int threads, int iters, int with_barrier, int nrows, int ncols, float* sink, cudaStream_t s) {
  const int store = with_barrier < 2 ? 1 : 0;  // with_barrier 2/3 = 0/1 without the column-sum stores
  with_barrier &= 1;
  const int smem = (1008 * 3 + (store ? 2016 : 0)) * 4 + 72 * 32 + 256;
  if (threads <= 64) {
    cudaFuncSetAttribute(probe_sweep_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe_sweep_kernel<64><<<blocks, threads, smem, s>>>(iters, with_barrier, nrows, ncols, sink, store);
  } else if (threads <= 128) {
    cudaFuncSetAttribute(probe_sweep_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe_sweep_kernel<128><<<blocks, threads, smem, s>>>(iters, with_barrier, nrows, ncols, sink, store);
  } else {
    cudaFuncSetAttribute(probe_sweep_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe_sweep_kernel<256><<<blocks, threads, smem, s>>>(iters, with_barrier, nrows, ncols, sink, store);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}
}  // namespace mch
