// mch_probes.cu -- measurement kernels for bench.py and profiles/*.py (NOT part of the product
// library or its ABI): arithmetic-pipe peaks for the roofline denominators, issue-mix probes,
// and the pair sweep alone (the Monte-Carlo kernel's inner tile loop without its control
// section).  Built by mchammer_b200.build into profiles/probes/libmch_probes.so.
#include <cstdint>
#include <cuda_runtime.h>

#include "../../mchammer_b200/csrc/mch_device.cuh"

namespace {
using mch::f32x2;
using mch::pack2;
using mch::unpack2;
using mch::fma2;
using namespace mch;

// ------------------------------------------------------------------ arithmetic probes
template <typename T>
__global__ void probe_fma_kernel(int iters, float* sink) {
  T a0 = (T)threadIdx.x * (T)1e-3, a1 = a0 + (T)1, a2 = a0 + (T)2, a3 = a0 + (T)3;
  T a4 = a0 + (T)4, a5 = a0 + (T)5, a6 = a0 + (T)6, a7 = a0 + (T)7;
  const T m = (T)0.999, c = (T)1e-3;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  const T r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (r == (T)-12345.678) *sink = (float)r;  // never true; keeps the loop alive
}

__global__ void probe_rsqrt_kernel(int iters, float* sink) {
  float a0 = 1.0f + threadIdx.x, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
  float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
  for (int i = 0; i < iters; ++i) {
    a0 = rsqrtf(a0); a1 = rsqrtf(a1); a2 = rsqrtf(a2); a3 = rsqrtf(a3);
    a4 = rsqrtf(a4); a5 = rsqrtf(a5); a6 = rsqrtf(a6); a7 = rsqrtf(a7);
  }
  const float r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (r == -12345.678f) *sink = r;
}

// Issue-mix probes (profiles/probe_mix.py): 8 independent instructions per iteration.
//   MODE 3: 8 FFMA2        MODE 4: 6 FFMA2 + 2 MUFU.RSQ
//   MODE 5: 6 FFMA + 2 MUFU.RSQ     MODE 6: 12 FFMA + 2 MUFU.RSQ (the scalar work of mode 4)
//   MODE 7: mode 4 + 1 broadcast LDS.128 feeding the FFMA2s
template <int MODE>
__global__ void probe_mix_kernel(int iters, float* sink) {
  const float t = 1e-3f * threadIdx.x;
  f32x2 p[6];
  float f[12], r0 = 1.5f + t, r1 = 2.5f + t;
  const f32x2 m2 = pack2(0.999f, 0.998f), c2 = pack2(1e-3f, 2e-3f);
#pragma unroll
  for (int k = 0; k < 6; ++k) p[k] = pack2(t + k, t - k);
#pragma unroll
  for (int k = 0; k < 12; ++k) f[k] = t + 0.5f * k;
  f32x2 q0 = pack2(t, 1.f), q1 = pack2(t, 2.f);
  __shared__ ulonglong2 rowbuf[64];
  if (MODE == 7) {
    if (threadIdx.x < 64) rowbuf[threadIdx.x] = make_ulonglong2(m2, c2);
    __syncthreads();
  }
  for (int i = 0; i < iters; ++i) {
    if (MODE == 3 || MODE == 4) {
#pragma unroll
      for (int k = 0; k < 6; ++k) p[k] = fma2(p[k], m2, c2);
    }
    if (MODE == 7) {  // mode 4 with its operands broadcast from shared memory (one LDS.128 per round)
      const ulonglong2 q = rowbuf[i & 63];
#pragma unroll
      for (int k = 0; k < 6; ++k) p[k] = fma2(p[k], q.x, q.y);
    }
    if (MODE == 3) { q0 = fma2(q0, m2, c2); q1 = fma2(q1, m2, c2); }
    if (MODE == 5 || MODE == 6) {
#pragma unroll
      for (int k = 0; k < (MODE == 5 ? 6 : 12); ++k) f[k] = fmaf(f[k], 0.999f, 1e-3f);
    }
    if (MODE != 3) { r0 = rsqrtf(r0); r1 = rsqrtf(r1); }
  }
  float lo, hi, r = r0 + r1;
#pragma unroll
  for (int k = 0; k < 6; ++k) { unpack2(p[k], lo, hi); r += lo + hi; }
  unpack2(q0, lo, hi); r += lo + hi;
  unpack2(q1, lo, hi); r += lo + hi;
#pragma unroll
  for (int k = 0; k < 12; ++k) r += f[k];
  if (r == -12345.678f) *sink = r;
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

static int launch_probe_sweep(int blocks, int threads, int iters, int with_barrier, int nrows, int ncols, float* sink, cudaStream_t s) {
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

}  // namespace

extern "C" {

/* Arithmetic-pipe probe: every thread runs `iters` rounds of 8 independent FMAs in `dtype`
 * (0 = f64, 1 = f32; 2 = 8 f32 MUFU.RSQ instead); 2*8*iters*blocks*threads flops per launch.
 * dtype 3..7 are issue-mix probes of 8 (mode 6: 14) instructions per round -- 3: 8 FFMA2,
 * 4: 6 FFMA2 + 2 MUFU.RSQ, 5: 6 FFMA + 2 MUFU.RSQ, 6: 12 FFMA + 2 MUFU.RSQ, 7: mode 4 with one
 * broadcast LDS.128.  Returns 0, or -1 (bad argument) / -3 (launch failed). */
int mch_probe_fma(int dtype, int blocks, int threads, int iters, float* sink, void* stream) {
  if (blocks <= 0 || threads <= 0 || threads > 1024 || iters <= 0 || !sink) return -1;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == 0) probe_fma_kernel<double><<<blocks, threads, 0, s>>>(iters, sink);
  else if (dtype == 1) probe_fma_kernel<float><<<blocks, threads, 0, s>>>(iters, sink);
  else if (dtype == 2) probe_rsqrt_kernel<<<blocks, threads, 0, s>>>(iters, sink);
  else if (dtype == 3) probe_mix_kernel<3><<<blocks, threads, 0, s>>>(iters, sink);
  else if (dtype == 4) probe_mix_kernel<4><<<blocks, threads, 0, s>>>(iters, sink);
  else if (dtype == 5) probe_mix_kernel<5><<<blocks, threads, 0, s>>>(iters, sink);
  else if (dtype == 6) probe_mix_kernel<6><<<blocks, threads, 0, s>>>(iters, sink);
  else if (dtype == 7) probe_mix_kernel<7><<<blocks, threads, 0, s>>>(iters, sink);
  else return -1;
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -3;
}

int mch_probe_sweep(int blocks, int threads, int iters, int with_barrier, int n_rows, int n_cols, float* sink, void* stream) {
  if (blocks <= 0 || threads < 32 || threads > 256 || threads % 32 || iters <= 0 || n_rows < 1 || n_rows > 60 ||
      n_cols < 1 || n_cols > 950 || !sink)
    return -1;
  return launch_probe_sweep(blocks, threads, iters, with_barrier, n_rows, n_cols, sink,
                            reinterpret_cast<cudaStream_t>(stream));
}


}  // extern "C"
