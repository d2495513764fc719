// mch_host.h -- host-side helpers shared by the translation units of libmchammer_b200.so.
#pragma once

#include <cuda_runtime.h>

#include "mch_optimize.cuh"

namespace mch {

// printf-style message into the calling thread's error buffer; returns `code`.
int fail(int code, const char* fmt, ...);

template <typename K, typename A>
int launch(K kernel, const A& args, int blocks, int threads, int smem, cudaStream_t stream, const char* what) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return fail(-3, "%s: cudaFuncSetAttribute(%d B): %s", what, smem, cudaGetErrorString(e));
  kernel<<<blocks, threads, smem, stream>>>(args);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail(-3, "%s: launch failed: %s", what, cudaGetErrorString(e));
  return 0;
}

// Kernel-variant dispatch of the resident Metropolis kernel (one TU per compute type, so the
// instantiations compile in parallel).  io_f64: element type of the position arrays.
int launch_optimize_f32(bool io_f64, int mu_kind, const OptArgs& a, int threads, int smem, cudaStream_t s);
int launch_optimize_f64(bool io_f64, int mu_kind, const OptArgs& a, int threads, int smem, cudaStream_t s);

int launch_probe_sweep(int blocks, int threads, int iters, int with_barrier, int nrows, int ncols, float* sink, cudaStream_t s);

template <typename T, typename Tio, int NTMAX>
int launch_optimize_nt(int mk, const OptArgs& a, int threads, int smem, cudaStream_t s) {
  switch (mk) {
    case MU_3: return launch(mch_optimize_kernel<T, Tio, MU_3, NTMAX>, a, a.C, threads, smem, s, "mch_optimize_batch");
    case MU_INT: return launch(mch_optimize_kernel<T, Tio, MU_INT, NTMAX>, a, a.C, threads, smem, s, "mch_optimize_batch");
    default: return launch(mch_optimize_kernel<T, Tio, MU_REAL, NTMAX>, a, a.C, threads, smem, s, "mch_optimize_batch");
  }
}

template <typename T, typename Tio>
int launch_optimize_io(int mk, const OptArgs& a, int threads, int smem, cudaStream_t s) {
  if (sizeof(T) == 4 && threads <= 32) return launch_optimize_nt<T, Tio, 32>(mk, a, threads, smem, s);
  if (sizeof(T) == 4 && threads <= 64) return launch_optimize_nt<T, Tio, 64>(mk, a, threads, smem, s);
  return threads <= small_block<T>() ? launch_optimize_nt<T, Tio, small_block<T>()>(mk, a, threads, smem, s)
                                     : launch_optimize_nt<T, Tio, large_block<T>()>(mk, a, threads, smem, s);
}

}  // namespace mch
