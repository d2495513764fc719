// mch_host.h -- host-side helpers shared by the translation units of libmchammer_b200.so.
#pragma once

#include <cuda_runtime.h>

#include "mch_optimize.cuh"

namespace mch {

struct WideArgs;

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

// Same, as thread-block clusters of `cluster` CTAs (blocks = chains x cluster).
template <typename K, typename A>
int launch_clustered(K kernel, const A& args, int blocks, int threads, int smem, int cluster, cudaStream_t stream,
                     const char* what) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return fail(-3, "%s: cudaFuncSetAttribute(%d B): %s", what, smem, cudaGetErrorString(e));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)blocks, 1, 1);
  cfg.blockDim = dim3((unsigned)threads, 1, 1);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kernel, args);
  if (e != cudaSuccess) return fail(-3, "%s: cluster launch (%d CTAs per chain) failed: %s", what, cluster, cudaGetErrorString(e));
  return 0;
}

// Same, allowing the non-portable cluster size 16.
template <typename K, typename A>
int launch_clustered_wide(K kernel, const A& args, int blocks, int threads, int smem, int cluster, cudaStream_t stream,
                          const char* what) {
  if (cluster > 8) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return fail(-3, "%s: cluster of %d CTAs not allowed: %s", what, cluster, cudaGetErrorString(e));
  }
  return launch_clustered(kernel, args, blocks, threads, smem, cluster, stream, what);
}

int launch_wide_f32(bool io_f64, int mu_kind, const WideArgs& a, int threads, int smem, cudaStream_t s);
int launch_wide_f64(bool io_f64, int mu_kind, const WideArgs& a, int threads, int smem, cudaStream_t s);

// Kernel-variant dispatch of the resident Metropolis kernel (one TU per compute type, so the
// instantiations compile in parallel).  io_f64: element type of the position arrays.
int launch_optimize_f32(bool io_f64, int mu_kind, const OptArgs& a, int threads, int smem, cudaStream_t s);
int launch_optimize_f64(bool io_f64, int mu_kind, const OptArgs& a, int threads, int smem, cudaStream_t s);

template <typename T, typename Tio, int NTMAX>
int launch_optimize_nt(int mk, const OptArgs& a, int threads, int smem, cudaStream_t s) {
#ifdef MCH_QUICK  // tuning builds (profiles/quick_variant.sh): mu = 3 only, seconds instead of minutes
  if (mk != MU_3) return fail(-3, "MCH_QUICK build: mu = 3 only");
  return launch(mch_optimize_kernel<T, Tio, MU_3, NTMAX>, a, a.C, threads, smem, s, "mch_optimize_batch");
#else
  switch (mk) {
    case MU_3: return launch(mch_optimize_kernel<T, Tio, MU_3, NTMAX>, a, a.C, threads, smem, s, "mch_optimize_batch");
    case MU_INT: return launch(mch_optimize_kernel<T, Tio, MU_INT, NTMAX>, a, a.C, threads, smem, s, "mch_optimize_batch");
    default: return launch(mch_optimize_kernel<T, Tio, MU_REAL, NTMAX>, a, a.C, threads, smem, s, "mch_optimize_batch");
  }
#endif
}

// a chain spread over a.cluster CTAs: the large-block instantiation, cluster-capable
template <typename T, typename Tio>
int launch_optimize_cluster(int mk, const OptArgs& a, int threads, int smem, cudaStream_t s) {
  constexpr int NT = 512;  // latency case: 512 threads at 128 registers in both precisions (no spills in the control section)
  const int blocks = a.C * a.cluster;
#ifdef MCH_QUICK
  if (mk != MU_3) return fail(-3, "MCH_QUICK build: mu = 3 only");
  return launch_clustered_wide(mch_optimize_kernel<T, Tio, MU_3, NT, true>, a, blocks, threads, smem, a.cluster, s, "mch_optimize_batch");
#else
  switch (mk) {
    case MU_3: return launch_clustered_wide(mch_optimize_kernel<T, Tio, MU_3, NT, true>, a, blocks, threads, smem, a.cluster, s, "mch_optimize_batch");
    case MU_INT: return launch_clustered_wide(mch_optimize_kernel<T, Tio, MU_INT, NT, true>, a, blocks, threads, smem, a.cluster, s, "mch_optimize_batch");
    default: return launch_clustered_wide(mch_optimize_kernel<T, Tio, MU_REAL, NT, true>, a, blocks, threads, smem, a.cluster, s, "mch_optimize_batch");
  }
#endif
}

// the EARLY instantiations (mch_optimize.cuh; mu = 3, one CTA per chain): the same block-size families as below,
// 257..512 threads on a 512-thread instantiation of their own (the cluster-capable one has no early form)
template <typename T, typename Tio>
int launch_optimize_early(const OptArgs& a, int threads, int smem, cudaStream_t s) {
  auto go = [&](auto kernel) { return launch(kernel, a, a.C, threads, smem, s, "mch_optimize_batch"); };
  if constexpr (sizeof(T) == 4) {
    if (a.regs64 || threads > 512) return go(mch_optimize_kernel<T, Tio, MU_3, 1024, false, true>);
    if (threads <= 64) return go(mch_optimize_kernel<T, Tio, MU_3, 64, false, true>);
  }
  if (threads <= 128) return go(mch_optimize_kernel<T, Tio, MU_3, 128, false, true>);
  if (threads <= 256) return go(mch_optimize_kernel<T, Tio, MU_3, 256, false, true>);
  return go(mch_optimize_kernel<T, Tio, MU_3, 512, false, true>);
}

template <typename T, typename Tio>
int launch_optimize_io(int mk, const OptArgs& a, int threads, int smem, cudaStream_t s) {
  if (a.cluster > 1) return launch_optimize_cluster<T, Tio>(mk, a, threads, smem, s);
  if (a.early && mk == MU_3) return launch_optimize_early<T, Tio>(a, threads, smem, s);  // the host sets a.early only for what exists
  // blocks that share an SM with so many others that 128 registers per thread do not fit
  // (a.regs64, set by the host), and blocks of more than 512 threads: the 64-register instantiation
  if constexpr (sizeof(T) == 4) {
    if (a.regs64 || threads > 512) return launch_optimize_nt<T, Tio, 1024>(mk, a, threads, smem, s);
    if (threads <= 32) return launch_optimize_nt<T, Tio, 32>(mk, a, threads, smem, s);
    if (threads <= 64) return launch_optimize_nt<T, Tio, 64>(mk, a, threads, smem, s);
  }
  if (threads <= 128) return launch_optimize_nt<T, Tio, 128>(mk, a, threads, smem, s);
  if (threads <= 256) return launch_optimize_nt<T, Tio, 256>(mk, a, threads, smem, s);
  return launch_optimize_cluster<T, Tio>(mk, a, threads, smem, s);  // 257..512 threads: a cluster of one
}

}  // namespace mch
