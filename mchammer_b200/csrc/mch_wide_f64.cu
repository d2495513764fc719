// fp64-compute instantiations of the split-state / general-subunit Metropolis kernel.
#include "mch_host.h"
#include "mch_wide.cuh"

namespace mch {
int launch_wide_f64(bool io_f64, int mk, const WideArgs& a, int threads, int smem, cudaStream_t s) {
#ifdef MCH_QUICK
  if (!io_f64) return fail(-3, "MCH_QUICK build: f64 position arrays only");
  return launch_wide_io<double, double>(mk, a, threads, smem, s);
#else
  return io_f64 ? launch_wide_io<double, double>(mk, a, threads, smem, s)
                : launch_wide_io<double, float>(mk, a, threads, smem, s);
#endif
}
}  // namespace mch
