// fp32-compute instantiations of the resident Metropolis kernel (fast mode).
#include "mch_host.h"

namespace mch {
int launch_optimize_f32(bool io_f64, int mk, const OptArgs& a, int threads, int smem, cudaStream_t s) {
#ifdef MCH_QUICK
  if (!io_f64) return fail(-3, "MCH_QUICK build: f64 position arrays only");
  return launch_optimize_io<float, double>(mk, a, threads, smem, s);
#else
  return io_f64 ? launch_optimize_io<float, double>(mk, a, threads, smem, s)
                : launch_optimize_io<float, float>(mk, a, threads, smem, s);
#endif
}

}  // namespace mch
