// fp32-compute instantiations of the resident Metropolis kernel (fast mode).
#include "mch_host.h"

namespace mch {
int launch_optimize_f32(bool io_f64, int mk, const OptArgs& a, int threads, int smem, cudaStream_t s) {
  return io_f64 ? launch_optimize_io<float, double>(mk, a, threads, smem, s)
                : launch_optimize_io<float, float>(mk, a, threads, smem, s);
}

}  // namespace mch
