// mch_api.cu -- extern "C" surface of libmchammer_b200.so (see include/mchammer_b200.h).
// Host-side work here: argument checks, kernel-variant dispatch, launch geometry, and the
// numpy-compatible PCG64 seeding.  No device allocation, no synchronisation.
#include "../../include/mchammer_b200.h"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "mch_collapse.cuh"
#include "mch_device.cuh"
#include "mch_host.h"
#include "mch_optimize.cuh"
#include "mch_potential.cuh"
#include "mch_wide.cuh"

namespace {
thread_local char g_error[512] = "";
}

namespace mch {
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return code;
}
}  // namespace mch

namespace {
using mch::fail;
using mch::launch;

constexpr int kMaxSmem = 227 * 1024;

mch::Params to_params(const mch_params& p) {
  mch::Params q;
  q.step_size = p.step_size;
  q.target_bond_length = p.target_bond_length;
  q.bond_epsilon = p.bond_epsilon;
  q.nonbond_epsilon = p.nonbond_epsilon;
  q.nonbond_sigma = p.nonbond_sigma;
  q.nonbond_mu = p.nonbond_mu;
  q.beta = p.beta;
  return q;
}

// mu == 3 -> cubic fast path; other small positive integers -> square-and-multiply;
// anything else -> exp2/log2 (fp32) or pow (fp64).
int mu_kind(double mu, int* mu_int) {
  *mu_int = 0;
  if (mu == 3.0) return mch::MU_3;
  if (mu >= 1.0 && mu <= 64.0 && std::floor(mu) == mu) {
    *mu_int = (int)mu;
    return mch::MU_INT;
  }
  return mch::MU_REAL;
}

int sm_count() {  // SMs of the current device (queried once per device)
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cache[dev] == 0) {
    int n = 0;
    cache[dev] = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
  }
  return cache[dev];
}

// Block size when the caller leaves it open (measured: profiles/r02_chain_sweep.txt).
//   per_sm  = chains that will share an SM = min(chains that fit its shared memory, ceil(batch / SMs))
//   by_smem = chains that fit an SM's shared memory
// A batch that fills the GPU runs best with few warps per chain and as many resident chains as
// shared memory allows: every warp of a chain but one idles while the chain's serial control
// section runs, and with two warps per chain that loss (almost) disappears while eight resident
// chains keep the pipes busy.  A batch that leaves room (fewer chains per SM than would fit) is
// bound by the latency of a chain instead, so a chain gets more warps the fewer neighbours it
// has -- always at 128 registers per thread -- up to the point where more warps have no column
// units left to work on (S1 on one GPU, chain-steps/s relative to the full batch: 1024 chains
// 0.83 -> 0.89 with 128 threads, 512 chains 0.71 -> 0.77, 256 chains 0.47 -> 0.64 with 256).
// Molecules of which only two chains fit an SM (M12L24) fill the register file in fp32 with
// 2 x 512 threads at 64 registers (*regs64).
int auto_threads(int n_atoms, int per_sm, int by_smem, int compute_dtype, bool* regs64) {
  *regs64 = false;
  const bool f64 = compute_dtype == MCH_F64;
  if (n_atoms <= 128) return (per_sm <= 1 && f64) ? 128 : 32;
  if (!f64 && by_smem == 2 && per_sm == 2) { *regs64 = true; return 512; }
  // the widest block whose warps all get column units (a warp takes 128 columns per tile call)
  int useful = ((n_atoms / (f64 ? 2 : 4) + 31) / 32) * 32;
  useful = useful < 128 ? 128 : (useful > 512 ? 512 : useful);
  int t = per_sm >= 8 ? 64 : (per_sm >= 3 ? 128 : (per_sm == 2 ? 256 : 512));
  if (f64 && t < 128) t = 128;
  return t < useful ? t : useful;
}

int check_threads(int requested, int fallback, int limit, int* out) {
  int t = requested > 0 ? requested : fallback;
  if (t % 32 != 0 || t < 32 || t > limit)
    return fail(MCH_ERR_INVALID, "threads must be a multiple of 32 in [32, %d], got %d", limit, t);
  *out = t;
  return MCH_OK;
}

template <typename T, typename Tio>
int launch_potential_mu(int mk, const mch::PotArgs& a, int threads, int smem, cudaStream_t s) {
  switch (mk) {
    case mch::MU_3: return launch(mch::mch_potential_kernel<T, Tio, mch::MU_3>, a, a.C, threads, smem, s, "mch_potential_batch");
    case mch::MU_INT: return launch(mch::mch_potential_kernel<T, Tio, mch::MU_INT>, a, a.C, threads, smem, s, "mch_potential_batch");
    default: return launch(mch::mch_potential_kernel<T, Tio, mch::MU_REAL>, a, a.C, threads, smem, s, "mch_potential_batch");
  }
}

bool dtype_ok(int d) { return d == MCH_F64 || d == MCH_F32; }

// ------------------------------------------------------------------ numpy SeedSequence
// Published algorithm (numpy/random/bit_generator.pyx, after M.E. O'Neill's seed_seq_fe):
// 4-word uint32 pool, hashmix/mix, then generate_state; PCG64 takes 4 uint64 = 8 words.
void seed_sequence_8(uint64_t seed, uint32_t out[8]) {
  const uint32_t XSHIFT = 16, INIT_A = 0x43b0d7e5u, MULT_A = 0x931e8875u;
  const uint32_t INIT_B = 0x8b51f9ddu, MULT_B = 0x58f38dedu;
  const uint32_t MIX_L = 0xca01f9ddu, MIX_R = 0x4973f715u;
  uint32_t entropy[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  const int n_entropy = entropy[1] != 0 ? 2 : 1;  // numpy drops leading zero words (0 -> [0])
  uint32_t hash_const = INIT_A;
  auto hashmix = [&](uint32_t v) {
    v ^= hash_const;
    hash_const *= MULT_A;
    v *= hash_const;
    v ^= v >> XSHIFT;
    return v;
  };
  auto mix = [&](uint32_t x, uint32_t y) {
    uint32_t r = MIX_L * x - MIX_R * y;
    r ^= r >> XSHIFT;
    return r;
  };
  uint32_t pool[4];
  for (int i = 0; i < 4; ++i) pool[i] = hashmix(i < n_entropy ? entropy[i] : 0u);
  for (int src = 0; src < 4; ++src)
    for (int dst = 0; dst < 4; ++dst)
      if (src != dst) pool[dst] = mix(pool[dst], hashmix(pool[src]));
  uint32_t hc = INIT_B;
  for (int i = 0; i < 8; ++i) {
    uint32_t v = pool[i & 3];
    v ^= hc;
    hc *= MULT_B;
    v *= hc;
    v ^= v >> XSHIFT;
    out[i] = v;
  }
}

}  // namespace

extern "C" {

const char* mch_last_error(void) { return g_error; }
int mch_version(void) { return 100; }

namespace {
// Shared memory of one chain and whether the control work is deferred (second column-sum
// buffer).  Deferral hides the chain's serial control section, which matters for small and
// medium molecules; for a large molecule (fewer than three chains per SM) the control section
// is negligible next to the sweep, and the buffer is dropped when it would cost a resident
// chain (the 5004-atom cage: 115 KB with it -> one chain per SM, 95 KB without -> two) or
// would not fit at all.
int plan_chain_smem(int n_atoms, int n_subunits, int n_bonds, int max_su, int compute_dtype, int* defer,
                    int cluster = 1) {
  const int tsize = compute_dtype == MCH_F64 ? 8 : 4;
  if (cluster > 1) {  // a chain on a cluster always defers (the partial energy rows travel under the sweep)
    *defer = 1;
    return mch::opt_layout(n_atoms, n_subunits, n_bonds, max_su, tsize, 1, cluster).total;
  }
  const int with = mch::opt_layout(n_atoms, n_subunits, n_bonds, max_su, tsize, 1).total;
  const int without = mch::opt_layout(n_atoms, n_subunits, n_bonds, max_su, tsize, 0).total;
  const bool keep = with <= kMaxSmem && (kMaxSmem / with >= 3 || kMaxSmem / with == kMaxSmem / without);
  *defer = keep ? 1 : 0;
  return keep ? with : without;
}
}  // namespace

int mch_optimize_smem_bytes(int n_atoms, int n_subunits, int n_bonds, int max_subunit_atoms, int compute_dtype) {
  if (n_atoms <= 0 || n_subunits <= 0 || n_bonds < 0 || max_subunit_atoms <= 0 || !dtype_ok(compute_dtype))
    return fail(MCH_ERR_INVALID, "mch_optimize_smem_bytes: bad sizes");
  const long long est = 4LL * n_atoms * 8 + 8LL * n_subunits * n_subunits;
  if (est > (1LL << 30)) return fail(MCH_ERR_TOO_LARGE, "molecule too large");
  int defer;
  const int total = plan_chain_smem(n_atoms, n_subunits, n_bonds, max_subunit_atoms, compute_dtype, &defer);
  if (total > kMaxSmem) return fail(MCH_ERR_TOO_LARGE, "chain needs %d B of shared memory (limit %d)", total, kMaxSmem);
  return total;
}

namespace {
// The split-state kernel (csrc/mch_wide.cuh): explicit row lists, state split over a cluster.
int optimize_wide(const mch_optimize_desc* d, void* stream) {
  if (d->n_row_lists <= 0 || !d->bond_slots || !d->row_offsets || !d->row_slots || !d->row_weights ||
      !d->run_offsets || !d->runs || d->max_rows <= 0)
    return fail(MCH_ERR_INVALID, "mch_optimize_batch: this input needs the explicit row lists (bond_slots, row_*, run_*)");
  const int tsize = d->compute_dtype == MCH_F64 ? 8 : 4;
  int G = 0, chunk = 0, smem = 0;
  for (int g = 1; g <= mch::kWideMaxCluster; g *= 2) {
    const int ch = mch::align_up((d->n_atoms + g - 1) / g, 64);
    const long long est = 3LL * ch * tsize;
    if (est > kMaxSmem) continue;
    const int total = mch::wide_layout(ch, d->n_bonds, d->max_rows, tsize).total;
    if (total <= kMaxSmem) { G = g; chunk = ch; smem = total; break; }
  }
  if (G == 0) return fail(MCH_ERR_TOO_LARGE, "molecule of %d atoms does not fit the shared memory of a cluster of %d SMs", d->n_atoms, mch::kWideMaxCluster);
  if (d->cluster_size > 0) {
    if (d->cluster_size != 1 && d->cluster_size != 2 && d->cluster_size != 4 && d->cluster_size != 8 && d->cluster_size != 16)
      return fail(MCH_ERR_INVALID, "mch_optimize_batch: cluster_size must be 1, 2, 4, 8 or 16 for the split-state kernel");
    if (d->cluster_size < G) return fail(MCH_ERR_TOO_LARGE, "the molecule needs a cluster of at least %d CTAs", G);
    G = d->cluster_size;
  } else {
    // few chains of a big molecule: more CTAs per chain while SMs are free and every CTA keeps >= 512 columns
    const long long batch = d->resident_chains_hint > d->n_chains ? d->resident_chains_hint : d->n_chains;
    while (G < 8 && batch * 2 * G <= sm_count() && d->n_atoms / (2 * G) >= 512) G *= 2;
  }
  chunk = mch::align_up((d->n_atoms + G - 1) / G, 64);
  smem = mch::wide_layout(chunk, d->n_bonds, d->max_rows, tsize).total;
  int threads;
  const int auto_t = chunk <= 64 ? 32 : (chunk <= 256 ? 64 : (chunk <= 1024 ? 128 : (chunk <= 4096 ? 256 : 512)));
  int rc = check_threads(d->threads_per_chain, auto_t, 512, &threads);
  if (rc != MCH_OK) return rc;
  mch::WideArgs a;
  a.pos_in = d->pos_in; a.pos_in_stride = d->pos_in_chain_stride; a.pos_out = d->pos_out;
  a.C = d->n_chains; a.N = d->n_atoms; a.S = d->n_row_lists; a.B = d->n_bonds;
  a.G = G; a.chunk = chunk; a.max_rows = d->max_rows;
  a.perm = d->perm; a.bond_slot = d->bond_slots; a.bond_su = d->bond_su;
  a.row_off = d->row_offsets; a.row_slot = d->row_slots; a.row_w = d->row_weights;
  a.run_off = d->run_offsets; a.runs = d->runs;
  a.p = to_params(d->params);
  const int mk = mu_kind(d->params.nonbond_mu, &a.mu_int);
  a.num_steps = d->num_steps < 1 ? 1 : d->num_steps;
  a.rng = d->rng_state;
  a.U = d->system_potential; a.Unb = d->nonbonded_potential; a.maxd = d->max_bond_distance;
  a.n_accept = d->n_accepted; a.last_info = d->last_step_info;
  a.trace_f = d->trace_potentials; a.trace_i = d->trace_info; a.traj = d->trajectory;
  a.pairs = reinterpret_cast<unsigned long long*>(d->pair_evals);
  a.inexact_undo = d->inexact_undo;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const bool io_f64 = d->io_dtype == MCH_F64;
  return d->compute_dtype == MCH_F64 ? mch::launch_wide_f64(io_f64, mk, a, threads, smem, s)
                                     : mch::launch_wide_f32(io_f64, mk, a, threads, smem, s);
}
}  // namespace

int mch_optimize_batch(const mch_optimize_desc* d, void* stream) {
  if (!d) return fail(MCH_ERR_INVALID, "mch_optimize_batch: null descriptor");
  if (d->n_chains <= 0 || d->n_atoms <= 0 || d->n_subunits <= 0 || d->n_bonds <= 0 || d->max_subunit_atoms <= 0)
    return fail(MCH_ERR_INVALID, "mch_optimize_batch: sizes must be positive (chains %d atoms %d subunits %d bonds %d max_su %d)",
                d->n_chains, d->n_atoms, d->n_subunits, d->n_bonds, d->max_subunit_atoms);
  if (d->n_bonds > 65535 || d->n_subunits > 8191)
    return fail(MCH_ERR_INVALID, "mch_optimize_batch: at most 65535 bonds and 8191 subunits");
  if (!dtype_ok(d->io_dtype) || !dtype_ok(d->compute_dtype)) return fail(MCH_ERR_INVALID, "mch_optimize_batch: unknown dtype");
  if (!d->pos_in || !d->pos_out || !d->perm || !d->su_offsets || !d->bonds || !d->bond_su || !d->rng_state ||
      !d->system_potential || !d->nonbonded_potential || !d->max_bond_distance || !d->n_accepted || !d->last_step_info)
    return fail(MCH_ERR_INVALID, "mch_optimize_batch: null required pointer");
  if (d->general < 0 || d->general > 2) return fail(MCH_ERR_INVALID, "mch_optimize_batch: general must be 0, 1 or 2");
  if (d->general != 0) return optimize_wide(d, stream);
  int smem = mch_optimize_smem_bytes(d->n_atoms, d->n_subunits, d->n_bonds, d->max_subunit_atoms, d->compute_dtype);
  if (smem == MCH_ERR_TOO_LARGE && d->n_row_lists > 0) return optimize_wide(d, stream);  // does not fit one CTA
  if (smem < 0) return smem;
  // chains that are (or will be) resident together: a caller that cuts one batch into several
  // launches (BatchPlan's pipelined slices) says so, and every slice gets the batch's choices
  const int sms = sm_count();
  const long long batch = d->resident_chains_hint > d->n_chains ? d->resident_chains_hint : d->n_chains;
  // Thread-block cluster per chain: worth it when the GPU would otherwise be mostly idle and the
  // sweep is long enough to pay for one cluster barrier per step (big molecule, few chains).
  int cluster = d->cluster_size;
  if (cluster != 0 && cluster != 1 && cluster != 2 && cluster != 4 && cluster != 8 && cluster != 16)
    return fail(MCH_ERR_INVALID, "mch_optimize_batch: cluster_size must be 0 (auto), 1, 2, 4, 8 or 16, got %d", cluster);
  // an explicit block size outside what a chain on a cluster takes keeps the chain on one CTA
  const bool threads_fit_cluster = d->threads_per_chain == 0 || (d->threads_per_chain >= 64 && d->threads_per_chain <= 512);
  if (cluster == 0) {
    cluster = 1;
    if (d->n_atoms >= 768 && d->num_steps > 1 && threads_fit_cluster) {  // measured: profiles/r01_cluster_chain.txt
      // 16 CTAs (the non-portable cluster size; one such cluster per GPC) from 4096 atoms: the sweep of one
      // chain is bound by the arithmetic pipes of the SMs it runs on (profiles/r02_cluster_chain.txt)
      for (int g = d->n_atoms >= 4096 ? 16 : (d->n_atoms >= 2048 ? 8 : 4); g >= 2; g /= 2)
        if (batch * g <= sms && (g < 16 || batch <= 8)) {
          int defer_unused;
          if (plan_chain_smem(d->n_atoms, d->n_subunits, d->n_bonds, d->max_subunit_atoms, d->compute_dtype,
                              &defer_unused, g) <= kMaxSmem) cluster = g;
          break;
        }
    }
  }
  int defer = 0;
  if (cluster > 1) {
    smem = plan_chain_smem(d->n_atoms, d->n_subunits, d->n_bonds, d->max_subunit_atoms, d->compute_dtype, &defer, cluster);
    if (smem > kMaxSmem)
      return fail(MCH_ERR_TOO_LARGE, "a chain on a cluster needs %d B of shared memory per CTA (limit %d)", smem, kMaxSmem);
  } else {
    (void)plan_chain_smem(d->n_atoms, d->n_subunits, d->n_bonds, d->max_subunit_atoms, d->compute_dtype, &defer);
  }
  const int by_smem = kMaxSmem / smem;
  const long long spread = (batch + sms - 1) / sms;
  const int per_sm = (int)(spread < by_smem ? spread : by_smem);
  bool regs64 = false;
  int threads;
  int rc = check_threads(d->threads_per_chain,
                         cluster > 1 ? 512 : auto_threads(d->n_atoms, per_sm, by_smem, d->compute_dtype, &regs64),
                         d->compute_dtype == MCH_F64 ? 512 : 1024, &threads);
  if (rc != MCH_OK) return rc;
  if (d->threads_per_chain > 0) regs64 = false;  // an explicit block size: the 128-register family up to 512 threads
  if (cluster > 1 && (threads < 64 || threads > 512))  // only reachable with an explicit cluster_size
    return fail(MCH_ERR_INVALID, "a chain on a cluster takes 64 to 512 threads per CTA, got %d", threads);

  mch::OptArgs a;
  a.pos_in = d->pos_in; a.pos_in_stride = d->pos_in_chain_stride; a.pos_out = d->pos_out;
  a.C = d->n_chains; a.N = d->n_atoms; a.S = d->n_subunits; a.B = d->n_bonds; a.max_su = d->max_subunit_atoms;
  a.perm = d->perm; a.su_off = d->su_offsets; a.bonds = d->bonds; a.bond_su = d->bond_su;
  a.p = to_params(d->params);
  const int mk = mu_kind(d->params.nonbond_mu, &a.mu_int);
  a.num_steps = d->num_steps < 1 ? 1 : d->num_steps;
  a.rng = d->rng_state;
  a.U = d->system_potential; a.Unb = d->nonbonded_potential; a.maxd = d->max_bond_distance;
  a.n_accept = d->n_accepted; a.last_info = d->last_step_info;
  a.trace_f = d->trace_potentials; a.trace_i = d->trace_info; a.traj = d->trajectory;
  a.pairs = reinterpret_cast<unsigned long long*>(d->pair_evals);
  a.inexact_undo = d->inexact_undo;
  a.defer = defer;
  a.cluster = cluster;
  a.regs64 = regs64 ? 1 : 0;
  // early rejection (opt-in, MCH_FLAG_EARLY_REJECTION): one-CTA chains of at least two warps (fp64: four) that defer
  // their control work, mu = 3, no per-step records
  const bool f64c = d->compute_dtype == MCH_F64;
  a.early = ((d->flags & MCH_FLAG_EARLY_REJECTION) && cluster == 1 && defer && mk == mch::MU_3 &&
             threads >= (f64c ? 128 : 64) &&
             !(!f64c && d->inexact_undo) &&  // fp32 undo: positions drift by whole fp32 ulps
             !d->trace_potentials && !d->trace_info && !d->trajectory && d->params.beta > 0.0 &&
             d->params.nonbond_epsilon >= 0.0 && d->params.nonbond_sigma >= 0.0) ? 1 : 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const bool io_f64 = d->io_dtype == MCH_F64;
  return d->compute_dtype == MCH_F64 ? mch::launch_optimize_f64(io_f64, mk, a, threads, smem, s)
                                     : mch::launch_optimize_f32(io_f64, mk, a, threads, smem, s);
}

int mch_potential_batch(const void* pos, int io_dtype, int compute_dtype, int n_mol, int n_atoms,
                        const int32_t* bonds, int n_bonds, const mch_params* params,
                        double* system_potential, double* nonbonded_potential,
                        double* max_bond_distance, void* stream) {
  if (!pos || !params || !system_potential || !nonbonded_potential || !max_bond_distance)
    return fail(MCH_ERR_INVALID, "mch_potential_batch: null required pointer");
  if (n_mol <= 0 || n_atoms <= 0 || n_bonds < 0 || (n_bonds > 0 && !bonds))
    return fail(MCH_ERR_INVALID, "mch_potential_batch: bad sizes");
  if (!dtype_ok(io_dtype) || !dtype_ok(compute_dtype)) return fail(MCH_ERR_INVALID, "mch_potential_batch: unknown dtype");
  if (n_atoms > (1 << 20)) return fail(MCH_ERR_TOO_LARGE, "molecule too large");
  const mch::PotLayout L = mch::pot_layout(n_atoms, compute_dtype == MCH_F64 ? 8 : 4);
  if (L.total > kMaxSmem) return fail(MCH_ERR_TOO_LARGE, "molecule needs %d B of shared memory (limit %d)", L.total, kMaxSmem);
  mch::PotArgs a;
  a.pos = pos; a.C = n_mol; a.N = n_atoms; a.B = n_bonds; a.bonds = bonds;
  a.p = to_params(*params);
  const int mk = mu_kind(params->nonbond_mu, &a.mu_int);
  a.U = system_potential; a.Unb = nonbonded_potential; a.maxd = max_bond_distance;
  const int threads = n_atoms <= 64 ? 32 : (n_atoms <= 256 ? 64 : (n_atoms <= 1024 ? 128 : 256));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (compute_dtype == MCH_F64) {
    return io_dtype == MCH_F64 ? launch_potential_mu<double, double>(mk, a, threads, L.total, s)
                               : launch_potential_mu<double, float>(mk, a, threads, L.total, s);
  }
  return io_dtype == MCH_F64 ? launch_potential_mu<float, double>(mk, a, threads, L.total, s)
                             : launch_potential_mu<float, float>(mk, a, threads, L.total, s);
}

int mch_collapse_batch(const mch_collapse_desc* d, void* stream) {
  if (!d) return fail(MCH_ERR_INVALID, "mch_collapse_batch: null descriptor");
  if (d->n_mol <= 0 || d->n_atoms <= 0 || d->n_subunits <= 0 || d->n_bonds < 0 || d->max_steps < 0 ||
      d->n_moving_subunits < 0 || d->n_moving_subunits > d->n_subunits)
    return fail(MCH_ERR_INVALID, "mch_collapse_batch: bad sizes");
  if (!dtype_ok(d->io_dtype) || !dtype_ok(d->compute_dtype)) return fail(MCH_ERR_INVALID, "mch_collapse_batch: unknown dtype");
  if (!d->pos_in || !d->pos_out || !d->perm || !d->su_offsets || !d->n_heavy || (d->n_bonds > 0 && !d->bonds) ||
      !d->steps_run || !d->min_distance || !d->max_bond_distance || !d->status)
    return fail(MCH_ERR_INVALID, "mch_collapse_batch: null required pointer");
  if ((d->trace_max_bond || d->trajectory) && d->trace_capacity <= 0)
    return fail(MCH_ERR_INVALID, "mch_collapse_batch: trace buffers need trace_capacity > 0");
  if (d->n_atoms > (1 << 20) || d->n_subunits > (1 << 16)) return fail(MCH_ERR_TOO_LARGE, "molecule too large");
  int threads;
  // few molecules (at most one per SM): latency case, a molecule gets the largest block that still has work
  const int batch_threads = d->n_atoms <= 128 ? 32 : (d->n_atoms <= 384 ? 64 : (d->n_atoms <= 3072 ? 128 : 256));
  const int alone_threads = d->n_atoms <= 128 ? 32 : (d->n_atoms <= 512 ? 128 : 256);
  int rc = check_threads(d->threads_per_mol, d->n_mol <= sm_count() ? alone_threads : batch_threads, 256, &threads);
  if (rc != MCH_OK) return rc;
  const int nh_total = (d->n_heavy_total > 0 && d->n_heavy_total <= d->n_atoms) ? d->n_heavy_total : d->n_atoms;
  const mch::ColLayout L = mch::col_layout(d->n_atoms, d->n_subunits, d->n_bonds, d->compute_dtype == MCH_F64 ? 8 : 4, nh_total);
  if (L.total > kMaxSmem) return fail(MCH_ERR_TOO_LARGE, "molecule needs %d B of shared memory (limit %d)", L.total, kMaxSmem);
  mch::ColArgs a;
  a.pos_in = d->pos_in; a.pos_in_stride = d->pos_in_mol_stride; a.pos_out = d->pos_out;
  a.C = d->n_mol; a.N = d->n_atoms; a.S = d->n_subunits; a.B = d->n_bonds;
  a.S_moving = d->n_moving_subunits > 0 ? d->n_moving_subunits : d->n_subunits;
  a.perm = d->perm; a.su_off = d->su_offsets; a.n_heavy = d->n_heavy; a.bonds = d->bonds;
  a.step_size = d->step_size; a.threshold = d->distance_threshold; a.scale_steps = d->scale_steps;
  a.max_steps = d->max_steps;
  a.nh_total = nh_total;
  a.steps_run = d->steps_run; a.min_dist = d->min_distance; a.maxd_last = d->max_bond_distance; a.status = d->status;
  a.trace_maxd = d->trace_max_bond; a.traj = d->trajectory; a.trace_cap = d->trace_capacity;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (d->compute_dtype == MCH_F64) {
    return d->io_dtype == MCH_F64
               ? launch(mch::mch_collapse_kernel<double, double>, a, a.C, threads, L.total, s, "mch_collapse_batch")
               : launch(mch::mch_collapse_kernel<double, float>, a, a.C, threads, L.total, s, "mch_collapse_batch");
  }
  return d->io_dtype == MCH_F64
             ? launch(mch::mch_collapse_kernel<float, double>, a, a.C, threads, L.total, s, "mch_collapse_batch")
             : launch(mch::mch_collapse_kernel<float, float>, a, a.C, threads, L.total, s, "mch_collapse_batch");
}

int mch_pcg64_seed(const uint64_t* seeds, int n, uint64_t* out) {
  if (!seeds || !out || n < 0) return fail(MCH_ERR_INVALID, "mch_pcg64_seed: bad arguments");
  for (int k = 0; k < n; ++k) {
    uint32_t w[8];
    seed_sequence_8(seeds[k], w);
    uint64_t v[4];
    for (int i = 0; i < 4; ++i) v[i] = (uint64_t)w[2 * i] | ((uint64_t)w[2 * i + 1] << 32);
    // pcg_setseq_128_srandom_r(initstate = v0:v1, initseq = v2:v3)
    mch::Pcg64 g;
    g.inc_hi = (v[2] << 1) | (v[3] >> 63);
    g.inc_lo = (v[3] << 1) | 1ULL;
    g.hi = 0; g.lo = 0; g.has32 = 0; g.half = 0;
    mch::pcg_advance(g);
    const uint64_t lo = g.lo + v[1];
    g.hi = g.hi + v[0] + (lo < g.lo ? 1ULL : 0ULL);
    g.lo = lo;
    mch::pcg_advance(g);
    mch::pcg_store(g, out + 6LL * k);
  }
  return MCH_OK;
}

int mch_pcg64_stream_check(uint64_t* rng_state, int rounds, uint32_t bound, int64_t* ints_out, double* doubles_out) {
  if (!rng_state || !ints_out || !doubles_out || rounds < 0 || bound == 0) return fail(MCH_ERR_INVALID, "mch_pcg64_stream_check: bad arguments");
  mch::Pcg64 g = mch::pcg_load(rng_state);
  int ni = 0, nd = 0;
  for (int k = 0; k < rounds; ++k) {
    ints_out[ni++] = mch::pcg_bounded(g, bound);
    ints_out[ni++] = mch::pcg_bounded(g, 2u);
    doubles_out[nd++] = mch::pcg_double(g);
    ints_out[ni++] = mch::pcg_bounded(g, 2u);
    if (k % 3 == 0) doubles_out[nd++] = mch::pcg_double(g);
  }
  mch::pcg_store(g, rng_state);
  return MCH_OK;
}

}  // extern "C"
