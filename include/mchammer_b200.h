/* mchammer_b200.h -- C ABI of libmchammer_b200.so (CUDA, sm_100a).
 *
 * The reference (andrewtarzia/MCHammer) is pure Python and has NO foreign-function layer;
 * the boundary of its Monte-Carlo hot path is a handful of Python methods.  Each entry
 * point below names the reference interface it stands in for.  A binding needs only
 * ctypes: every pointer marked [dev] is a CUDA device pointer on
 * the current device (e.g. torch.Tensor.data_ptr()), `stream` is a cudaStream_t passed
 * as void* (NULL = legacy default stream).  Calls are asynchronous on `stream`; no call
 * synchronises, allocates device memory, or keeps global mutable state, so the library
 * is thread-safe per stream.  Nothing throws across the ABI: every function returns
 * MCH_OK or a negative code, and mch_last_error() gives the text for the calling thread.
 */
#ifndef MCHAMMER_B200_H
#define MCHAMMER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCH_OK 0
#define MCH_ERR_INVALID (-1)   /* bad argument (null pointer, size <= 0, unknown dtype ...) */
#define MCH_ERR_TOO_LARGE (-2) /* molecule does not fit shared memory (one SM, or a cluster of 16 with row lists) */
#define MCH_ERR_CUDA (-3)      /* CUDA runtime error; text in mch_last_error()             */

#define MCH_F64 0
#define MCH_F32 1

/* mch_optimize_desc.flags.  EARLY_REJECTION: proposals that a pair-free lower bound of U_new - U already rejects
 * (reference test_move, mc_operations.py:39-52, with the next random double read ahead) are booked
 * by the chain's control warp without a pair sweep -- the same decisions, the same random stream and the same bits
 * in every output as the full evaluation; only pair_evals shrinks.  Honoured for one-CTA chains of at least 64
 * threads (fp64 compute: 128) that defer their control work (the library's choice whenever a second column-sum
 * buffer costs no resident chain), with mu = 3, beta > 0, non-negative epsilon / sigma and no per-step records;
 * ignored (full evaluation) otherwise. */
#define MCH_FLAG_EARLY_REJECTION 1

/* Constructor arguments of Optimizer (reference src/mchammer/_internal/optimizer.py:27-38). */
typedef struct mch_params {
  double step_size;
  double target_bond_length;
  double bond_epsilon;
  double nonbond_epsilon;
  double nonbond_sigma;
  double nonbond_mu;
  double beta;
} mch_params;

/* ---------------------------------------------------------------------------------------
 * mch_optimize_batch -- C independent Metropolis chains, all num_steps resident on device.
 * Stands in for Optimizer.get_result / get_trajectory (optimizer.py:308-438) and everything
 * below them: _run_first_step (:164-200), _run_step (:202-306), compute_potential
 * (:122-142), translate_atoms_along_vector (mc_operations.py:15-28), test_move
 * (mc_operations.py:39-52).
 *
 * Topology (shared by all chains) is given in *slots*: a permutation of the atoms that makes
 * every subunit a contiguous range (mchammer_b200.lowering builds it):
 *   perm[slot]            atom id living in that slot
 *   su_offsets[s..s+1]    slot range of subunit s (dict order of the caller's `subunits`)
 *   bonds[b][0..1]        atom ids of long bond b, in the caller's order (sign of the bond vector)
 *   bond_su[b][0..1]      subunit of each end = first dict entry containing the atom (:220-221)
 * rng_state[c] = numpy PCG64 state of chain c as 6 words
 *   {state_hi, state_lo, inc_hi, inc_lo, has_uint32, uinteger}; advanced in place, so a
 *   second call continues the stream exactly like a second get_result on one Optimizer.
 * trace_info word: bond index | passed << 16 | move kind << 17 | moved subunit << 18; -1 at step 0.
 * ------------------------------------------------------------------------------------- */
typedef struct mch_optimize_desc {
  int32_t n_chains, n_atoms, n_subunits, n_bonds;
  int32_t max_subunit_atoms;  /* largest subunit, in atoms                              */
  int32_t num_steps;          /* counts step 0, like the reference (num_steps-1 moves)   */
  int32_t io_dtype;           /* MCH_F64 / MCH_F32: element type of pos_in/pos_out/trajectory */
  int32_t compute_dtype;      /* MCH_F64: parity mode; MCH_F32: fast mode                */
  int32_t inexact_undo;       /* 1: on reject redo (p - v) + v like optimizer.py:273-277 */
  int32_t threads_per_chain;  /* 0 = auto; else multiple of 32, <= 1024 (fp32) / 512 (fp64) */
  int32_t cluster_size;       /* CTAs (SMs) per chain: 0 = auto, 1 = one CTA, 2 / 4 / 8 / 16 = thread-block cluster */
  int32_t resident_chains_hint; /* 0, or the number of chains of the whole batch when the caller cuts one batch
                                 into several launches that run concurrently (pipelined slices): block size,
                                 cluster and occupancy choices are then made for the batch, not per slice */
  const void* pos_in;         /* [dev] [n_chains or 1][n_atoms][3]                        */
  int64_t pos_in_chain_stride;/* elements between chains; 0 = all chains start from chain 0 */
  void* pos_out;              /* [dev] [n_chains][n_atoms][3]                             */
  const int32_t* perm;        /* [dev] [n_atoms]                                          */
  const int32_t* su_offsets;  /* [dev] [n_subunits + 1]                                   */
  const int32_t* bonds;       /* [dev] [n_bonds][2]                                       */
  const int32_t* bond_su;     /* [dev] [n_bonds][2]                                       */
  mch_params params;
  uint64_t* rng_state;        /* [dev] [n_chains][6] in/out                               */
  double* system_potential;   /* [dev] [n_chains] out                                     */
  double* nonbonded_potential;/* [dev] [n_chains] out                                     */
  double* max_bond_distance;  /* [dev] [n_chains] out                                     */
  int32_t* n_accepted;        /* [dev] [n_chains] out                                     */
  int32_t* last_step_info;    /* [dev] [n_chains] out: trace_info word of the last step   */
  double* trace_potentials;   /* [dev] [n_chains][num_steps][3] or NULL: U, U_nb, max bond */
  int32_t* trace_info;        /* [dev] [n_chains][num_steps] or NULL                      */
  void* trajectory;           /* [dev] [n_chains][num_steps][n_atoms][3] io_dtype or NULL */
  uint64_t* pair_evals;       /* [dev] [n_chains] or NULL: pair terms evaluated           */
  /* Explicit moving sets (optional; mchammer_b200.lowering builds them).  They are REQUIRED when
   * the caller's subunits overlap (`general` = 1: a move translates every atom the chosen subunit
   * LISTS, shared atoms included -- reference optimizer.py:248-252 -- while su_offsets/bond_su
   * keep the first-owner partition of :220-221), and they let molecules that do not fit one
   * SM's shared memory (or whose subunit count outgrows the S x S energy table) run with their
   * state split over a thread-block cluster instead of being refused with MCH_ERR_TOO_LARGE.
   * `general` = 2 forces that split-state kernel for any input (tests, measurements). */
  int32_t general;
  int32_t n_row_lists;        /* the caller's subunits (n_subunits minus the static group); 0 = lists absent */
  int32_t max_rows;           /* longest row list                                          */
  int32_t flags;              /* 0, or MCH_FLAG_* bits                                     */
  const int32_t* bond_slots;  /* [dev] [n_bonds][2] slot of each bond end                  */
  const int32_t* row_offsets; /* [dev] [n_row_lists + 1]                                   */
  const int32_t* row_slots;   /* [dev] slots of the atoms each subunit lists, ascending, each once */
  const int32_t* row_weights; /* [dev] how often the caller's list names that atom (centroid weights) */
  const int32_t* run_offsets; /* [dev] [n_row_lists + 1]                                   */
  const int32_t* runs;        /* [dev] [..][2] maximal contiguous slot ranges of each row list */
} mch_optimize_desc;

int mch_optimize_batch(const mch_optimize_desc* desc, void* stream);

/* Shared memory one chain needs (bytes), or MCH_ERR_TOO_LARGE. Host-only query. */
int mch_optimize_smem_bytes(int n_atoms, int n_subunits, int n_bonds, int max_subunit_atoms,
                            int compute_dtype);

/* ---------------------------------------------------------------------------------------
 * mch_potential_batch -- potential of n_mol molecules that share a bond list.
 * Stands in for Optimizer.compute_potential / _compute_nonbonded_potential
 * (optimizer.py:114-142) and the max-bond-distance scan (:174-183).
 * ------------------------------------------------------------------------------------- */
int mch_potential_batch(const void* pos /* [dev] [n_mol][n_atoms][3] */, int io_dtype,
                        int compute_dtype, int n_mol, int n_atoms,
                        const int32_t* bonds /* [dev] [n_bonds][2], may be NULL if 0 */,
                        int n_bonds, const mch_params* params,
                        double* system_potential /* [dev] [n_mol] */,
                        double* nonbonded_potential /* [dev] [n_mol] */,
                        double* max_bond_distance /* [dev] [n_mol] */, void* stream);

/* ---------------------------------------------------------------------------------------
 * mch_collapse_batch -- Collapser.get_result / get_trajectory (collapser.py:200-340) for
 * n_mol molecules of one topology.  Slot order additionally puts the non-H atoms of each
 * subunit first (n_heavy[s] of them); H atoms are ignored by the contact test (:77).
 * ------------------------------------------------------------------------------------- */
typedef struct mch_collapse_desc {
  int32_t n_mol, n_atoms, n_subunits, n_bonds;
  int32_t io_dtype, compute_dtype;
  int32_t scale_steps;        /* Collapser(scale_steps=...)                               */
  int32_t max_steps;          /* safety cap; status 1 when reached                        */
  int32_t trace_capacity;     /* steps the trace/trajectory buffers can hold              */
  int32_t threads_per_mol;    /* 0 = auto                                                 */
  int32_t n_moving_subunits;  /* the first n_moving_subunits subunits are the caller's dict entries; a trailing
                                 group (atoms in no subunit) never moves, is not contact-tested and does not
                                 enter the step scales (collapser.py:56-82, :106-157).  0 = all n_subunits */
  int32_t n_heavy_total;      /* sum of n_heavy[] (sizes a shared-memory buffer); 0 = unknown, n_atoms is assumed */
  double step_size, distance_threshold;
  const void* pos_in;         /* [dev] [n_mol or 1][n_atoms][3]                           */
  int64_t pos_in_mol_stride;
  void* pos_out;              /* [dev] [n_mol][n_atoms][3]                                */
  const int32_t* perm;        /* [dev] [n_atoms]                                          */
  const int32_t* su_offsets;  /* [dev] [n_subunits + 1]                                   */
  const int32_t* n_heavy;     /* [dev] [n_subunits]                                       */
  const int32_t* bonds;       /* [dev] [n_bonds][2]                                       */
  int32_t* steps_run;         /* [dev] [n_mol] out                                        */
  double* min_distance;       /* [dev] [n_mol] out: value tested against threshold/2      */
  double* max_bond_distance;  /* [dev] [n_mol] out: after the last step                   */
  int32_t* status;            /* [dev] [n_mol] out                                        */
  double* trace_max_bond;     /* [dev] [n_mol][trace_capacity] or NULL                    */
  void* trajectory;           /* [dev] [n_mol][trace_capacity][n_atoms][3] or NULL        */
} mch_collapse_desc;

int mch_collapse_batch(const mch_collapse_desc* desc, void* stream);

/* ---------------------------------------------------------------------------------------
 * Host-side random stream set-up (no GPU needed).
 * mch_pcg64_seed: state words of numpy.random.default_rng(seed) for each seed
 *   (SeedSequence -> PCG64, what optimizer.py:90-93 constructs).  out: [n][6].
 * mch_pcg64_stream_check: replays on the host the draw pattern of one Monte-Carlo step
 *   (choice(bound), choice(2), random(), choice(2), and a 5th random() every third round)
 *   with the same code the kernels use; used to test the stream model against numpy.
 * ------------------------------------------------------------------------------------- */
int mch_pcg64_seed(const uint64_t* seeds, int n, uint64_t* rng_state_out);
int mch_pcg64_stream_check(uint64_t* rng_state /* [6] in/out */, int rounds, uint32_t bound,
                           int64_t* ints_out /* [3*rounds] */,
                           double* doubles_out /* [rounds + (rounds+2)/3] */);

const char* mch_last_error(void);
int mch_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MCHAMMER_B200_H */
