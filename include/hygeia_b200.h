/* include/hygeia_b200.h -- C ABI of libhygeia_b200.so: the B200-native drop-in for Hygeia's inference hot path.
 *
 * Boundary being replaced (paths relative to /root/reference):
 *   single-group operator   runOnlineCombinedInferenceCpp     src/single_group/src/cpp/singleGroup.cpp:76-189
 *     (the Rcpp export the Rscript src/single_group/bin/estimate_parameters_and_regimes:303-322 calls)
 *   its model setup          setKnownParameters/UnknownParameters  src/single_group/src/cpp/singleGroup.h:173-335
 *   its emission term        evaluateLogObservationDensity    src/single_group/src/cpp/singleGroup.h:610-627
 *   prior draw               sampleFromParameterPriorCpp      src/single_group/src/cpp/singleGroup.cpp:18-35
 *
 * Conventions: C linkage, plain pointers and sizes, caller-owned buffers, `int` status (0 = ok, < 0 = error, message via
 * hyg_last_error).  No exceptions or exit() cross this boundary.  One context per device / host thread; contexts are
 * independent.  There is NO CPU fallback: every entry point that computes fails with HYG_ERR_CUDA when no device is usable.
 */
#ifndef HYGEIA_B200_H
#define HYGEIA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HYG_OK 0
#define HYG_ERR_ARG (-1)
#define HYG_ERR_UNSUPPORTED (-2)
#define HYG_ERR_CUDA (-3)
#define HYG_ERR_STATE (-4)
#define HYG_ERR_CAPACITY (-5)   /* a bounded workspace overflowed and results would differ from the reference (see status words) */

typedef struct hyg_ctx hyg_ctx;

/* ---- context ---------------------------------------------------------------------------------------------------- */
hyg_ctx* hyg_create(int device);          /* NULL if the device cannot be initialised (see hyg_create_error) */
void hyg_destroy(hyg_ctx* ctx);
const char* hyg_last_error(hyg_ctx* ctx); /* message of the last failing call on this context */
const char* hyg_create_error(void);       /* why the last hyg_create returned NULL */
const char* hyg_version(void);
void* hyg_stream(hyg_ctx* ctx);           /* the cudaStream_t every kernel of this context is launched on */

/* ---- model (= vartheta and theta of the reference) -------------------------------------------------------------- */
/* setKnownParameters (singleGroup.h:173-195): vartheta = (u, R, alpha[R], beta[R], isKappaFixed, kappa[R]) */
int hyg_sg_set_model(hyg_ctx* ctx, uint32_t R, uint32_t u, const double* alpha, const double* beta, int kappa_fixed, const double* kappa);
int hyg_sg_set_vartheta(hyg_ctx* ctx, const double* vartheta, uint32_t n);
/* setUnknownParameters (singleGroup.h:197-335): theta in R^(R*R) -> P, omega and the sojourn tables, built on the host in
 * the reference's order of operations and uploaded.  t_max bounds the table length (longest chain that will be run). */
int hyg_sg_set_theta(hyg_ctx* ctx, const double* theta, uint32_t dim, uint64_t t_max);
/* read back what the kernels will use: P[R*R] row-major, omega[R]; rho/exit for d = 1..d_max (each R x d_max, row-major) */
int hyg_sg_get_tables(hyg_ctx* ctx, double* P, double* omega, uint32_t d_max, double* rho, uint8_t* exit_status);

/* ---- data sets: one per chromosome / count matrix ----------------------------------------------------------------
 * Counts are uint16 (the reference itself narrows to int16: src/two_group/run_inference_two_groups.py:246-253), layout
 * [S][T] with the SITE index fastest.  `on_device` != 0: the pointers are device pointers with row pitch `pitch`
 * elements (pitch even, >= T, the pad readable) and are used in place; otherwise they are host pointers with row pitch
 * `pitch` elements (>= T; 0 = T -- a window of a wider host matrix can be staged without repacking) and are copied
 * host -> device by this call. */
int hyg_sg_add_dataset(hyg_ctx* ctx, uint64_t T, uint32_t S, const uint16_t* n_total, const uint16_t* n_meth, int on_device, uint64_t pitch);
int hyg_sg_clear(hyg_ctx* ctx);           /* drop all data sets and chains */

/* ---- chains: one per (data set, seed) --------------------------------------------------------------------------- */
#define HYG_SG_STATUS_WORDS 8
typedef struct hyg_sg_chain {
  uint32_t dataset;             /* index returned by the order of hyg_sg_add_dataset calls */
  uint64_t seed;                /* Philox key: u[t] = philox(seed, chain_id, t) ...                          */
  uint32_t chain_id;
  const double* uniforms;       /* ... unless T injected per-site uniforms are given (host pointer, may be NULL) */
  const uint32_t* positions;    /* T genomic positions (host) or NULL -> 0..T-1; first column of regime_probs */
  /* outputs, host pointers, any may be NULL */
  double* regime_probs;         /* T x (1+R) row-major: position, p_1..p_R  (= regimeProbabilityEstimates).  If this is
                                   pinned (page-locked) memory the kernel writes the rows straight into it as sites are
                                   finalised and hyg_sg_download only synchronises; pageable memory is staged + copied */
  double* logz;                 /* T : running log normalising constant after site t                          */
  double* theta_trace;          /* T x D (parameter-estimation mode)                                          */
  /* step-level taps for parity tests, host pointers, any may be NULL */
  int32_t* k_kept;              /* T : K of optimal resampling; -1 growth phase; -2 keep-largest              */
  uint8_t* drew_uniform;        /* T */
  int32_t* n_pending;           /* T : lag-set size after site t                                              */
  int32_t* n_curr;              /* T : particle count after site t                                            */
  int32_t* finalised_at;        /* T : step at which site t was emitted                                       */
  uint64_t* support_hash;       /* T : order-independent hash of the finite-weight support {(d, r)} after site t     */
  uint8_t* tie_flags;           /* T : bit 0 two of the sorted weights were exactly equal; bit 1 such a tie DECIDED which
                                       particle survived (the one place where this library and the reference, whose order
                                       of equal weights is whatever std::sort leaves, can differ: DESIGN.md quirk C-14)   */
  int32_t status[HYG_SG_STATUS_WORDS];
                                /* out: [0] forced emissions because the lag set was full, [1] max lag-set size,
                                        [2] owned sites emitted by force at the end of a segment's right halo (segmented
                                        execution only), [3] sites stepped through, halos included, [4] sites whose sort was
                                        redone on the full words (two weights equal in their top 56 bits), [5] sites where an exact tie of weights decided a
                                        particle's fate, [6] double systematic draws repaired, [7] segmented execution:
                                        number of overlap rows whose two computations differ by more than 1e-6              */
  double overlap_max_abs;       /* out: segmented execution: max |p - p'| over the rows a segment's right halo recomputed
                                        after the next segment (left-halo check); 0 for whole chains                        */
} hyg_sg_chain;

/* Algorithm switches: same meaning as the scalar arguments of runOnlineCombinedInferenceCpp (singleGroup.cpp:83-95). */
typedef struct hyg_sg_run_args {
  uint32_t n_particles_max;             /* <= 256 */
  uint32_t smc_proposal_type;           /* must be 1 (change-point model)            */
  uint32_t smc_resample_type;           /* must be 2 (optimal finite-state)          */
  int32_t use_online_marginal_smoothing;
  double epsilon;
  int32_t use_online_parameter_estimation;
  int32_t normalise_gradients;
  int32_t use_adam;
  uint32_t n_steps_without_parameter_update;
  double learning_rate_exponent;
  double learning_rate_factor;
  uint32_t lag_capacity;                /* pending-site capacity of the fixed-lag smoother per chain (0 -> 1024).  The reference's
                                           lag set is unbounded (OnlineMarginalSmoothing.h:148-195); here the rows live in a
                                           global-memory workspace (12 KB per pending site), so the capacity can be made as
                                           large as the longest lag.  A site that arrives when the set is full is emitted with
                                           its filtering estimate and COUNTED in status[0]; hyg_sg_download fails with
                                           HYG_ERR_CAPACITY after the launch unless allow_forced_emission is set.  The one-call
                                           operator hyg_sg_run_online_combined_inference then runs the recursion again with four
                                           times the capacity (same draws) until nothing overflowed                          */
  int32_t allow_forced_emission;        /* keep going (status[0] > 0) instead of failing when the lag set overflowed        */
  int32_t resample_full_sort;           /* sort the particles on the full (log-weight, regime, sojourn) words at every site instead of
                                           only where two weights agree in their top 56 bits (same decisions; parity tests)   */
} hyg_sg_run_args;

void hyg_sg_default_run_args(hyg_sg_run_args* args);

/* Segmented execution of the recursion (throughput mode).  segment_sites = 0 (default): every chain is one sequential run from
 * its first to its last site, exactly as OnlineCombinedInference::run (OnlineCombinedInference.h:48-118).  segment_sites > 0:
 * every chain is cut into near-equal segments of at most segment_sites sites which run concurrently; a segment starts
 * halo_left sites early from the R-particle initial system (Smc::initialise) -- the filter forgets its initial condition -- and
 * runs up to halo_right sites past its end, until its last site has been finalised by the fixed-lag smoother.  Uniforms stay
 * indexed by the site, and log Z_t is stitched from the per-segment increments.  The reference uses the same device for its
 * two-group path (segment_size 100000, buffer_size 5000: src/two_group/run_inference_two_groups.py:64-72,195-218).  Deviation
 * from the whole-chain run with the default halos: <= 3e-10 on the posteriors, 0 differing regime calls, <= 4e-8 absolute on
 * log Z increments (tools/segment_study.py; GPU tests).  Ignored in parameter-estimation mode (theta evolves along the chain). */
#define HYG_SEGMENT_AUTO UINT64_MAX  /* segment size chosen per launch so that the resident CTAs of the device finish together */
int hyg_sg_set_segmentation(hyg_ctx* ctx, uint64_t segment_sites, uint64_t halo_left, uint64_t halo_right);
/* Pinned regime_probs buffers are written by the kernel directly (default on); 0 forces device staging + a D2H copy. */
int hyg_sg_set_zero_copy_outputs(hyg_ctx* ctx, int enable);
/* Number of (chain, segment) units the last hyg_sg_filter launched, the segment size it used (0 = whole chains) and the
 * number of persistent CTAs that shared the units (any pointer may be NULL). */
int hyg_sg_filter_units(hyg_ctx* ctx, uint32_t* n_units, uint64_t* segment_sites, uint32_t* resident_ctas);

/* Stage the chains (uploads injected uniforms, allocates device outputs). */
int hyg_sg_set_chains(hyg_ctx* ctx, const hyg_sg_chain* chains, uint32_t n_chains);
/* K1: emission tables logObs[T x R] of every data set (device resident).  Asynchronous on the context's stream. */
int hyg_sg_emission(hyg_ctx* ctx);
/* K2: the recursion over all staged chains (device resident).  Asynchronous on the context's stream; the status words are
 * checked by hyg_sg_download (HYG_ERR_CAPACITY, see hyg_sg_run_args::lag_capacity). */
int hyg_sg_filter(hyg_ctx* ctx, const hyg_sg_run_args* args);
/* Device pointers of chain `chain`'s staged outputs, for consumers that stay on the device (e.g. an NCCL reduction of the
 * posteriors over seeds): regime_probs [T][1+R] (NULL when it was not requested or is written straight into pinned host
 * memory -- see hyg_sg_set_zero_copy_outputs), logz [T].  Valid until the next hyg_sg_set_chains / hyg_sg_clear. */
int hyg_sg_device_outputs(hyg_ctx* ctx, uint32_t chain, double** regime_probs, double** logz);
/* Device -> host copy of the outputs of every staged chain into the host pointers of hyg_sg_set_chains; synchronises. */
int hyg_sg_download(hyg_ctx* ctx, hyg_sg_chain* chains, uint32_t n_chains);
int hyg_sync(hyg_ctx* ctx);
/* Device time (ms, CUDA events on the context's stream) of the last hyg_sg_emission / hyg_sg_filter; synchronises. */
int hyg_sg_timings(hyg_ctx* ctx, float* ms_emission, float* ms_filter, uint32_t* emission_launches, uint32_t* filter_launches);
/* Copy the emission table of data set `dataset` to the host (T x R doubles). */
int hyg_sg_get_logobs(hyg_ctx* ctx, uint32_t dataset, double* logobs);

/* The operator itself: one chain, host buffers in, host buffers out -- argument for argument runOnlineCombinedInferenceCpp
 * (singleGroup.cpp:76-96) except that counts are uint16 [S][T] (site fastest) and the seed indexes Philox.
 * regime_probs: T x (1+R); theta_trace: T x D (parameter mode) or NULL; seconds: wall time of the call;
 * status: HYG_SG_STATUS_WORDS ints as in hyg_sg_chain (may be NULL). */
int hyg_sg_run_online_combined_inference(hyg_ctx* ctx, const double* vartheta, uint32_t n_vartheta, const double* theta_init, uint32_t dim_theta,
                                         uint64_t T, uint32_t S, const uint32_t* positions, const uint16_t* n_total, const uint16_t* n_meth,
                                         const hyg_sg_run_args* args, uint64_t seed, const double* uniforms,
                                         double* regime_probs, double* theta_trace, double* logz, double* seconds, int32_t* status);

/* ---- two-group (case/control) path -------------------------------------------------------------------------------
 * Boundary being replaced: hygeia/filter_and_smoother_algorithm.py::run as called by `hygeia infer`
 * (src/two_group/run_inference_two_groups.py:261-276; parameters :110-167, outputs :233-240,294-322).
 * Emission tables come from K1: add the control and the case count matrices as two data sets (hyg_sg_add_dataset, after
 * hyg_sg_set_model has fixed alpha/beta) and call hyg_sg_emission. */
typedef struct hyg_tg_model {
  uint32_t R;                      /* regimes                                                                    */
  uint32_t minimum_duration;       /* u (flag --minimum_duration, default 3)                                     */
  uint32_t num_resampled;          /* M ancestors kept per site (--num_resampled_particles, default 50; <= 64)   */
  uint32_t num_backward;           /* backward trajectories (--num_samples_backward, default 25; <= 32)          */
  const double* log_p_control;     /* R x R row-major log transition matrix of the control regimes (diagonal ignored) */
  const double* omega_control;     /* R                                                                          */
  const double* omega_case;        /* R (the CLI uses 0.8 for every regime)                                      */
  const double* kappa_control;     /* R (2.0)                                                                    */
  const double* kappa_case;        /* R (2.0)                                                                    */
  double merge_prob;               /* P(split -> merged), 0.1                                                    */
  double split_prob;               /* P(merged -> split), 0.01                                                   */
  /* optional precomputed hazards rho[r][d], d = 0..d_max, row pitch d_max + 1 (NULL -> computed here in fp64)   */
  const double* rho_control;
  const double* rho_case;
  uint32_t d_max;
  /* which hazard is computed when no tables are supplied: HYG_TG_HAZARD_REFERENCE (0, default) = as the reference evaluates
   * it, in fp32 with the fixed value 0.1 where that is not finite (case_control_regime_model.py:111-168: from d = 94 on for
   * omega = 0.8); HYG_TG_HAZARD_EXACT (1) = the negative-binomial hazard in fp64 */
  uint32_t hazard_mode;
  /* tuning / test hook: the resampling sort covers only the heaviest particles of a site -- at least sort_preselect[0] of them
   * on the first attempt, sort_preselect[1] on the second, all of them on the third; an attempt is repeated only when a tooth of
   * the systematic comb falls behind the sorted prefix, so the results never depend on these numbers.  0 -> M + 110 and
   * 3 M + 250; values below M + 96 are raised to it (the K search looks at the first M + 96 positions) */
  uint32_t sort_preselect[2];
  /* test hook: when at least this many particles are selected for the sort, it runs in the CTA's global scratch area instead of
   * shared memory (the path that more than 2048 finite particles of a site would take); 0 -> only then */
  uint32_t sort_scratch_from;
} hyg_tg_model;
#define HYG_TG_HAZARD_REFERENCE 0u
#define HYG_TG_HAZARD_EXACT 1u

typedef struct hyg_tg_chain {
  uint32_t control_dataset;        /* data-set indices (hyg_sg_add_dataset order); both must have the same T     */
  uint32_t case_dataset;
  uint64_t seed;
  uint32_t chain_id;
  /* outputs, host pointers */
  int32_t* trajectories;           /* T x B x 5 : merged, d_control, r_control, d_case, r_case.  Pinned (page-locked) host
                                      memory is written by the kernel directly (no staging copy), pageable memory by a copy */
  double* log_normalizing_constant;/* 1                                                                          */
  int32_t* taps;                   /* optional T x 4 : particles proposed, K, finite-weight particles, sort attempts (0 = no sort) */
} hyg_tg_chain;

int hyg_tg_set_model(hyg_ctx* ctx, const hyg_tg_model* model, uint64_t t_max);
/* K4/K5 over all chains: particle filter then backward simulation; synchronous (results are in the host buffers on return). */
int hyg_tg_run(hyg_ctx* ctx, const hyg_tg_chain* chains, uint32_t n_chains, float* ms_device);
/* host copy of the hazard tables the kernels use, rho[R][d_max + 1]: the exact one and the reference-mode one */
int hyg_tg_hazard_table(const double* omega, const double* kappa, uint32_t R, uint32_t u, uint32_t d_max, double* rho);
int hyg_tg_reference_hazard_table(const double* omega, const double* kappa, uint32_t R, uint32_t u, uint32_t d_max, double* rho);

/* ---- DMP calling -------------------------------------------------------------------------
 * Boundary being replaced: the reductions of `hygeia aggregate` and `hygeia get_dmps`
 * (src/two_group/aggregate_results.py:125-147,181; src/two_group/get_dmps.py:63-76,111-126) and the two procedures of
 * src/two_group/multiple_testing.py.  Inputs of hyg_tg_site_statistics are the aggregated trajectory matrices the reference
 * writes as merge_states_/control_regimes_/case_regimes_chrom_<chrom>.csv.gz: int8 [T][P], site-major, P = seeds x backward
 * trajectories.  Outputs (fp64, bit-identical to the NumPy expressions): split_prob[T] = mean(merged == 0);
 * null_stat[T] = 1 - #(control != case) / P; control_freq / case_freq [T][R] = bincount / P (may be NULL);
 * pair_stat[T][R][R] = 1 - #(control == i and case == j) / P (--test_regime_combinations; may be NULL).
 * on_device != 0: every pointer is a device pointer (inputs readable 16 bytes past their end) and nothing is copied. */
int hyg_tg_site_statistics(hyg_ctx* ctx, uint64_t T, uint32_t P, uint32_t R, const int8_t* merged, const int8_t* control_regimes,
                           const int8_t* case_regimes, int on_device, double* split_prob, double* null_stat, double* control_freq,
                           double* case_freq, double* pair_stat, float* ms_device);
/* FDR_procedure(test_statistics, fdr_threshold) -> (k, Qk, threshold)  (multiple_testing.py:3-12); host pointers */
int hyg_fdr_procedure(hyg_ctx* ctx, uint64_t n, const double* test_statistics, double fdr_threshold, uint64_t* k, double* Qk, double* threshold);
/* weighted_FDR_procedure(test_statistics, fdr_threshold, weights_false_positives, weights_false_negatives)
 * -> (ranking_indices[:s], Nsums[s-1])  (multiple_testing.py:13-22); `indices` has room for n entries; ties in the ranking
 * keep the smaller index first (NumPy leaves their order unspecified) */
int hyg_weighted_fdr_procedure(hyg_ctx* ctx, uint64_t n, const double* test_statistics, double fdr_threshold, const double* weights_false_positives,
                               const double* weights_false_negatives, uint64_t* n_selected, uint64_t* indices, double* Nk);

/* sampleFromParameterPriorCpp (singleGroup.cpp:18-35): theta ~ N(0, I_D), Philox-based (host). */
int hyg_sg_sample_theta_prior(uint32_t dim, uint64_t seed, double* theta);
/* Host copy of the by-site uniform the kernels use. */
double hyg_philox_uniform(uint64_t seed, uint32_t chain_id, uint64_t t);

#ifdef __cplusplus
}
#endif
#endif
