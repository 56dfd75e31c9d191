/* oracle/sg_oracle.h -- TEST INFRASTRUCTURE, not product code.
 *
 * C ABI of the CPU restatement of Hygeia's single-group inference path
 * (oracle/sg_oracle.cpp).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.
 *
 * Parity status: PINNED against the reference itself -- the reference's own
 * headers are compiled in this container (oracle/_ref, see oracle/Makefile and
 * oracle/ref_driver.cpp) and tests/test_oracle_vs_ref.py plus the committed
 * fixtures in tests/golden/ check this restatement against them.  The
 * reference repository ships no tests or golden vectors of its own
 * (SURVEY.md section 4).
 */
#ifndef HYG_SG_ORACLE_H
#define HYG_SG_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct hygo_sg_args {
  /* model: vartheta = (u, R, alpha[R], beta[R], isKappaFixed, kappa[R]); theta in R^(R*R) */
  const double* vartheta;
  uint32_t n_vartheta;
  const double* theta;
  uint32_t dim_theta;
  /* data */
  uint64_t T;
  uint32_t S;
  const uint32_t* positions;   /* T, may be NULL (then 0..T-1)          */
  const uint32_t* n_total;     /* S x T, sample index fastest            */
  const uint32_t* n_meth;      /* S x T, sample index fastest            */
  const double* logobs;        /* optional precomputed T x R table (then n_total/n_meth may be NULL) */
  /* algorithm switches (singleGroup.cpp:83-95) */
  uint32_t n_particles_max;
  int32_t use_smoothing;
  double epsilon;
  int32_t use_param_est;
  int32_t normalise_gradients;
  int32_t use_adam;
  uint32_t n_steps_without_update;
  double lr_exponent;
  double lr_factor;
  /* randomness: one uniform per site, used only when the resampler needs it */
  const double* uniforms_by_site; /* T, required */
  /* outputs (any may be NULL) */
  double* regime_probs;    /* T x (1+R): position, p_0..p_{R-1}                       */
  double* theta_trace;     /* T x D                                                   */
  double* logz;            /* T : log-sum of unnormalised weights after site t        */
  int32_t* n_curr;         /* T : particle count after site t                         */
  int32_t* k_kept;         /* T : K of optimal resampling; -1 growth, -2 keep-largest */
  int32_t* finalised_at;   /* T : step at which site t's estimate was emitted         */
  uint8_t* drew_uniform;   /* T                                                       */
  int32_t* n_pending;      /* T : lag-set size after site t                           */
  int16_t* ancestors;      /* T x (n_particles_max - R), -1 padded                    */
  double* seconds;
  int32_t* tie_pairs;      /* T : #{adjacent equal pairs among the finite weights in the sorted order used at site t}
                                  (0 when site t did not sort) -- where the reference's unstable sort decides     */
  double* weights_prev;    /* T x n_particles_max: self-normalised weights the resampler of site t saw (NaN padded) */
  int32_t* d_prev;         /* T x n_particles_max: sojourn d of the same particles (0 padded); regime in the top byte: d | r << 24 */
  /* Order of exactly equal weights in sort_index:
   *   0 = the reference's: libstdc++ std::sort with a value-only comparator (arrangement-dependent, == oracle/_ref);
   *   1 = canonical: log-weight descending, then (regime, sojourn) ascending -- a rule that does not depend on the storage
   *       order of the particles, which is what the CUDA path implements.  Both agree wherever no two weights are equal. */
  int32_t tie_order;
  uint64_t* support_hash;  /* T : order-independent hash of {(d, r)} over the finite-weight particles after site t */
  uint8_t* tie_flags;      /* T : bit 0 = two finite weights exactly equal in the sort of site t; bit 1 = such a tie DECIDED
                                  something (one of two tied particles survived the resampling, the other did not) */
} hygo_sg_args;

int hygo_sg_run(const hygo_sg_args* a);

/* logObs[T x R] (singleGroup.h:610-627 + misc.h:630-640) */
int hygo_sg_emission(const double* alpha, const double* beta, uint32_t R, uint64_t T, uint32_t S,
                     const uint32_t* n_total, const uint32_t* n_meth, double* logobs);

/* P[R x R], omega[R], and for d = 1..d_max: rho, exit flag, d(log rho)/d(theta_omega)  (singleGroup.h:197-335) */
int hygo_sg_tables(const double* vartheta, uint32_t n_vartheta, const double* theta, uint32_t dim_theta, uint32_t d_max,
                   double* rho, uint8_t* exit_status, double* grad_omega_log_rho, double* P, double* omega);

double hygo_log_beta_binomial(uint32_t x, uint32_t n, double a, double b);

#ifdef __cplusplus
}
#endif
#endif
