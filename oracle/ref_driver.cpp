// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Builds the *reference's own* single-group C++ (headers #include'd in place
// from /root/reference/src/single_group/src/cpp, never copied) against the
// stand-in oracle/shim/RcppArmadillo.h and exposes it behind a tiny C ABI so
// tests (ctypes) and bench.py's cpu_baseline / --impl reference legs can call
// it.  Output: oracle/_ref/libhyg_ref.so (git-ignored, travels with gpurun).
//
// What is restated here (glue only; the algorithm is the reference's code):
//   * the wiring of Model -> Smc -> OnlineMarginalSmoothing ->
//     OnlineParameterEstimation -> OnlineCombinedInference done by
//     runOnlineCombinedInferenceCpp            (singleGroup.cpp:104-178)
//     (that file itself needs Rcpp::List, so it cannot be compiled here);
//   * for the step-level variant, the ten-line outer loop of
//     OnlineCombinedInference::run             (OnlineCombinedInference.h:58-117)
//     so that log Z_t, K, the ancestors and the lag set can be observed after
//     every site.  hygref_sg_run(stepwise=0) calls the reference's run()
//     untouched; tests check that both variants return identical posteriors.
//
// Uniform draws: arma::randu() is routed to a hook.  If `uniforms_by_site` is
// given, the hook returns uniforms_by_site[smc.getStep()] (SURVEY.md fact 6:
// injected draws must be indexed by site, not by position in a stream);
// otherwise it draws from a std::mt19937_64 seeded with `rng_seed`.
#include <cstdint>
#include <cstring>
#include <ctime>
#include <random>

#include <RcppArmadillo.h>
namespace arma {
std::function<double()> g_randu;
std::function<double()> g_randn;
}  // namespace arma

#define private public
#include "singleGroup.h"
#undef private

typedef Model<ModelParameters, LatentVariable, Covariate, Observation> RefModel;
typedef Smc<ModelParameters, LatentVariable, Covariate, Observation, SmcParameters, Particle<LatentVariable> > RefSmc;
typedef OnlineMarginalSmoothing<ModelParameters, LatentVariable, Covariate, Observation, SmcParameters, Particle<LatentVariable> > RefOms;
typedef OnlineParameterEstimation<ModelParameters, LatentVariable, Covariate, Observation, SmcParameters, Particle<LatentVariable> > RefOpe;
typedef OnlineCombinedInference<ModelParameters, LatentVariable, Covariate, Observation, SmcParameters, Particle<LatentVariable> > RefOci;

namespace {

arma::colvec to_col(const double* p, uint64_t n) {
  arma::colvec v(n);
  for (uint64_t i = 0; i < n; i++) v(i) = p[i];
  return v;
}

struct Inputs {
  arma::uvec positions;
  arma::umat n_total, n_meth;  // S x T, column-major => sample fastest
};

Inputs make_inputs(uint64_t T, uint32_t S, const uint32_t* positions, const uint32_t* n_total, const uint32_t* n_meth) {
  Inputs in;
  in.positions.set_size(T);
  in.n_total.set_size(S, T);
  in.n_meth.set_size(S, T);
  for (uint64_t t = 0; t < T; t++) in.positions(t) = positions ? positions[t] : t;
  for (uint64_t i = 0; i < T * S; i++) {
    in.n_total.mem[i] = n_total[i];
    in.n_meth.mem[i] = n_meth[i];
  }
  return in;
}

}  // namespace

extern "C" {

struct hygref_sg_args {
  // model (= vartheta of model_functions.R:44-54, and thetaInit)
  const double* vartheta;
  uint32_t n_vartheta;
  const double* theta;
  uint32_t dim_theta;
  // data, S x T with the sample index fastest (arma::umat column-major)
  uint64_t T;
  uint32_t S;
  const uint32_t* positions;
  const uint32_t* n_total;
  const uint32_t* n_meth;
  // algorithm switches, same meaning/order as singleGroup.cpp:83-95
  uint32_t n_particles_max;
  uint32_t smc_proposal_type;
  uint32_t smc_resample_type;
  int32_t use_smoothing;
  double epsilon;
  int32_t use_param_est;
  int32_t normalise_gradients;
  int32_t use_adam;
  uint32_t n_steps_without_update;
  double lr_exponent;
  double lr_factor;
  // randomness
  const double* uniforms_by_site;  // T doubles or NULL
  uint64_t rng_seed;
  // 0: call the reference's OnlineCombinedInference::run();  1: restated outer loop with per-step taps
  int32_t stepwise;
  // outputs (any may be NULL)
  double* regime_probs;   // T x (1+R), row-major (position first)
  double* theta_trace;    // T x D
  double* logz;           // T                 (stepwise only)
  int32_t* n_curr;        // T                 (stepwise only) particle count after step t
  int32_t* finalised_at;  // T                 (stepwise only) step at which site t was finalised
  uint8_t* drew_uniform;  // T                 1 if randu() was consumed at site t
  int32_t* n_pending;     // T                 (stepwise only) lag-set size after step t
  double* seconds;        // 1
  // optional particle-system dump after step `dump_at` (stepwise only; -1 = off)
  int64_t dump_at;
  double* dump_logw;      // n_particles_max
  double* dump_W;         // n_particles_max
  uint32_t* dump_d;       // n_particles_max
  uint32_t* dump_r;       // n_particles_max
};

int hygref_sg_run(const hygref_sg_args* a) {
  const uint64_t T = a->T;
  std::mt19937_64 eng(a->rng_seed);
  std::uniform_real_distribution<double> unif(0.0, 1.0);
  std::normal_distribution<double> norm(0.0, 1.0);

  Rng rng;
  rng.setSeed(static_cast<unsigned int>(a->rng_seed));
  RefModel model(rng);
  arma::colvec vartheta = to_col(a->vartheta, a->n_vartheta);
  arma::colvec theta = to_col(a->theta, a->dim_theta);
  Inputs in = make_inputs(T, a->S, a->positions, a->n_total, a->n_meth);

  clock_t t1 = clock();
  // --- singleGroup.cpp:116-120 ---
  model.setKnownParameters(vartheta);
  model.setCovariates(convertArmaUmatToCovariates(in.positions, in.n_total));
  model.setObservations(convertArmaUmatToObservations(in.n_meth));
  model.setUnknownParameters(theta);
  // --- singleGroup.cpp:126-130 ---
  RefSmc smc(rng, model);
  smc.setSmcResampleType(static_cast<SmcResampleType>(a->smc_resample_type));
  smc.setSmcProposalType(static_cast<SmcProposalType>(a->smc_proposal_type));
  smc.setNParticlesMax(a->n_particles_max);
  smc.setNRegimes(model.getModelParameters().getNMethylationRegimes());
  // --- singleGroup.cpp:136-142 ---
  RefOms oms(rng, model, smc);
  if (a->use_smoothing) oms.setEpsilon(a->epsilon);
  // --- singleGroup.cpp:148-160 ---
  GradientAscent ga;
  RefOpe ope(rng, model, smc, ga);
  if (a->use_param_est) {
    ga.setNormaliseGradients(a->normalise_gradients != 0);
    ga.setUseAdam(a->use_adam != 0);
    ga.setLearningRateExponent(a->lr_exponent);
    ga.setLearningRateFactor(a->lr_factor);
    ope.setTheta(theta);
    ope.setNStepsWithoutParameterUpdate(a->n_steps_without_update);
  }
  // --- singleGroup.cpp:166-170 ---
  RefOci oci(rng, model, smc, oms, ope);
  oci.setNSteps(model.getNObservations());
  oci.setUseOnlineMarginalSmoothing(a->use_smoothing != 0);
  oci.setUseOnlineParameterEstimation(a->use_param_est != 0);

  if (a->drew_uniform) std::memset(a->drew_uniform, 0, T);
  arma::g_randu = [&]() -> double {
    uint64_t t = smc.getStep();
    if (a->drew_uniform && t < T) a->drew_uniform[t] = 1;
    if (a->uniforms_by_site) return a->uniforms_by_site[t < T ? t : T - 1];
    return unif(eng);
  };
  arma::g_randn = [&]() -> double { return norm(eng); };

  std::vector<arma::colvec> regime_est, theta_est;
  const uint32_t R = model.getModelParameters().getNMethylationRegimes();

  if (!a->stepwise) {
    oci.run(regime_est, theta_est);  // the reference's own loop, untouched
  } else {
    // --- OnlineCombinedInference.h:55-101, restated so that we can tap per-step state ---
    std::vector<arma::colvec> fe_aux;
    std::vector<unsigned int> ti_aux;
    std::vector<int32_t> fin_step;
    smc.initialise();
    if (a->use_smoothing) { fe_aux.reserve(T); ti_aux.reserve(T); oms.initialise(fe_aux, ti_aux); }
    if (a->use_param_est) { theta_est.reserve(T); ope.initialise(theta_est); }
    auto tap = [&](uint64_t t) {
      if (a->logz) a->logz[t] = smc.getLogSumOfUnnormalisedWeightsCurr();
      if (a->n_curr) a->n_curr[t] = static_cast<int32_t>(smc.getNParticlesCurr());
      if (a->n_pending) a->n_pending[t] = static_cast<int32_t>(oms.psiTimeIndices_.size());
      while (fin_step.size() < ti_aux.size()) fin_step.push_back(static_cast<int32_t>(t));
      if (a->dump_at >= 0 && static_cast<uint64_t>(a->dump_at) == t) {
        for (uint32_t n = 0; n < smc.getNParticlesCurr(); n++) {
          if (a->dump_logw) a->dump_logw[n] = smc.getLogUnnormalisedWeightsCurr()(n);
          if (a->dump_W) a->dump_W[n] = smc.getSelfNormalisedWeightsCurr()(n);
          if (a->dump_d) a->dump_d[n] = smc.getParticlesCurr(n).getDistance();
          if (a->dump_r) a->dump_r[n] = smc.getParticlesCurr(n).getRegime();
        }
      }
    };
    tap(0);
    for (uint64_t t = 1; t < T; t++) {
      smc.iterate();
      smc.evaluateBackwardKernels();
      if (a->use_smoothing) {
        if (t == T - 1) oms.setIsFinalStep(true);
        oms.update(fe_aux, ti_aux);
      }
      if (a->use_param_est) ope.update(theta_est);
      tap(t);
    }
    // --- OnlineCombinedInference.h:106-117 ---
    if (a->use_smoothing) {
      regime_est.resize(T);
      for (uint64_t s = 0; s < ti_aux.size(); s++) {
        unsigned int t = ti_aux[s];
        regime_est[t].set_size(fe_aux[s].size() + 1);
        regime_est[t](0) = model.getCovariates(t).getGenomicPosition();
        regime_est[t].subvec(1, regime_est[t].size() - 1) = fe_aux[s];
        if (a->finalised_at) a->finalised_at[t] = fin_step[s];
      }
    }
  }
  clock_t t2 = clock();
  if (a->seconds) *a->seconds = (static_cast<double>(t2) - static_cast<double>(t1)) / CLOCKS_PER_SEC;

  if (a->regime_probs && a->use_smoothing) {
    for (uint64_t t = 0; t < T; t++)
      for (uint32_t k = 0; k < 1 + R; k++)
        a->regime_probs[t * (1 + R) + k] = (regime_est[t].size() == 1 + R) ? regime_est[t](k) : std::numeric_limits<double>::quiet_NaN();
  }
  if (a->theta_trace && a->use_param_est) {
    for (uint64_t t = 0; t < theta_est.size() && t < T; t++)
      for (uint32_t k = 0; k < a->dim_theta; k++) a->theta_trace[t * a->dim_theta + k] = theta_est[t](k);
  }
  arma::g_randu = nullptr;
  arma::g_randn = nullptr;
  return 0;
}

// Emission table logObs[T x R] straight from Model::evaluateLogObservationDensity (singleGroup.h:610-627).
int hygref_sg_emission(const double* vartheta, uint32_t n_vartheta, uint64_t T, uint32_t S,
                       const uint32_t* n_total, const uint32_t* n_meth, double* logobs) {
  Rng rng;
  RefModel model(rng);
  model.setKnownParameters(to_col(vartheta, n_vartheta));
  Inputs in = make_inputs(T, S, nullptr, n_total, n_meth);
  model.setCovariates(convertArmaUmatToCovariates(in.positions, in.n_total));
  model.setObservations(convertArmaUmatToObservations(in.n_meth));
  const uint32_t R = model.getModelParameters().getNMethylationRegimes();
  LatentVariable lv;
  lv.setSojournTime(1);
  for (uint64_t t = 0; t < T; t++)
    for (uint32_t r = 0; r < R; r++) {
      lv.setMethylationRegimeType(r);
      logobs[t * R + r] = model.evaluateLogObservationDensity(t, lv);
    }
  return 0;
}

// Sojourn tables from ModelParameters (singleGroup.h:118-150,197-335): rho, exit flag and d/dtheta_omega log rho
// for d = 1..d_max (0-based output index d-1), each R x d_max row-major.
int hygref_sg_tables(const double* vartheta, uint32_t n_vartheta, const double* theta, uint32_t dim_theta, uint32_t d_max,
                     double* rho, uint8_t* exit_status, double* grad_omega_log_rho, double* P, double* omega) {
  Rng rng;
  RefModel model(rng);
  model.setKnownParameters(to_col(vartheta, n_vartheta));
  model.setUnknownParameters(to_col(theta, dim_theta));
  ModelParameters& mp = model.modelParameters_;
  const uint32_t R = mp.getNMethylationRegimes();
  for (uint32_t r = 0; r < R; r++) {
    for (uint32_t d = 1; d <= d_max; d++) {
      // mimic the access pattern of the filter: d grows by one per site
      double v = mp.getRho(d, r);
      if (rho) rho[r * d_max + d - 1] = v;
      if (exit_status) exit_status[r * d_max + d - 1] = mp.getExitStatus(d, r) ? 1 : 0;
      if (grad_omega_log_rho) grad_omega_log_rho[r * d_max + d - 1] = mp.getGradThetaOmegaLogRho(d, r);
    }
  }
  if (P) for (uint32_t i = 0; i < R; i++) for (uint32_t j = 0; j < R; j++) P[i * R + j] = mp.getP()(i, j);
  if (omega) for (uint32_t i = 0; i < R; i++) omega[i] = mp.getOmega(i);
  return 0;
}

// Beta-binomial log-density, misc.h:630-640.
double hygref_log_beta_binomial(uint32_t x, uint32_t n, double a, double b) { return evaluateLogBetaBinomialDensity(x, n, a, b); }

// Data simulated by the reference's own generative model (Model.h:62-75; singleGroup.h:485-557).
// n_total is given (S x T); writes n_meth (S x T) and latent (2 x T: sojourn, regime).
int hygref_sg_simulate(const double* vartheta, uint32_t n_vartheta, const double* theta, uint32_t dim_theta, uint64_t T, uint32_t S,
                       const uint32_t* n_total, uint64_t seed, uint32_t* n_meth, uint32_t* latent) {
  std::mt19937_64 eng(seed);
  std::uniform_real_distribution<double> unif(0.0, 1.0);
  arma::g_randu = [&]() -> double { return unif(eng); };
  Rng rng;
  rng.setSeed(static_cast<unsigned int>(seed));
  RefModel model(rng);
  model.setKnownParameters(to_col(vartheta, n_vartheta));
  model.setUnknownParameters(to_col(theta, dim_theta));
  std::vector<uint32_t> zeros(T * S, 0);
  Inputs in = make_inputs(T, S, nullptr, n_total, zeros.data());
  model.setCovariates(convertArmaUmatToCovariates(in.positions, in.n_total));
  model.simulateData(T);
  for (uint64_t t = 0; t < T; t++) {
    for (uint32_t s = 0; s < S; s++) n_meth[t * S + s] = static_cast<uint32_t>(model.getObservations(t)(s));
    if (latent) {
      latent[2 * t] = model.getLatentVariables(t).getSojournTime();
      latent[2 * t + 1] = model.getLatentVariables(t).getMethylationRegimeType();
    }
  }
  arma::g_randu = nullptr;
  return 0;
}

}  // extern "C"
