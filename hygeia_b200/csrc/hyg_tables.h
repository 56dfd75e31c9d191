// hygeia_b200/csrc/hyg_tables.h -- host-side model tables (plain C++, no CUDA).
//
// Everything here is evaluated on the HOST with libm, in the reference's own order of operations, and uploaded
// once per theta, so the constants the kernels consume are the reference's constants bit for bit
// (non-fast-math build).  Reference: /root/reference/src/single_group/src/cpp/singleGroup.h:173-335,
// misc/misc.h:630-640,673-693.
#ifndef HYG_TABLES_H
#define HYG_TABLES_H

#include <stdint.h>

#include <string>
#include <vector>

namespace hyg {

struct SgHostModel {
  int R = 0, u = 0, D = 0;
  std::vector<double> alpha, beta, kappa, omega, theta;
  double P[8][8];
  double logP[8][8];
  uint32_t dcap = 0;
  std::vector<double> tab;   // [R][dcap][2] = {c_new, lc}
  std::vector<double> tabg;  // [R][dcap]    = d log rho / d theta_omega
  std::string err;

  // vartheta = (u, R, alpha[R], beta[R], isKappaFixed, kappa[R])       (singleGroup.h:173-195)
  int set_known(const double* vartheta, uint32_t n);
  int set_known(int R_, int u_, const double* alpha_, const double* beta_, int kappa_fixed, const double* kappa_);
  // theta -> P, omega, sojourn tables for d = 1..dcap                  (singleGroup.h:197-335)
  int set_theta(const double* theta_, uint32_t dim, uint64_t t_max);
};

// Triangular emission table: out[((n*(n+1))/2 + x) * R + r] = logBetaBinomial(x; n, alpha_r, beta_r), n <= nmax
// (misc.h:630-640, the nine lgamma terms added left to right).
void build_emission_table(const double* alpha, const double* beta, int R, int nmax, std::vector<double>& out);
double log_beta_binomial(uint32_t x, uint32_t n, double a, double b);

// Two-group hazard rho[r][d] = pmf(d-u) / P(X >= d-u), X ~ NB(kappa_r, omega_r), d = 0..d_max (0 below u, 0.1 where not
// finite): what case_control_regime_model.py:111-168 computes through TFP's log_prob / log_survival_function.
void build_hazard_table(const double* omega, const double* kappa, int R, int u, uint32_t d_max, std::vector<double>& rho);
// the same table with the reference's fp32 evaluation and its fixed value 0.1 where that is not finite
void build_reference_hazard_table(const double* omega, const double* kappa, int R, int u, uint32_t d_max, std::vector<double>& rho);

}  // namespace hyg
#endif
