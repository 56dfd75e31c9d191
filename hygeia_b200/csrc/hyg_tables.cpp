// hygeia_b200/csrc/hyg_tables.cpp -- see hyg_tables.h.
#include "hyg_tables.h"

#include <cmath>
#include <limits>

namespace hyg {

static const double NEG_INF = -std::numeric_limits<double>::infinity();

// misc.h:630-640
double log_beta_binomial(uint32_t x, uint32_t range, double shape1, double shape2) {
  if (x <= range) {
    return std::lgamma(range + 1) - std::lgamma(x + 1) - std::lgamma(range - x + 1) + std::lgamma(x + shape1) +
           std::lgamma(range - x + shape2) - std::lgamma(range + shape1 + shape2) + std::lgamma(shape1 + shape2) -
           std::lgamma(shape1) - std::lgamma(shape2);
  }
  return NEG_INF;
}

void build_emission_table(const double* alpha, const double* beta, int R, int nmax, std::vector<double>& out) {
  out.assign(static_cast<size_t>(nmax + 1) * (nmax + 2) / 2 * R, 0.0);
  for (int n = 0; n <= nmax; n++)
    for (int x = 0; x <= n; x++)
      for (int r = 0; r < R; r++)
        out[(static_cast<size_t>(n) * (n + 1) / 2 + x) * R + r] = log_beta_binomial(x, n, alpha[r], beta[r]);
}

// misc.h:673-693
static double log_negative_binomial(uint32_t x, double size, double prob) {
  if (x == 0 && prob == 0) return 0.0;
  if (prob == 0) return NEG_INF;
  return std::lgamma(x + size) - std::lgamma(size) - std::lgamma(x + 1) + size * std::log(1 - prob) + x * std::log(prob);
}

int SgHostModel::set_known(const double* vt, uint32_t n) {
  if (n < 2) { err = "vartheta too short"; return -1; }
  const int u_ = static_cast<int>(vt[0]);
  const int R_ = static_cast<int>(vt[1]);
  if (R_ < 2 || R_ > 7) { err = "number of regimes must be in [2, 7]"; return -1; }
  if (n < static_cast<uint32_t>(2 * R_ + 3)) { err = "vartheta too short"; return -1; }
  const int kf = vt[2 * R_ + 2] != 0.0;
  if (kf && n < static_cast<uint32_t>(3 * R_ + 3)) { err = "vartheta lacks kappa"; return -1; }
  return set_known(R_, u_, vt + 2, vt + 2 + R_, kf, kf ? vt + 2 * R_ + 3 : nullptr);
}

int SgHostModel::set_known(int R_, int u_, const double* alpha_, const double* beta_, int kappa_fixed, const double* kappa_) {
  if (R_ < 2 || R_ > 7) { err = "number of regimes must be in [2, 7]"; return -1; }
  if (u_ < 2) { err = "u must be >= 2 (the reference indexes d-1 at d = u-1, singleGroup.h:309)"; return -1; }
  if (!kappa_fixed || !kappa_) {
    // The reference's kappa-estimation path reuses the omega index for the kappa gradient (singleGroup.h:664-668,329;
    // SURVEY.md C-7) and is never enabled by the Nextflow modules.
    err = "is_kappa_fixed = FALSE is not supported";
    return -2;
  }
  R = R_; u = u_; D = R * R;
  alpha.assign(alpha_, alpha_ + R);
  beta.assign(beta_, beta_ + R);
  kappa.assign(kappa_, kappa_ + R);
  return 0;
}

int SgHostModel::set_theta(const double* th, uint32_t dim, uint64_t t_max) {
  if (R == 0) { err = "set_model first"; return -1; }
  if (static_cast<int>(dim) != D) { err = "theta must have R*R entries"; return -3; }
  theta.assign(th, th + D);
  // P: row r = exp(normaliseExp(theta block r)) with a zero inserted at column r (singleGroup.h:201-212; misc.h:793-798)
  for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) { P[r][c] = 0.0; logP[r][c] = NEG_INF; }
  for (int r = 0; r < R; r++) {
    const double* blk = th + r * (R - 1);
    double mx = blk[0];
    for (int k = 1; k < R - 1; k++) if (blk[k] > mx) mx = blk[k];
    double sm = 0.0;
    for (int k = 0; k < R - 1; k++) sm += std::exp(blk[k] - mx);
    const double logz = mx + std::log(sm);
    int k = 0;
    for (int c = 0; c < R; c++) {
      if (c == r) continue;
      P[r][c] = std::exp(blk[k++] - logz);
      logP[r][c] = std::log(P[r][c]);
    }
  }
  omega.resize(R);
  for (int r = 0; r < R; r++) omega[r] = 1.0 / (1.0 + std::exp((-1.0) * th[R * (R - 1) + r]));  // misc.h:34-37

  // Sojourn tables by the reference's forward recurrence (singleGroup.h:289-332), fp64, same order.
  // The table stops where it has become stationary: h underflowed to exactly 0 (rho = 0 from then on) or the exit
  // flag is set (rho = 1 from then on); entry dcap-1 stands for every larger d.
  const uint64_t hard_cap = (t_max < 2) ? 2 : t_max;
  std::vector<std::vector<double> > rho(R), glr(R);
  std::vector<std::vector<char> > ex(R);
  uint64_t longest = 2;
  for (int r = 0; r < R; r++) {
    std::vector<double>& rh = rho[r];
    std::vector<double>& gl = glr[r];
    std::vector<char>& e = ex[r];
    double H_prev = 0.0, gH_prev = 0.0;
    bool exit_prev = false;
    for (uint64_t d = 0; d < hard_cap; d++) {  // d is the 0-based index; sojourn time = d + 1
      if (d + 1 < static_cast<uint64_t>(u)) { rh.push_back(0.0); gl.push_back(0.0); e.push_back(0); continue; }
      const uint32_t k = static_cast<uint32_t>(d + 1 - u);
      const double h = std::exp(log_negative_binomial(k, kappa[r], omega[r]));
      double H_here = 0.0, rho_d;
      char exit_d;
      double H_prev_used = H_prev;
      if (exit_prev || H_prev >= 1.0) {
        H_prev_used = 0.99999;  // the reference overwrites bigH[d-1] before the gradient lines read it
        rho_d = 1.0;
        exit_d = 1;
        H_here = 0.0;  // never written by the reference on this branch; overwritten with 0.99999 on the next d
      } else {
        H_here = H_prev + h;
        rho_d = h / (1.0 - H_prev);
        exit_d = 0;
      }
      const double glh = (static_cast<double>(k) / omega[r] - kappa[r] / (1.0 - omega[r])) * (2.0 + std::exp(-omega[r]) + std::exp(omega[r]));
      const double gH = gH_prev + h * glh;
      const double glr_d = glh + gH_prev / (1.0 - H_prev_used);
      rh.push_back(rho_d); gl.push_back(glr_d); e.push_back(exit_d);
      const bool stationary = (exit_d && exit_prev) || (!exit_d && h == 0.0 && k > 8 * (kappa[r] + 1.0) / (1.0 - omega[r] + 1e-300));
      H_prev = H_here; gH_prev = gH; exit_prev = exit_d;
      if (stationary) break;  // (the gradient table is not stationary, but it is only read in parameter mode, which builds its tables on the device)
    }
    if (rh.size() > longest) longest = rh.size();
  }
  dcap = static_cast<uint32_t>(longest);
  tab.assign(static_cast<size_t>(R) * dcap * 2, 0.0);
  tabg.assign(static_cast<size_t>(R) * dcap, 0.0);
  for (int r = 0; r < R; r++) {
    for (uint32_t d = 0; d < dcap; d++) {
      const size_t i = (d < rho[r].size()) ? d : rho[r].size() - 1;
      const double rh = rho[r][i];
      const bool e = ex[r][i] != 0;
      const bool can_change = (d + 1 >= static_cast<uint32_t>(u));
      // new segment (singleGroup.h:584-596): log rho + log P, or log P alone under the exit flag
      const double c_new = !can_change ? 0.0 : (e ? 1.0 : rh);
      // continuation (singleGroup.h:597-605)
      const double lc = (!e && rh <= 1) ? std::log(1.0 - rh) : NEG_INF;
      tab[(static_cast<size_t>(r) * dcap + d) * 2 + 0] = c_new;
      tab[(static_cast<size_t>(r) * dcap + d) * 2 + 1] = lc;
      tabg[static_cast<size_t>(r) * dcap + d] = glr[r][i];
    }
  }
  return 0;
}


void build_hazard_table(const double* omega, const double* kappa, int R, int u, uint32_t d_max, std::vector<double>& rho) {
  // Backward recurrence of the inverse hazard g(k) = P(X >= k) / pmf(k):  g(k) = 1 + omega (k + kappa)/(k + 1) g(k+1), with
  // g(inf) = 1/(1 - omega); started far enough beyond d_max that the start-up error (contracted by ~omega per step) is
  // below 1e-18.  No lgamma, no underflow, no cancellation.
  rho.assign(static_cast<size_t>(R) * (d_max + 1), 0.0);
  for (int r = 0; r < R; r++) {
    const double om = omega[r], ka = kappa[r];
    const long long n_extra = static_cast<long long>(std::min(std::ceil(-41.5 / std::log(om)), 5e7));
    double g = 1.0 / (1.0 - om);
    const long long kmax = static_cast<long long>(d_max) - u;
    for (long long k = kmax + n_extra - 1; k >= 0; k--) {
      g = 1.0 + (om * (static_cast<double>(k) + ka) / (static_cast<double>(k) + 1.0)) * g;
      if (k <= kmax) {
        const double v = 1.0 / g;
        rho[static_cast<size_t>(r) * (d_max + 1) + static_cast<size_t>(k + u)] = std::isfinite(v) ? v : 0.1;
      }
    }
  }
}

void build_reference_hazard_table(const double* omega, const double* kappa, int R, int u, uint32_t d_max, std::vector<double>& rho) {
  // The hazard AS THE REFERENCE EVALUATES IT (case_control_regime_model.py:111-168): exp(log_prob(d - u) -
  // log_survival_function(d - u - 1)) of tfd.NegativeBinomial(total_count = kappa, probs = omega) in fp32, replaced by the
  // fixed value 0.1 wherever that is not finite.  TFP's log survival function is log1p(-cdf); an fp32 cdf rounds to 1 once
  // 1 - cdf < 2^-25, so from that sojourn on the hazard is 0.1 whatever omega is (kappa = 2: d = 94 for omega = 0.8, 197 for
  // 0.9, 265 / 401 / 808 / 4071 for 0.925 / 0.95 / 0.975 / 0.995), and it carries fp32 cancellation noise before that.
  // Restated here with fp64 special functions rounded to fp32 at the points where TFP holds fp32 values.
  rho.assign(static_cast<size_t>(R) * (d_max + 1), 0.0);
  for (int r = 0; r < R; r++) {
    const float p32 = static_cast<float>(omega[r]);
    const double tc = static_cast<double>(static_cast<float>(kappa[r]));
    const double lg = static_cast<double>(static_cast<float>(std::log(static_cast<double>(p32)) - std::log1p(-static_cast<double>(p32))));
    const double softplus_pos = std::max(lg, 0.0) + std::log1p(std::exp(-std::fabs(lg)));      // log(1 + e^lg)
    const double softplus_neg = std::max(-lg, 0.0) + std::log1p(std::exp(-std::fabs(lg)));     // log(1 + e^-lg)
    const double w = 1.0 / (1.0 + std::exp(-lg));                                              // probability of one more site
    double pmf = std::exp(-tc * softplus_pos), cdf = 0.0;                                      // pmf(0) = (1 - w)^tc
    for (uint32_t d = static_cast<uint32_t>(u); d <= d_max; d++) {
      const double x = static_cast<double>(d) - u;
      const float logh = static_cast<float>(-tc * softplus_pos - x * softplus_neg
                                            - (std::lgamma(1.0 + x) + std::lgamma(tc) - std::lgamma(1.0 + x + tc)) - std::log(tc + x));
      float logsf = 0.0f;
      if (d > static_cast<uint32_t>(u)) {
        cdf += pmf;                                        // cdf(x - 1) = sum of pmf(0 .. x-1)
        pmf *= w * (x - 1.0 + tc) / x;                     // pmf(x) from pmf(x - 1)
        logsf = log1pf(-static_cast<float>(std::min(cdf, 1.0)));
      }
      const float v = expf(logh - logsf);
      rho[static_cast<size_t>(r) * (d_max + 1) + d] = std::isfinite(v) ? static_cast<double>(v) : static_cast<double>(0.1f);
    }
  }
}

}  // namespace hyg
