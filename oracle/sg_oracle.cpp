// oracle/sg_oracle.cpp -- TEST INFRASTRUCTURE, not product code (see sg_oracle.h).
//
// A plain-loop CPU restatement of Hygeia's single-group inference path:
// discrete particle filter over (sojourn d, regime r) with optimal finite-state
// resampling, forward-only adaptive fixed-lag smoothing and online score-based
// parameter estimation.  Every function cites the reference lines it follows
// (paths relative to /root/reference/src/single_group/src/cpp).  The order of
// floating-point operations follows the reference so that, built without
// -ffast-math, it reproduces oracle/_ref/libhyg_ref_strict.so bit for bit
// (tests/test_oracle_vs_ref.py).  Two deliberate differences, both neutral for
// the results: the emission table logObs[T x R] is evaluated once per
// (site, regime) instead of 1744 times per site.  sort_index
// takes the reference's own permutation under exact ties (see sort_index_desc).
#include "sg_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <ctime>
#include <limits>
#include <vector>

namespace {

const double NEG_INF = -std::numeric_limits<double>::infinity();

// misc.h:630-640
double log_beta_binomial(uint32_t x, uint32_t range, double shape1, double shape2) {
  if (x <= range) {
    return std::lgamma(range + 1) - std::lgamma(x + 1) - std::lgamma(range - x + 1) + std::lgamma(x + shape1) +
           std::lgamma(range - x + shape2) - std::lgamma(range + shape1 + shape2) + std::lgamma(shape1 + shape2) -
           std::lgamma(shape1) - std::lgamma(shape2);
  }
  return NEG_INF;
}

// misc.h:673-693
double log_negative_binomial(uint32_t x, double size, double prob) {
  if (x == 0 && prob == 0) return 0.0;
  if (prob == 0) return NEG_INF;
  return std::lgamma(x + size) - std::lgamma(size) - std::lgamma(x + 1) + size * std::log(1 - prob) + x * std::log(prob);
}

// misc.h:793-798 (+ the two-argument overload :800-806)
double normalise_exp(std::vector<double>& v, size_t n) {  // in place: v <- v - logZ, returns logZ
  double mx = v[0];
  for (size_t i = 1; i < n; i++) if (v[i] > mx) mx = v[i];
  double s = 0.0;
  for (size_t i = 0; i < n; i++) s += std::exp(v[i] - mx);
  double logz = mx + std::log(s);
  for (size_t i = 0; i < n; i++) v[i] = v[i] - logz;
  return logz;
}

// misc.h:748-760
double sum_exp(const std::vector<double>& v, size_t n) {
  double mx = v[0];
  for (size_t i = 1; i < n; i++) if (v[i] > mx) mx = v[i];
  if (mx > NEG_INF) {
    double s = 0.0;
    for (size_t i = 0; i < n; i++) s += std::exp(v[i] - mx);
    return mx + std::log(s);
  }
  return NEG_INF;
}

// ModelParameters, singleGroup.h:69-455 (kappa fixed)
struct Params {
  uint32_t u, R, D;
  std::vector<double> alpha, beta, kappa, omega;
  std::vector<double> P;  // R x R row-major, zero diagonal
  // sojourn tables, per regime, 0-based index d-1
  std::vector<std::vector<double> > h, H, rho, glh, gH, glr;
  std::vector<std::vector<char> > exit_status;
  std::vector<uint32_t> d_max;

  // singleGroup.h:173-195
  int set_known(const double* vt, uint32_t n) {
    u = static_cast<uint32_t>(vt[0]);
    R = static_cast<uint32_t>(vt[1]);
    if (n < 2 * R + 3) return -1;
    alpha.assign(vt + 2, vt + 2 + R);
    beta.assign(vt + 2 + R, vt + 2 + 2 * R);
    bool kappa_fixed = vt[2 * R + 2] != 0.0;
    if (!kappa_fixed || n < 3 * R + 3) return -2;  // kappa estimation: reference path is broken (SURVEY C-7); unsupported
    kappa.assign(vt + 2 * R + 3, vt + 3 * R + 3);
    D = R * R;
    return 0;
  }
  // singleGroup.h:197-269
  void set_unknown(const double* theta) {
    P.assign(R * R, 0.0);
    std::vector<double> aux(R - 1);
    for (uint32_t r = 0; r < R; r++) {
      for (uint32_t k = 0; k < R - 1; k++) aux[k] = theta[r * (R - 1) + k];
      normalise_exp(aux, R - 1);
      uint32_t k = 0;
      for (uint32_t c = 0; c < R; c++) P[r * R + c] = (c == r) ? 0.0 : std::exp(aux[k++]);
    }
    omega.resize(R);
    for (uint32_t r = 0; r < R; r++) omega[r] = 1.0 / (1.0 + std::exp((-1.0) * theta[R * (R - 1) + r]));  // misc.h:34-37
    h.assign(R, std::vector<double>());
    H = rho = glh = gH = glr = h;
    exit_status.assign(R, std::vector<char>());
    d_max.assign(R, 10);
    for (uint32_t r = 0; r < R; r++) extend(d_max[r], r);
  }
  // singleGroup.h:271-335.  The reference recomputes from d = 0 on every extension (its vectors are
  // reserve()d, never resize()d, so size() == 0); the values are a pure function of (r, d).
  void extend(uint32_t d_new, uint32_t r) {
    h[r].assign(d_new, 0.0); H[r].assign(d_new, 0.0); rho[r].assign(d_new, 0.0);
    glh[r].assign(d_new, 0.0); gH[r].assign(d_new, 0.0); glr[r].assign(d_new, 0.0);
    exit_status[r].assign(d_new, 0);
    for (uint32_t d = u - 1; d < d_new; d++) {
      h[r][d] = std::exp(log_negative_binomial(d + 1 - u, kappa[r], omega[r]));
      if (exit_status[r][d - 1] || H[r][d - 1] >= 1.0) {
        H[r][d - 1] = 0.99999;
        rho[r][d] = 1.0;
        exit_status[r][d] = 1;
      } else {
        H[r][d] = H[r][d - 1] + h[r][d];
        rho[r][d] = h[r][d] / (1.0 - H[r][d - 1]);
        exit_status[r][d] = 0;
      }
      // misc.h:92-95: gradLogitEvaluatedAtInverseLogit(omega) = 2 + exp(-omega) + exp(omega)
      glh[r][d] = (static_cast<double>(d + 1 - u) / omega[r] - kappa[r] / (1.0 - omega[r])) * (2.0 + std::exp(-omega[r]) + std::exp(omega[r]));
      gH[r][d] = gH[r][d - 1] + h[r][d] * glh[r][d];
      glr[r][d] = glh[r][d] + gH[r][d - 1] / (1.0 - H[r][d - 1]);
    }
    d_max[r] = d_new;
  }
  void need(uint32_t d, uint32_t r) {
    if (d > d_max[r]) {
      uint32_t nd = d_max[r];
      while (nd < d) nd = nd * 2;
      extend(nd, r);
    }
  }
  double get_rho(uint32_t d, uint32_t r) { need(d, r); return rho[r][d - 1]; }        // singleGroup.h:118-125
  double get_glr(uint32_t d, uint32_t r) { need(d, r); return glr[r][d - 1]; }        // singleGroup.h:128-135
  bool get_exit(uint32_t d, uint32_t r) { need(d, r); return exit_status[r][d - 1]; } // singleGroup.h:147-150

  // singleGroup.h:568-608
  double log_trans(uint32_t d_curr, uint32_t r_curr, uint32_t d_prev, uint32_t r_prev) {
    double ld = NEG_INF;
    if (d_curr == 1 && r_curr != r_prev && d_prev >= u) {
      double rh = get_rho(d_prev, r_prev);
      if (get_exit(d_prev, r_prev)) ld = std::log(P[r_prev * R + r_curr]);
      else ld = std::log(rh) + std::log(P[r_prev * R + r_curr]);
    } else if (d_curr > 1 && r_curr == r_prev) {
      double rh = get_rho(d_prev, r_prev);
      if (!get_exit(d_prev, r_prev) && rh <= 1) ld = std::log(1.0 - rh);
    }
    return ld;
  }
  // singleGroup.h:640-706 (kappa fixed): writes D entries into g
  void grad_log_trans(double* g, uint32_t d_curr, uint32_t r_curr, uint32_t d_prev, uint32_t r_prev) {
    for (uint32_t i = 0; i < D; i++) g[i] = 0.0;
    uint32_t idx = R * (R - 1) + r_prev;
    g[idx] = get_glr(d_prev, r_prev);
    if (d_curr == 1 && r_curr != r_prev && d_prev >= u) {
      uint32_t k = r_prev * (R - 1);
      for (uint32_t c = 0; c < R; c++) {
        if (c == r_prev) continue;
        double a = (-1.0) * P[r_prev * R + c];
        if (c == r_curr) a = a + 1;
        g[k++] = a;
      }
    } else if (d_curr > 1 && r_curr == r_prev) {
      double rh = get_rho(d_prev, r_prev);
      if (!get_exit(d_prev, r_prev) && rh < 1.0) g[idx] = -g[idx] * rh / (1.0 - rh);
      else for (uint32_t i = 0; i < D; i++) g[i] = 0.0;
    } else {
      for (uint32_t i = 0; i < D; i++) g[i] = 0.0;
    }
  }
};

struct Particle { uint32_t d, r; };

// resample.h:85-117 with T = (linspace(0,N-1,N) + u) / N
void systematic_base(double u, std::vector<uint32_t>& parent, const std::vector<double>& w, uint32_t N) {
  std::vector<double> Tq(N), Q(w.size());
  for (uint32_t j = 0; j < N; j++) Tq[j] = (static_cast<double>(j) + u) / N;
  double acc = 0.0;
  for (size_t i = 0; i < w.size(); i++) { acc += w[i]; Q[i] = acc; }
  uint32_t i = 0, j = 0;
  while (j < N) {
    if (i >= Q.size()) { parent[j] = static_cast<uint32_t>(Q.size() - 1); ++j; continue; }  // reference would read out of bounds; clamp
    if (Tq[j] <= Q[i]) { parent[j] = i; ++j; } else { ++i; }
  }
}

// arma::sort_index(v, "descend") (resample.h:304,373; Smc.h:438).  Armadillo packs (value, index) pairs and calls
// std::sort with a comparator that looks at the VALUE ONLY, so the order of exactly equal weights is whatever libstdc++'s
// introsort leaves behind -- a pure function of the arrangement, but neither "by index" nor stable.  Equal weights are
// systematic here: every resampled particle gets the common weight logSum - logC (resample.h:361-364), and
// log(1 - rho(d, r)) stops depending on d once the sojourn table has gone stationary (SURVEY C-4).  Which of two tied
// particles survives the next systematic resampling moves the smoothed posteriors by up to 1e-4 on sparse one-sample
// chains, so the restatement has to take the same permutation: the same std::sort call on the
// same pair type with the same comparator as oracle/shim/RcppArmadillo.h (the stand-in the reference is compiled against).
std::vector<uint32_t> sort_index_desc(const std::vector<double>& v, size_t n) {
  std::vector<std::pair<double, uint32_t> > pk(n);
  for (size_t i = 0; i < n; i++) pk[i] = std::make_pair(v[i], static_cast<uint32_t>(i));
  std::sort(pk.begin(), pk.end(), [](const std::pair<double, uint32_t>& a, const std::pair<double, uint32_t>& b) { return a.first > b.first; });
  std::vector<uint32_t> idx(n);
  for (size_t i = 0; i < n; i++) idx[i] = pk[i].second;
  return idx;
}

// order-independent hash of a particle's support point (tap; splitmix64 finaliser)
inline uint64_t mix64(uint64_t x) {
  uint64_t z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

struct Filter {
  Params mp;
  uint32_t R, Nmax, N_curr, N_prev;
  std::vector<Particle> pc, pp;
  std::vector<double> lw_c, lw_p, W_c, W_p;
  std::vector<uint32_t> anc;
  double lsum_c, lsum_p;
  std::vector<std::vector<double> > bk;  // R x N_prev
  const double* logobs;                  // T x R
  const double* unif;
  uint64_t step;
  int k_last;
  bool drew;
  int tie_order;  // 0 = reference (std::sort), 1 = canonical (log-weight desc, then regime, then sojourn)
  int tie_flags;  // tap: bit 0 = exact tie among finite weights in this site's sort, bit 1 = a tie decided a particle's fate
  // Canonical order: the key is the LOG-weight in both branches (W = exp(logw - logsum) is a monotone image of it; where W
  // collapses distinct log-weights -- underflow to 0 -- the reference's own keep-largest branch sorts log-weights too).
  std::vector<uint32_t> sort_particles(const std::vector<double>& v_reference) {
    if (tie_order == 0) return sort_index_desc(v_reference, N_prev);
    std::vector<uint32_t> idx(N_prev);
    for (uint32_t i = 0; i < N_prev; i++) idx[i] = i;
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) {
      if (lw_p[a] != lw_p[b]) return lw_p[a] > lw_p[b];
      const uint32_t ia = (pp[a].r << 28) | pp[a].d, ib = (pp[b].r << 28) | pp[b].d;
      return ia < ib;
    });
    return idx;
  }
  int tie_pairs;  // tap: adjacent equal finite values in the sorted order of this site
  void count_ties(const std::vector<double>& v, const std::vector<uint32_t>& sorted) {
    for (size_t i = 0; i + 1 < sorted.size(); i++) {
      const double a = v[sorted[i]], b = v[sorted[i + 1]];
      if (a == b && std::isfinite(a) && (a != 0.0 || tie_order)) tie_pairs++;  // W == 0 / logw == -inf entries are interchangeable
    }
  }

  // Smc.h:576-579
  void self_normalise() {
    std::vector<double> tmp(lw_c.begin(), lw_c.begin() + N_curr);
    lsum_c = normalise_exp(tmp, N_curr);
    W_c.assign(N_curr, 0.0);
    for (uint32_t n = 0; n < N_curr; n++) W_c[n] = std::exp(tmp[n]);
  }
  // Smc.h:114-188 (change-point proposal)
  void initialise() {
    step = 0;
    N_curr = R;
    pc.assign(N_curr, Particle());
    lw_c.assign(N_curr, 0.0);
    for (uint32_t n = 0; n < N_curr; n++) {
      pc[n].d = 1; pc[n].r = n;                                         // :463-471
      lw_c[n] = -std::log(static_cast<double>(R)) + logobs[0 * R + n];  // :485-491, :582-586, singleGroup.h:559-566
    }
    self_normalise();
    k_last = -1; drew = false; tie_pairs = 0; tie_flags = 0;
  }
  // resample.h:289-409
  void optimal_finite_state(uint32_t M) {
    uint32_t N = N_prev;
    for (uint32_t n = 0; n < lw_c.size(); n++) lw_c[n] = 0.0;  // :301
    std::vector<uint32_t> sorted = sort_particles(W_p);
    count_ties(tie_order ? lw_p : W_p, sorted);
    std::vector<double> q(N), logq(N), Q(N);
    for (uint32_t i = 0; i < N; i++) { q[i] = W_p[sorted[i]]; logq[i] = std::log(q[i]); }
    { double acc = 0.0; for (uint32_t i = N; i-- > 0;) { acc += q[i]; Q[i] = acc; } }  // reverse(cumsum(reverse(q)))
    uint32_t k_old = 1, k_new = 0;
    double logC = 0.0;
    while (k_new != k_old) {
      k_old = k_new;
      logC = std::log(static_cast<double>(M - k_old)) - std::log(Q[k_old]);
      uint32_t cnt = 0;
      for (uint32_t i = k_old; i < N; i++) if (logq[i] > -logC * 1.0) cnt++;
      k_new = k_old + cnt;
    }
    if (std::isfinite(logC)) {
      uint32_t K = k_new, L = M - K;
      for (uint32_t i = 0; i < K; i++) { anc[i] = sorted[i]; lw_c[i] = lw_p[sorted[i]]; }
      if (K < M) {
        std::vector<double> res(logq.begin() + K, logq.end());
        normalise_exp(res, res.size());
        for (size_t i = 0; i < res.size(); i++) res[i] = std::exp(res[i]);
        std::vector<uint32_t> ind(L);
        drew = true;
        systematic_base(unif[step], ind, res, L);  // resample.h:119-127: one arma::randu()
        for (uint32_t j = 0; j < L; j++) anc[K + j] = sorted[ind[j] + K];
        // tap: did an exact tie decide who survives?  (two equal residual weights, one drawn, the other not)
        std::vector<char> drawn(N - K, 0);
        for (uint32_t j = 0; j < L; j++) drawn[ind[j]] = 1;
        for (uint32_t i = K; i + 1 < N; i++)
          if ((tie_order ? lw_p[sorted[i]] == lw_p[sorted[i + 1]] : q[i] == q[i + 1]) && q[i] != 0.0 && drawn[i - K] != drawn[i + 1 - K]) tie_flags |= 2;   // (the canonical order sorts log-weights)
      }
      for (uint32_t n = K; n < M; n++) lw_c[n] = lsum_p - logC;
      k_last = static_cast<int>(K);
    } else {
      std::vector<uint32_t> idx = sort_particles(lw_p);
      tie_pairs = 0;
      count_ties(lw_p, idx);
      if (M < N && lw_p[idx[M - 1]] == lw_p[idx[M]] && std::isfinite(lw_p[idx[M]])) tie_flags |= 2;
      for (uint32_t i = 0; i < M; i++) { anc[i] = idx[i]; lw_c[i] = lw_p[idx[i]]; }
      k_last = -2;
    }
  }
  // Smc.h:406-450
  void resample_cp() {
    uint32_t M = N_curr - R;
    anc.assign(M, 0);
    if (N_curr < N_prev + R) {
      uint32_t F = 0;
      for (uint32_t n = 0; n < N_prev; n++) if (std::isfinite(lw_p[n])) F++;
      if (F > M) {
        optimal_finite_state(M);
      } else {
        std::vector<uint32_t> idx = sort_particles(lw_p);
        count_ties(lw_p, idx);
        if (M < N_prev && lw_p[idx[M - 1]] == lw_p[idx[M]] && std::isfinite(lw_p[idx[M]])) tie_flags |= 2;
        for (uint32_t i = 0; i < M; i++) { anc[i] = idx[i]; lw_c[i] = lw_p[idx[i]]; }
        k_last = -2;
      }
    } else {
      for (uint32_t n = 0; n < N_prev; n++) { anc[n] = n; lw_c[n] = lw_p[n]; }
      k_last = -1;
    }
  }
  // Smc.h:190-286
  void iterate() {
    step++;
    N_prev = N_curr;
    N_curr = (N_prev + R > Nmax) ? Nmax : N_prev + R;
    pp = pc; lw_p = lw_c; W_p = W_c; lsum_p = lsum_c;
    pc.resize(N_curr); lw_c.resize(N_curr, 0.0);
    uint32_t M = N_curr - R;
    drew = false;
    tie_pairs = 0; tie_flags = 0;
    resample_cp();
    if (tie_pairs > 0) tie_flags |= 1;
    // :504-522
    for (uint32_t n = 0; n < M; n++) { pc[n].d = pp[anc[n]].d + 1; pc[n].r = pp[anc[n]].r; }
    for (uint32_t r = 0; r < R; r++) { pc[M + r].d = 1; pc[M + r].r = r; }
    // :536-574
    const double* lo = logobs + step * R;
    for (uint32_t n = 0; n < M; n++)
      lw_c[n] += mp.log_trans(pc[n].d, pc[n].r, pp[anc[n]].d, pp[anc[n]].r) + lo[pc[n].r];
    std::vector<double> aux(N_prev);
    for (uint32_t r = 0; r < R; r++) {
      for (uint32_t n = 0; n < N_prev; n++)
        aux[n] = (mp.log_trans(1, r, pp[n].d, pp[n].r) + lo[r]) + lw_p[n];
      lw_c[M + r] = sum_exp(aux, N_prev);
    }
    self_normalise();
  }
  // Smc.h:288-326
  void backward_kernels() {
    uint32_t M = N_curr - R;
    bk.assign(R, std::vector<double>(N_prev));
    for (uint32_t r = 0; r < R; r++) {
      for (uint32_t n = 0; n < N_prev; n++)
        bk[r][n] = lw_p[n] + mp.log_trans(pc[M + r].d, pc[M + r].r, pp[n].d, pp[n].r);
      normalise_exp(bk[r], N_prev);
      bool nan = false;
      for (uint32_t n = 0; n < N_prev; n++) { bk[r][n] = std::exp(bk[r][n]); if (std::isnan(bk[r][n])) nan = true; }
      if (nan) for (uint32_t n = 0; n < N_prev; n++) bk[r][n] = 0.0;
    }
  }
  // Smc.h:329-338
  double filtered_mean(const std::vector<double>& x) const {
    double est = 0.0;
    for (uint32_t n = 0; n < N_curr; n++) est = est + W_c[n] * x[n];
    return est;
  }
  // Smc.h:352-362
  double filtered_variance(const std::vector<double>& x) const {
    double est = 0.0, m = filtered_mean(x);
    for (uint32_t n = 0; n < N_curr; n++) est += W_c[n] * std::pow(x[n] - m, 2.0);
    return est;
  }
};

}  // namespace

extern "C" {

double hygo_log_beta_binomial(uint32_t x, uint32_t n, double a, double b) { return log_beta_binomial(x, n, a, b); }

// singleGroup.h:610-627: sum over samples, s = 0..S-1 in order
int hygo_sg_emission(const double* alpha, const double* beta, uint32_t R, uint64_t T, uint32_t S,
                     const uint32_t* n_total, const uint32_t* n_meth, double* logobs) {
  for (uint64_t t = 0; t < T; t++)
    for (uint32_t r = 0; r < R; r++) {
      double ld = 0.0;
      for (uint32_t s = 0; s < S; s++) ld += log_beta_binomial(n_meth[t * S + s], n_total[t * S + s], alpha[r], beta[r]);
      logobs[t * R + r] = ld;
    }
  return 0;
}

int hygo_sg_tables(const double* vartheta, uint32_t n_vartheta, const double* theta, uint32_t dim_theta, uint32_t d_max,
                   double* rho, uint8_t* exit_status, double* grad_omega_log_rho, double* P, double* omega) {
  Params mp;
  int rc = mp.set_known(vartheta, n_vartheta);
  if (rc) return rc;
  if (dim_theta != mp.D) return -3;
  mp.set_unknown(theta);
  for (uint32_t r = 0; r < mp.R; r++) {
    mp.need(d_max, r);
    for (uint32_t d = 1; d <= d_max; d++) {
      if (rho) rho[r * d_max + d - 1] = mp.rho[r][d - 1];
      if (exit_status) exit_status[r * d_max + d - 1] = mp.exit_status[r][d - 1];
      if (grad_omega_log_rho) grad_omega_log_rho[r * d_max + d - 1] = mp.glr[r][d - 1];
    }
  }
  if (P) std::memcpy(P, mp.P.data(), sizeof(double) * mp.R * mp.R);
  if (omega) std::memcpy(omega, mp.omega.data(), sizeof(double) * mp.R);
  return 0;
}

int hygo_sg_run(const hygo_sg_args* a) {
  Filter f;
  int rc = f.mp.set_known(a->vartheta, a->n_vartheta);
  if (rc) return rc;
  if (a->dim_theta != f.mp.D) return -3;
  const uint32_t R = f.mp.R, D = f.mp.D;
  const uint64_t T = a->T;
  if (!a->uniforms_by_site) return -4;
  clock_t t1 = clock();
  f.mp.set_unknown(a->theta);
  std::vector<double> lo_own;
  if (a->logobs) {
    f.logobs = a->logobs;
  } else {
    lo_own.resize(T * R);
    hygo_sg_emission(f.mp.alpha.data(), f.mp.beta.data(), R, T, a->S, a->n_total, a->n_meth, lo_own.data());
    f.logobs = lo_own.data();
  }
  f.R = R; f.Nmax = a->n_particles_max; f.unif = a->uniforms_by_site; f.tie_order = a->tie_order;
  const uint32_t Mmax = f.Nmax - R;

  // --- OnlineMarginalSmoothing state (OnlineMarginalSmoothing.h) ---
  std::vector<std::vector<std::vector<double> > > psi_c, psi_p;  // [pending][R][Nmax]
  std::vector<uint32_t> psi_t;
  bool is_final = false;
  auto init_psi = [&]() {  // :119-146
    std::vector<std::vector<double> > aux(R, std::vector<double>(f.Nmax, 0.0));
    for (uint32_t r = 0; r < R; r++)
      for (uint32_t n = 0; n < f.N_curr; n++) aux[r][n] = (r == f.pc[n].r) ? 1.0 : 0.0;  // singleGroup.h:821-837
    psi_t.push_back(static_cast<uint32_t>(f.step));
    psi_c.push_back(aux);
  };
  auto store = [&]() {  // :197-255
    std::vector<std::vector<std::vector<double> > > keep;
    std::vector<uint32_t> keep_t;
    for (size_t s = 0; s < psi_c.size(); s++) {
      bool emit = true;
      if (!is_final) {
        uint32_t r = 0;
        while (r < R && f.filtered_variance(psi_c[s][r]) < a->epsilon) r++;
        if (r < R) emit = false;
      }
      if (emit) {
        uint32_t t = psi_t[s];
        if (a->regime_probs) {
          a->regime_probs[t * (1 + R)] = a->positions ? a->positions[t] : t;  // OnlineCombinedInference.h:114
          for (uint32_t r = 0; r < R; r++) a->regime_probs[t * (1 + R) + 1 + r] = f.filtered_mean(psi_c[s][r]);
        }
        if (a->finalised_at) a->finalised_at[t] = static_cast<int32_t>(f.step);
      } else {
        keep.push_back(psi_c[s]);
        keep_t.push_back(psi_t[s]);
      }
    }
    psi_c.swap(keep);
    psi_t.swap(keep_t);
  };
  auto update_psi = [&]() {  // :148-177
    uint32_t M = f.N_curr - R;
    for (size_t s = 0; s < psi_c.size(); s++)
      for (uint32_t r = 0; r < R; r++) {
        for (uint32_t n = 0; n < M; n++) psi_c[s][r][n] = psi_p[s][r][f.anc[n]];
        for (uint32_t q = 0; q < R; q++) {
          psi_c[s][r][M + q] = 0;
          for (uint32_t n = 0; n < f.N_prev; n++) psi_c[s][r][M + q] = psi_c[s][r][M + q] + f.bk[q][n] * psi_p[s][r][n];
        }
      }
  };

  // --- OnlineParameterEstimation state (OnlineParameterEstimation.h, GradientAscent.h) ---
  std::vector<double> theta(a->theta, a->theta + D), phi_c, phi_p, grad_c(D, 0.0), grad_p(D, 0.0), adam_m(D, 0.0), adam_v(D, 0.0), g(D), acc(D);
  uint32_t ga_iter = 0;
  const double b1 = 0.9, b2 = 0.999, adam_eps = std::exp(-8.0 * std::log(10));  // GradientAscent.h:61-63
  auto filtered_mean_vec = [&](const std::vector<double>& phi, std::vector<double>& out) {  // Smc.h:340-349
    for (uint32_t k = 0; k < D; k++) out[k] = 0.0;
    for (uint32_t n = 0; n < f.N_curr; n++)
      for (uint32_t k = 0; k < D; k++) out[k] = out[k] + f.W_c[n] * phi[n * D + k];
  };
  auto trace = [&](uint64_t t) {
    if (a->theta_trace) std::memcpy(a->theta_trace + t * D, theta.data(), sizeof(double) * D);
  };
  auto tap = [&](uint64_t t) {
    if (a->logz) a->logz[t] = f.lsum_c;
    if (a->n_curr) a->n_curr[t] = static_cast<int32_t>(f.N_curr);
    if (a->k_kept) a->k_kept[t] = f.k_last;
    if (a->drew_uniform) a->drew_uniform[t] = f.drew ? 1 : 0;
    if (a->n_pending) a->n_pending[t] = static_cast<int32_t>(psi_t.size());
    if (a->tie_pairs) a->tie_pairs[t] = f.tie_pairs;
    if (a->tie_flags) a->tie_flags[t] = static_cast<uint8_t>(f.tie_flags);
    if (a->support_hash) {
      uint64_t h = 0;
      for (uint32_t n = 0; n < f.N_curr; n++)
        if (f.lw_c[n] > NEG_INF) h += mix64((static_cast<uint64_t>(f.pc[n].r) << 28) | f.pc[n].d);
      a->support_hash[t] = h;
    }
    if (a->weights_prev && t > 0) {
      double* w = a->weights_prev + t * f.Nmax;
      for (uint32_t n = 0; n < f.Nmax; n++) w[n] = (n < f.N_prev) ? f.W_p[n] : std::numeric_limits<double>::quiet_NaN();
    }
    if (a->d_prev && t > 0) {
      int32_t* w = a->d_prev + t * f.Nmax;
      for (uint32_t n = 0; n < f.Nmax; n++) w[n] = (n < f.N_prev) ? static_cast<int32_t>(f.pp[n].d | (f.pp[n].r << 24)) : 0;
    }
    if (a->ancestors) {
      for (uint32_t m = 0; m < Mmax; m++) a->ancestors[t * Mmax + m] = (t > 0 && m < f.anc.size()) ? static_cast<int16_t>(f.anc[m]) : -1;
    }
  };

  // --- OnlineCombinedInference.h:48-118 ---
  f.initialise();
  if (a->use_smoothing) { init_psi(); store(); }  // OnlineMarginalSmoothing.h:40-50
  if (a->use_param_est) {                         // OnlineParameterEstimation.h:42-49,115-132
    phi_c.assign(static_cast<size_t>(f.Nmax) * D, 0.0);
    phi_p = phi_c;
    filtered_mean_vec(phi_c, grad_c);
    trace(0);
  }
  tap(0);
  for (uint64_t t = 1; t < T; t++) {
    f.iterate();
    f.backward_kernels();
    if (a->use_smoothing) {
      if (t == T - 1) is_final = true;
      psi_p = psi_c;  // OnlineMarginalSmoothing.h:58
      update_psi();
      init_psi();
      store();
    }
    if (a->use_param_est) {
      // OnlineParameterEstimation.h:135-158
      uint32_t M = f.N_curr - R;
      phi_p = phi_c;
      for (uint32_t n = 0; n < M; n++) {
        uint32_t an = f.anc[n];
        f.mp.grad_log_trans(g.data(), f.pc[n].d, f.pc[n].r, f.pp[an].d, f.pp[an].r);
        for (uint32_t k = 0; k < D; k++) phi_c[n * D + k] = phi_p[an * D + k] + g[k];
      }
      for (uint32_t r = 0; r < R; r++) {
        for (uint32_t k = 0; k < D; k++) acc[k] = 0.0;
        for (uint32_t n = 0; n < f.N_prev; n++) {
          f.mp.grad_log_trans(g.data(), f.pc[M + r].d, f.pc[M + r].r, f.pp[n].d, f.pp[n].r);
          for (uint32_t k = 0; k < D; k++) acc[k] = acc[k] + f.bk[r][n] * (phi_p[n * D + k] + g[k]);
        }
        for (uint32_t k = 0; k < D; k++) phi_c[(M + r) * D + k] = acc[k];
      }
      if (t % a->n_steps_without_update == 0) {  // :54-59
        grad_p = grad_c;
        filtered_mean_vec(phi_c, grad_c);
        std::vector<double> gr(D);
        for (uint32_t k = 0; k < D; k++) gr[k] = grad_c[k] - grad_p[k];
        double lr = a->lr_factor / std::pow(static_cast<double>(ga_iter + 1.0), a->lr_exponent);  // GradientAscent.h:109-112
        if (a->use_adam) {                                                                        // :114-155
          double c2 = 1.0 - std::pow(b2, ga_iter + 1), c1 = 1.0 - std::pow(b1, ga_iter + 1);
          for (uint32_t k = 0; k < D; k++) {
            adam_m[k] = b1 * adam_m[k] + (1.0 - b1) * gr[k];
            adam_v[k] = b2 * adam_v[k] + ((1.0 - b2) * gr[k]) * gr[k];
            theta[k] = theta[k] + lr * adam_m[k] * std::pow(std::sqrt(adam_v[k] / c2) + adam_eps, -1.0) / c1;
          }
        } else if (a->normalise_gradients) {  // :94-97
          double nrm = 0.0;
          for (uint32_t k = 0; k < D; k++) nrm += std::fabs(gr[k]);
          for (uint32_t k = 0; k < D; k++) theta[k] = theta[k] + lr * ((nrm > 0) ? gr[k] / nrm : gr[k]);
        } else {
          for (uint32_t k = 0; k < D; k++) theta[k] = theta[k] + lr * gr[k];
        }
        ga_iter++;
        f.mp.set_unknown(theta.data());
      }
      trace(t);
    }
    tap(t);
  }
  clock_t t2 = clock();
  if (a->seconds) *a->seconds = (static_cast<double>(t2) - static_cast<double>(t1)) / CLOCKS_PER_SEC;
  return 0;
}

}  // extern "C"
