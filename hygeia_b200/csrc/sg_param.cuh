// hygeia_b200/csrc/sg_param.cuh -- K3: online score-based parameter estimation, fused into the K2 recursion.
//
// Reference (/root/reference/src/single_group/src/cpp):
//   score recursion  phi_n in R^D            algorithms/OnlineParameterEstimation.h:115-176
//   d/dtheta log transition density          singleGroup.h:640-706   (emission / initial gradients are zero, :629-638,708-717)
//   ADAM / plain ascent                      misc/GradientAscent.h:52-155
//   theta -> P, omega, sojourn tables        singleGroup.h:197-335
//
// Device formulation.  With c[r'][r] = P[r'][r] / sumE[r] the backward kernel of the new particle (1, r) is
// bk_r[n] = e_n c[r_n][r], so   phi'_{(1,r)} = sum_n bk_r[n] (phi_n + grad_n)   collapses to the D x R class sums
// G[k][r'] = sum_{n in class r'} e_n phi_n[k] (a 36 x 250 x 6 contraction per site, done by all 288 threads over
// 8 chunks of 32 particles) plus closed-form terms in E[r'] and Eg[r'] = sum_{n in r'} e_n dlogrho(d_n, r').  Particles
// stay in their slots (sg_filter.cuh), so a continuing particle's phi stays where it is and ONE entry (the omega slot of its
// regime) is touched; only the <= R rewritten slots get a new vector.  phi lives in shared memory, [D][257] (pitch 257: the
// chunked contraction reads are bank-conflict free), updated in place after the class sums have been taken.
// Every n_steps sites: g = sum_n W_n phi_n, ADAM on g - g_prev, theta -> P, omega, and the sojourn tables are REBUILT
// ON THE DEVICE (h in parallel over (r, d); the cumulative sums H and Gh sequentially, one warp per regime, because
// rho = h / (1 - H) is ill-conditioned and only the reference's own summation order reproduces its values) up to
// d = max particle sojourn + n_steps + 2, after which every particle reloads the entries it carries.
#ifndef HYG_SG_PARAM_CUH
#define HYG_SG_PARAM_CUH

#include "hyg_common.cuh"
#include "hyg_dev_structs.h"

#define HYG_PHI_PITCH 257

namespace hyg {

template <int R> struct SgPeSmem {
  static constexpr int D = R * R;
  double phi[D][HYG_PHI_PITCH];
  double part[8][D][R];     // per-chunk partial class sums
  double Gs[D][R];          // G[k][r']
  double theta[D], adam_m[D], adam_v[D], grad_cur[D], grad_prev[D];
  double eprev[HYG_NPMAX];  // e_n = W_n c_new(d_n, r_n) of the previous particles
  double Etot[8], Egtot[8]; // class totals E[r'], Eg[r']
  double omega[8], kappa[8];
  double scr[HYG_RMAX][256];          // table rebuild staging: [h | h*glh | H used | Gh prev] x 64 per regime
  unsigned char exf[HYG_RMAX][64];    // exit flags of the staged chunk
  unsigned int iter;
  uint32_t dneed;
};

// theta -> P (row softmax of R-1 logits, zero diagonal), log P, omega (singleGroup.h:201-216; misc.h:34-37,793-798)
template <int R>
__device__ __forceinline__ void pe_set_theta(SgModelDev& mdl, SgPeSmem<R>& pe) {
  const int tid = threadIdx.x;
  if (tid < R) {
    const int r = tid;
    const double* blk = pe.theta + r * (R - 1);
    double mx = blk[0];
    for (int k = 1; k < R - 1; k++) mx = blk[k] > mx ? blk[k] : mx;
    double sm = 0.0;
    for (int k = 0; k < R - 1; k++) sm += exp(blk[k] - mx);
    const double logz = mx + log(sm);
    int k = 0;
    for (int c = 0; c < R; c++) {
      if (c == r) { mdl.P[r][c] = 0.0; mdl.logP[r][c] = -HYG_INF; continue; }
      const double pv = exp(blk[k++] - logz);
      mdl.P[r][c] = pv;
      mdl.logP[r][c] = log(pv);
    }
    pe.omega[r] = 1.0 / (1.0 + exp(-pe.theta[R * (R - 1) + r]));
  }
  __syncthreads();
}

// Sojourn tables for d = 1..dneed by the reference's recurrence (singleGroup.h:289-332): h, glh in parallel over (r, d);
// H = cumsum(h) and Gh = cumsum(h * glh) sequentially in chunks of 64 (one warp per regime).
//   tab[r*dcap + d-1] = {c_new, lc},  tabg[r*dcap + d-1] = dlogrho/dtheta_omega;  wh / wg are scratch of the same shape.
template <int R>
__device__ void pe_rebuild_tables(SgModelDev& mdl, SgPeSmem<R>& pe, double2* tab, double* tabg, double* wh, double* wg, uint32_t dcap,
                                  uint32_t dneed) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t u = static_cast<uint32_t>(mdl.u);
  for (uint32_t idx = tid; idx < R * dneed; idx += HYG_NT) {
    const uint32_t r = idx / dneed, d = idx % dneed;   // 0-based index, sojourn time d + 1
    double h = 0.0, glh = 0.0;
    if (d + 1 >= u) {
      const double k = static_cast<double>(d + 1 - u), om = pe.omega[r], ka = pe.kappa[r];
      // misc.h:673-693
      double lnb;
      if (d + 1 == u && om == 0.0) lnb = 0.0;
      else if (om == 0.0) lnb = -HYG_INF;
      else lnb = lgamma(k + ka) - lgamma(ka) - lgamma(k + 1.0) + ka * log(1.0 - om) + k * log(om);
      h = exp(lnb);
      glh = (k / om - ka / (1.0 - om)) * (2.0 + exp(-om) + exp(om));   // singleGroup.h:322, misc.h:92-95
    }
    wh[r * dcap + d] = h;
    wg[r * dcap + d] = glh;
  }
  __syncthreads();
  if (warp < R) {
    // The cumulative sums H[d] = H[d-1] + h[d] and Gh[d] are accumulated SEQUENTIALLY, in the reference's order:
    // rho = h / (1 - H[d-1]) is ill-conditioned once H approaches 1, so a tree-shaped scan would give
    // different values exactly where the reference's own values are rounding noise.  Chunks of 64 entries are staged in
    // shared memory, lane 0 walks them, all lanes then finish the entries in parallel.
    const int r = warp;
    double* st = pe.scr[r];
    unsigned char* exf = pe.exf[r];
    double Hprev = 0.0, Gprev = 0.0;
    bool exit_prev = false;
    for (uint32_t base = 0; base < dneed; base += 64) {
#pragma unroll
      for (int jj = 0; jj < 2; jj++) {
        const uint32_t j = lane + 32 * jj, d = base + j;
        const bool in = d < dneed;
        const double h = in ? wh[r * dcap + d] : 0.0;
        const double glh = in ? wg[r * dcap + d] : 0.0;
        st[j] = h;
        st[64 + j] = h * glh;
      }
      __syncwarp();
      if (lane == 0) {
        for (uint32_t j = 0; j < 64 && base + j < dneed; j++) {
          const uint32_t d = base + j;
          if (d + 1 < u) { st[128 + j] = 0.0; st[192 + j] = 0.0; exf[j] = 0; continue; }
          const double h = st[j];
          double Hused = Hprev;
          bool ex;
          if (exit_prev || Hprev >= 1.0) { Hused = 0.99999; ex = true; }       // singleGroup.h:309-314
          else ex = false;
          st[128 + j] = Hused;
          st[192 + j] = Gprev;
          exf[j] = ex ? 1 : 0;
          if (!ex) Hprev = Hprev + h;                                            // :317 (not updated once the exit flag is set)
          Gprev = Gprev + st[64 + j];                                            // :323
          exit_prev = ex;
        }
      }
      __syncwarp();
#pragma unroll
      for (int jj = 0; jj < 2; jj++) {
        const uint32_t j = lane + 32 * jj, d = base + j;
        if (d < dneed) {
          double c_new, lc, glr;
          if (d + 1 < u) {
            c_new = 0.0; lc = 0.0; glr = 0.0;
          } else {
            const double h = st[j], Hu = st[128 + j], Gp = st[192 + j];
            const bool ex = exf[j] != 0;
            const double glh = wg[r * dcap + d];
            const double rho = ex ? 1.0 : h / (1.0 - Hu);                        // :312,:318
            c_new = rho;
            lc = (!ex && rho <= 1.0) ? log(1.0 - rho) : -HYG_INF;                // singleGroup.h:601-604
            glr = glh + Gp / (1.0 - Hu);                                         // :324
          }
          tab[r * dcap + d] = make_double2(c_new, lc);
          tabg[r * dcap + d] = glr;
        }
      }
      __syncwarp();
    }
  }
  if (tid == 0) { mdl.dcap = dneed; pe.dneed = dneed; }
  __syncthreads();
}

// One ADAM / ascent step on g - g_prev (GradientAscent.h:72-155); thread k < D.
template <int R>
__device__ __forceinline__ void pe_ascent(SgPeSmem<R>& pe, const SgRunDev& run) {
  constexpr int D = R * R;
  const int k = threadIdx.x;
  const unsigned it = pe.iter;
  const double lr = run.lr_factor / pow(static_cast<double>(it) + 1.0, run.lr_exponent);
  double nrm = 1.0;
  if (!run.use_adam && run.normalise_gradients) {
    nrm = 0.0;
    for (int j = 0; j < D; j++) nrm += fabs(pe.grad_cur[j] - pe.grad_prev[j]);
    if (!(nrm > 0.0)) nrm = 1.0;
  }
  if (k < D) {
    const double gr = pe.grad_cur[k] - pe.grad_prev[k];
    if (run.use_adam) {
      const double b1 = 0.9, b2 = 0.999, eps = exp(-8.0 * log(10.0));
      const double c1 = 1.0 - pow(b1, static_cast<double>(it + 1)), c2 = 1.0 - pow(b2, static_cast<double>(it + 1));
      const double m = b1 * pe.adam_m[k] + (1.0 - b1) * gr;
      const double v = b2 * pe.adam_v[k] + ((1.0 - b2) * gr) * gr;
      pe.adam_m[k] = m; pe.adam_v[k] = v;
      pe.theta[k] = pe.theta[k] + lr * m * pow(sqrt(v / c2) + eps, -1.0) / c1;
    } else {
      pe.theta[k] = pe.theta[k] + lr * (gr / nrm);
    }
  }
  __syncthreads();
  if (k == 0) pe.iter = it + 1;
}

}  // namespace hyg
#endif
