// tests/emu/emu_kernels.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Compiles the device code of hygeia_b200/csrc/*.cuh with g++ against tests/emu/cuda_emu.h (a CPU emulation of the
// CUDA execution model) so that `pytest -m "not gpu"` can check the KERNEL LOGIC against the oracle on the GPU-less
// build box.  The product never links this file; on a GPU box the same device code runs through nvcc (hyg_api.cu).
#include "cuda_emu.h"

#include <algorithm>
#include <vector>

#include "../../hygeia_b200/csrc/hyg_tables.h"
#include "../../hygeia_b200/csrc/sg_emission.cuh"
#include "../../hygeia_b200/csrc/sg_filter.cuh"
#include "../../hygeia_b200/csrc/hyg_tg.cuh"

extern "C" {

// K2 under emulation.  vartheta/theta as in the reference; logobs T x R; uniforms T or NULL (-> Philox(seed, chain_id)).
int hygemu_sg_filter(const double* vartheta, uint32_t n_vartheta, const double* theta, uint32_t dim_theta, uint32_t n_particles,
                     uint64_t T, const double* logobs, const double* unif, uint64_t seed, uint32_t chain_id,
                     int use_smoothing, double epsilon, int lcap,
                     int use_param_est, int normalise, int use_adam, uint32_t n_steps, double lr_exponent, double lr_factor, double* theta_trace,
                     double* probs, double* logz, int* k_kept, unsigned char* drew, int* n_pending, int* n_curr,
                     int* finalised_at, unsigned long long* support_hash, unsigned char* tie_flags, int* status,
                     uint64_t t_off, uint64_t own_lo, uint64_t own_hi, int last_segment, double* seg_inc, int force_full_sort) {
  hyg::SgHostModel hm;
  if (hm.set_known(vartheta, n_vartheta)) return -1;
  if (hm.set_theta(theta, dim_theta, T)) return -2;
  if (hm.R < 2 || hm.R > 6) return -3;
  hyg::SgModelDev mdl;
  mdl.R = hm.R; mdl.u = hm.u; mdl.n_particles = static_cast<int>(n_particles); mdl.dcap = hm.dcap;
  for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) { mdl.P[i][j] = hm.P[i][j]; mdl.logP[i][j] = hm.logP[i][j]; }
  mdl.tab = reinterpret_cast<const double2*>(hm.tab.data());
  mdl.tabg = hm.tabg.data();
  hyg::SgChainDev ch;
  std::memset(&ch, 0, sizeof(ch));
  ch.T = T; ch.logobs = logobs; ch.unif = unif; ch.seed = seed; ch.chain_id = chain_id;
  ch.probs = probs; ch.logz = logz; ch.k_kept = k_kept; ch.drew = drew; ch.n_pending = n_pending; ch.n_curr = n_curr;
  ch.finalised_at = finalised_at; ch.support_hash = support_hash; ch.tie_flags = tie_flags; ch.status = status;
  // segment view (whole chain: t_off = 0, own = [0, T), last_segment = 1); pointers are already local
  ch.t_off = t_off; ch.own_lo = own_lo; ch.own_hi = own_hi; ch.last_segment = last_segment; ch.seg_inc = seg_inc;
  hyg::SgRunDev run;
  run.use_smoothing = use_smoothing; run.epsilon = epsilon; run.lcap = lcap;
  const size_t stride = static_cast<size_t>(lcap) * hm.R * HYG_NPMAX + (5 * static_cast<size_t>(lcap) + 1) / 2 + 8;
  std::vector<double> ws(stride);
  unsigned int queue = 0;
  run.psi_ws = ws.data(); run.psi_stride = stride; run.queue = &queue; run.n_chains = 1; run.force_full_sort = force_full_sort;
  run.use_param_est = use_param_est; run.normalise_gradients = normalise; run.use_adam = use_adam; run.n_steps_without_update = n_steps;
  run.lr_exponent = lr_exponent; run.lr_factor = lr_factor;
  for (int r = 0; r < HYG_RMAX; r++) run.kappa[r] = r < hm.R ? hm.kappa[r] : 1.0;
  run.pe_dcap = static_cast<uint32_t>(T + 8 < 64 ? 64 : T + 8);
  run.pe_stride = 5ull * hm.R * run.pe_dcap;
  std::vector<double> pews(use_param_est ? run.pe_stride : 1);
  run.pe_ws = pews.data();
  ch.theta0 = theta; ch.theta_trace = theta_trace;
  const hyg::SgModelDev* pm = &mdl;
  const hyg::SgChainDev* pc = &ch;
  // one instantiation per number of regimes, as the library dispatches them (parameter mode: R = 6 only under emulation)
  if (use_param_est) {   // (the static stand-in for the dynamic shared memory is sized for R = 6, the largest)
    switch (hm.R) {
      case 2: emu::launch(dim3(1), dim3(HYG_NT), [=]() { hyg::sg_filter_entry<2, true>(pm, pc, run); }); break;
      case 3: emu::launch(dim3(1), dim3(HYG_NT), [=]() { hyg::sg_filter_entry<3, true>(pm, pc, run); }); break;
      case 4: emu::launch(dim3(1), dim3(HYG_NT), [=]() { hyg::sg_filter_entry<4, true>(pm, pc, run); }); break;
      case 5: emu::launch(dim3(1), dim3(HYG_NT), [=]() { hyg::sg_filter_entry<5, true>(pm, pc, run); }); break;
      case 6: emu::launch(dim3(1), dim3(HYG_NT), [=]() { hyg::sg_filter_entry<6, true>(pm, pc, run); }); break;
      default: return -2;
    }
    return 0;
  }
  switch (hm.R) {
    case 2: emu::launch(dim3(1), dim3(HYG_NT), [=]() { hyg::sg_filter_entry<2, false>(pm, pc, run); }); break;
    case 3: emu::launch(dim3(1), dim3(HYG_NT), [=]() { hyg::sg_filter_entry<3, false>(pm, pc, run); }); break;
    case 4: emu::launch(dim3(1), dim3(HYG_NT), [=]() { hyg::sg_filter_entry<4, false>(pm, pc, run); }); break;
    case 5: emu::launch(dim3(1), dim3(HYG_NT), [=]() { hyg::sg_filter_entry<5, false>(pm, pc, run); }); break;
    case 6: emu::launch(dim3(1), dim3(HYG_NT), [=]() { hyg::sg_filter_entry<6, false>(pm, pc, run); }); break;
    default: return -2;
  }
  return 0;
}

// K1 under emulation: counts uint16 [S][pitch] (site fastest), logobs T x R.
int hygemu_sg_emission(const double* alpha, const double* beta, int R, uint64_t T, uint32_t S, uint64_t pitch,
                       const uint16_t* n_total, const uint16_t* n_meth, int nmax_table, int nmax_smem, int grid, int block,
                       double* logobs) {
  if (R != 6) return -3;
  std::vector<double> tab;
  hyg::build_emission_table(alpha, beta, R, nmax_table, tab);
  // split the sites into two "data sets" so the flattened tile space with several sets is exercised
  const uint64_t T0 = (T > 2200) ? 2050 : T;  // first set ends inside a tile
  hyg::SgEmissionSet sets[2];
  uint32_t n_sets = (T0 < T) ? 2 : 1;
  sets[0].T = T0; sets[0].pitch = pitch; sets[0].n_total = n_total; sets[0].n_meth = n_meth; sets[0].logobs = logobs; sets[0].tile0 = 0; sets[0].S = S;
  unsigned long long tiles = ((T0 + 1) / 2 + HYG_EM_TILE - 1) / HYG_EM_TILE;
  if (n_sets == 2) {
    // second set starts at site T0 (T0 is even, so the 4-byte pair loads stay aligned)
    sets[1].T = T - T0; sets[1].pitch = pitch; sets[1].n_total = n_total + T0; sets[1].n_meth = n_meth + T0; sets[1].logobs = logobs + T0 * R;
    sets[1].tile0 = tiles; sets[1].S = S;
    tiles += ((T - T0 + 1) / 2 + HYG_EM_TILE - 1) / HYG_EM_TILE;
  }
  hyg::SgEmissionArgs a;
  a.sets = sets; a.n_sets = n_sets; a.n_tiles = tiles;
  a.table = tab.data(); a.nmax_table = nmax_table; a.nmax_smem = nmax_smem;
  for (int r = 0; r < R; r++) { a.alpha[r] = alpha[r]; a.beta[r] = beta[r]; }
  (void)block;
  emu::launch(dim3(grid), dim3(HYG_EM_NT), [=]() { hyg::sg_emission_entry<6>(a); });
  return 0;
}

// K4/K5 under emulation: two-group filter + backward simulation of one chain.
int hygemu_tg_run(int R, int u, int M, int B, const double* logP /*R x R*/, const double* logPm /*2 x 2*/, const double* rho_c,
                  const double* rho_k, uint32_t dmax, uint64_t T, const double* lo_c, const double* lo_k, uint64_t seed, uint32_t chain,
                  int* traj, double* log_norm, int* taps, int presel0, int presel1, int scratch_from) {
  hyg::TgModelDev mdl;
  std::memset(&mdl, 0, sizeof(mdl));
  mdl.R = R; mdl.u = u; mdl.M = M; mdl.B = B; mdl.dmax = dmax;
  for (int i = 0; i < R; i++) for (int j = 0; j < R; j++) mdl.logP[i][j] = logP[i * R + j];
  for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) mdl.logPm[i][j] = logPm[i * 2 + j];
  const size_t nrho = static_cast<size_t>(R) * (dmax + 1);
  std::vector<double2> lrc(nrho), lrk(nrho);
  for (size_t i = 0; i < nrho; i++) {
    lrc[i] = make_double2(std::log(rho_c[i]), std::log(1.0 - rho_c[i]));
    lrk[i] = make_double2(std::log(rho_k[i]), std::log(1.0 - rho_k[i]));
  }
  mdl.lrho_c = lrc.data(); mdl.lrho_k = lrk.data();
  mdl.nl_rm1 = -std::log(static_cast<double>(R) - 1.0); mdl.nl_rm2 = -std::log(static_cast<double>(R) - 2.0);
  mdl.presel[0] = std::max(presel0 ? presel0 : M + 110, M + 96); mdl.presel[1] = std::max(presel1 ? presel1 : 3 * M + 250, mdl.presel[0]);
  mdl.big_from = scratch_from ? std::min(scratch_from - 1, HYG_TG_SORTMAX) : HYG_TG_SORTMAX;
  hyg::TgChainDev ch;
  ch.T = T; ch.lo_c = lo_c; ch.lo_k = lo_k; ch.seed = seed; ch.chain = chain; ch.traj = traj; ch.log_norm = log_norm; ch.taps = taps;
  hyg::TgRunDev run;
  run.t_max = T;
  run.anc_pitch = std::max(M, R * R);
  run.scratch_off = (sizeof(hyg::TgStepRec) * T + sizeof(hyg::TgAncRec) * T * run.anc_pitch + 255) & ~static_cast<size_t>(255);
  run.ws_stride = run.scratch_off + static_cast<size_t>(HYG_TG_BIGMAX) * 18 + 256;
  std::vector<unsigned char> ws(run.ws_stride);
  unsigned int queue = 0;
  run.ws = ws.data(); run.queue = &queue; run.n_chains = 1;
  const hyg::TgModelDev* pm = &mdl;
  const hyg::TgChainDev* pc = &ch;
  emu::launch(dim3(1), dim3(HYG_TG_NT), [=]() { hyg::tg_entry(pm, pc, run); });
  return 0;
}

}  // extern "C"
