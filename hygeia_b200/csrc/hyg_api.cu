// hygeia_b200/csrc/hyg_api.cu -- host side of the C ABI declared in include/hygeia_b200.h.
//
// Owns device memory, streams and launches; the model constants are built on the host (hyg_tables.cpp) in the
// reference's order of operations and uploaded once per theta.  No CPU fallback: without a usable device
// hyg_create() fails and nothing computes.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <numeric>
#include <queue>
#include <functional>
#include <string>
#include <vector>

#include "../../include/hygeia_b200.h"
#include "hyg_tables.h"
#include "sg_emission.cuh"
#include "sg_filter.cuh"
#include "hyg_tg.cuh"
#include "hyg_dmp.cuh"


#define HYG_VERSION_STR "hygeia_b200 0.1.0 (sm_100a)"

namespace {

std::string g_create_error;

struct Dataset {
  uint64_t T = 0;
  uint32_t S = 0;
  uint64_t pitch = 0;
  const uint16_t* d_nt = nullptr;
  const uint16_t* d_nm = nullptr;
  bool owned = false;
  double* d_logobs = nullptr;
  bool emitted = false;         // d_logobs holds the emission table of these counts (hyg_sg_emission ran after they were added)
};

struct ChainBuf {
  hyg_sg_chain host;  // copy of the caller's descriptor
  uint64_t T = 0;
  double* d_unif = nullptr;
  uint32_t* d_pos = nullptr;
  double* d_probs = nullptr;
  bool probs_mapped = false;   // d_probs aliases the caller's pinned host buffer: rows stream out of K2, nothing to download
  double* d_logz = nullptr;
  int* d_k = nullptr;
  unsigned char* d_drew = nullptr;
  int* d_npend = nullptr;
  int* d_ncurr = nullptr;
  int* d_fin = nullptr;
  unsigned long long* d_hash = nullptr;
  unsigned char* d_tie = nullptr;
  int* d_status = nullptr;
  double* d_ovl = nullptr;               // segmented execution: [n_seg][HYG_OVL_ROWS][R] rows recomputed in the right halos
  unsigned long long* d_ovlmax = nullptr;
  uint32_t ovl_cap = 0;
  double* d_trace = nullptr;
  double* d_seginc = nullptr;   // segmented execution: log Z increment of every segment
  uint32_t n_seg = 1, seginc_cap = 0;
};

}  // namespace

struct hyg_ctx {
  int device = 0;
  int num_sms = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  hyg::SgHostModel hm;
  bool model_set = false, theta_set = false;
  bool model_kappa_fixed = true, model_had_kappa = false;
  uint64_t theta_t_max = 0;
  double2* d_tab = nullptr;
  double* d_tabg = nullptr;
  hyg::SgModelDev* d_mdl = nullptr;
  double* d_emtab = nullptr;
  hyg::SgEmissionSet* d_sets = nullptr;
  size_t d_sets_cap = 0;
  int nmax_table = 255;
  std::vector<Dataset> ds;
  std::vector<ChainBuf> chains;
  std::vector<uint32_t> order;  // launch order (longest first) -> caller index
  hyg::SgChainDev* d_chains = nullptr;
  double* d_psi = nullptr;
  size_t psi_bytes = 0;
  double* d_pe = nullptr;
  size_t pe_bytes = 0;
  double* d_theta0 = nullptr;
  unsigned int* d_queue = nullptr;
  size_t d_chains_cap = 0;
  uint64_t seg_sites = 0, seg_halo_left = 5000, seg_halo_right = 5000;   // hyg_sg_set_segmentation (0 = whole chains)
  hyg::SgLogzFix* d_fix = nullptr;
  size_t d_fix_cap = 0;
  hyg::SgOvlCheck* d_ovlchk = nullptr;
  size_t d_ovlchk_cap = 0;
  bool last_allow_forced = false;
  bool zero_copy_out = true;
  uint32_t n_units_last = 0;
  uint64_t seg_sites_last = 0;
  uint32_t grid_last = 0;
  uint32_t n_particles_staged = 0;
  std::multimap<size_t, void*> pool_free_blocks;
  std::map<void*, size_t> pool_live;
  cudaEvent_t ev_em0 = nullptr, ev_em1 = nullptr, ev_f0 = nullptr, ev_f1 = nullptr, ev_d0 = nullptr, ev_d1 = nullptr;
  bool timed_em = false, timed_f = false;
  uint32_t em_launches = 0, f_launches = 0;
  // two-group
  bool tg_set = false;
  hyg::TgModelDev tg_host;
  hyg::TgModelDev* d_tg_mdl = nullptr;
  double* d_tg_rho = nullptr;   // [2][R][dmax + 1]
};

namespace {

int fail(hyg_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}
#define HYG_CUDA(ctx, call)                                                                                        \
  do {                                                                                                             \
    cudaError_t e_ = (call);                                                                                       \
    if (e_ != cudaSuccess) return fail((ctx), HYG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));   \
  } while (0)

template <class T> void dfree(T*& p) {
  if (p) cudaFree(const_cast<typename std::remove_const<T>::type*>(p));
  p = nullptr;
}

// Per-context caching allocator for the per-sweep buffers (counts, emission tables, posteriors): a whole-genome sweep
// stages ~10 GB in ~150 buffers, and cudaMalloc/cudaFree of those dominated the end-to-end time of repeated sweeps.
// Freed blocks are kept and reused for requests of the same size; hyg_destroy releases them.
cudaError_t pool_alloc(hyg_ctx* c, void** p, size_t bytes);
template <class T> void pool_free(hyg_ctx* c, T*& p);

cudaError_t pool_alloc(hyg_ctx* c, void** p, size_t bytes) {
  if (bytes == 0) bytes = 8;
  auto it = c->pool_free_blocks.find(bytes);
  if (it != c->pool_free_blocks.end()) {
    *p = it->second;
    c->pool_free_blocks.erase(it);
    c->pool_live[*p] = bytes;
    return cudaSuccess;
  }
  cudaError_t e = cudaMalloc(p, bytes);
  if (e != cudaSuccess) {
    // out of memory: drop the cache and retry once
    for (auto& kv : c->pool_free_blocks) cudaFree(kv.second);
    c->pool_free_blocks.clear();
    cudaGetLastError();
    e = cudaMalloc(p, bytes);
  }
  if (e == cudaSuccess) c->pool_live[*p] = bytes;
  return e;
}
template <class T> void pool_free(hyg_ctx* c, T*& p) {
  if (!p) return;
  void* q = const_cast<void*>(static_cast<const void*>(p));
  auto it = c->pool_live.find(q);
  if (it != c->pool_live.end()) {
    c->pool_free_blocks.insert({it->second, q});
    c->pool_live.erase(it);
  } else {
    cudaFree(q);
  }
  p = nullptr;
}
void pool_release(hyg_ctx* c) {
  for (auto& kv : c->pool_free_blocks) cudaFree(kv.second);
  c->pool_free_blocks.clear();
}

void free_chains(hyg_ctx* c) {
  for (auto& b : c->chains) {
    if (b.probs_mapped) b.d_probs = nullptr;
    pool_free(c, b.d_unif); pool_free(c, b.d_pos); pool_free(c, b.d_probs); pool_free(c, b.d_logz); pool_free(c, b.d_k);
    pool_free(c, b.d_drew); pool_free(c, b.d_npend); pool_free(c, b.d_ncurr); pool_free(c, b.d_fin); pool_free(c, b.d_hash); pool_free(c, b.d_tie);
    pool_free(c, b.d_status); pool_free(c, b.d_trace); pool_free(c, b.d_seginc); pool_free(c, b.d_ovl); pool_free(c, b.d_ovlmax);
  }
  c->chains.clear();
  c->order.clear();
}

void free_datasets(hyg_ctx* c) {
  for (auto& d : c->ds) {
    if (d.owned) { pool_free(c, d.d_nt); pool_free(c, d.d_nm); }
    pool_free(c, d.d_logobs);
  }
  c->ds.clear();
}

// Segmented execution: shift the rows of segment j >= 1 by the log Z accumulated over the segments before it.
__global__ void sg_logz_fix_kernel(const hyg::SgLogzFix* units, unsigned int n_units) {
  __shared__ double s_off;
  for (unsigned int k = blockIdx.x; k < n_units; k += gridDim.x) {
    const hyg::SgLogzFix f = units[k];
    if (threadIdx.x == 0) {
      double off = 0.0;
      for (unsigned int i = 0; i < f.j; i++) off += f.seg_inc[i];
      s_off = off;
    }
    __syncthreads();
    const double off = s_off;
    for (unsigned long long t = threadIdx.x; t < f.len; t += blockDim.x) f.logz[t] += off;
    __syncthreads();
  }
}

// Segmented execution, left-halo check: the rows segment j recomputed in its right halo against the rows segment j+1 wrote.
__global__ void sg_overlap_check_kernel(const hyg::SgOvlCheck* units, unsigned int n_units) {
  for (unsigned int k = blockIdx.x; k < n_units; k += gridDim.x) {
    const hyg::SgOvlCheck u = units[k];
    if (threadIdx.x < u.rows) {
      double worst = 0.0;
      bool have = false;
      for (unsigned int q = 0; q < u.R; q++) {
        const double a = u.ovl[threadIdx.x * u.R + q];
        if (a != a) continue;   // row not finalised inside the halo
        have = true;
        const double b = u.probs[static_cast<size_t>(threadIdx.x) * (u.R + 1) + 1 + q];
        const double d = fabs(a - b);
        worst = (d > worst || d != d) ? d : worst;
      }
      if (have) {
        if (worst != worst) worst = 1.0;   // a NaN in the owner's row: report as a full-scale difference
        atomicMax(u.max_bits, static_cast<unsigned long long>(__double_as_longlong(worst)));
        if (worst > 1e-6) atomicAdd(u.status + 7, 1);
      }
    }
  }
}

// Segment size for HYG_SEGMENT_AUTO: the candidate (multiples of halo_left / 2) whose units, handed longest-first to `workers`
// persistent CTAs, finish earliest.  Cost of a unit = its sites + the left halo it has to step through first.
uint64_t choose_segment_sites(const std::vector<uint64_t>& chain_T, uint64_t halo_left, int workers) {
  uint64_t longest = 0;
  for (uint64_t T : chain_T) longest = std::max(longest, T);
  if (workers < 1) workers = 1;
  const uint64_t step = std::max<uint64_t>(halo_left, 256);
  uint64_t best_seg = 0, best_span = longest;   // whole chains: the longest chain bounds the launch
  if (chain_T.size() >= static_cast<size_t>(workers) * 4) return 0;   // enough independent chains already
  for (uint64_t seg = 2 * step; seg < longest + step; seg += step / 2) {
    std::vector<uint64_t> units;
    for (uint64_t T : chain_T) {
      const uint64_t ns = (T + seg - 1) / seg, len = (T + ns - 1) / ns;
      for (uint64_t j = 0; j < ns && j * len < T; j++) units.push_back(std::min(len, T - j * len) + (j ? halo_left : 0));
    }
    std::sort(units.begin(), units.end(), std::greater<uint64_t>());
    std::priority_queue<uint64_t, std::vector<uint64_t>, std::greater<uint64_t>> load;
    for (int w = 0; w < workers; w++) load.push(0);
    uint64_t span = 0;
    for (uint64_t u : units) { uint64_t l = load.top() + u; load.pop(); load.push(l); span = std::max(span, l); }
    if (span < best_span) { best_span = span; best_seg = seg; }
  }
  return best_seg;
}

template <int R> int launch_emission(hyg_ctx* c, const hyg::SgEmissionArgs& a, size_t smem, int grid) {
  HYG_CUDA(c, cudaFuncSetAttribute(hyg::sg_emission_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  hyg::sg_emission_kernel<R><<<grid, HYG_EM_NT, smem, c->stream>>>(a);
  HYG_CUDA(c, cudaGetLastError());
  return HYG_OK;
}

template <int R, bool PE> int launch_filter(hyg_ctx* c, const hyg::SgRunDev& run, int grid) {
  const size_t smem = PE ? sizeof(hyg::SgPeSmem<R>) : 0;
  if (PE) HYG_CUDA(c, cudaFuncSetAttribute(hyg::sg_filter_kernel<R, PE>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  if (!PE && run.n_chains <= c->num_sms) hyg::sg_filter_kernel_sparse<R><<<grid, HYG_NT, 0, c->stream>>>(c->d_mdl, c->d_chains, run);
  else hyg::sg_filter_kernel<R, PE><<<grid, HYG_NT, smem, c->stream>>>(c->d_mdl, c->d_chains, run);
  HYG_CUDA(c, cudaGetLastError());
  return HYG_OK;
}
template <int R, bool PE> int filter_occupancy(int* occ) {
  const size_t smem = PE ? sizeof(hyg::SgPeSmem<R>) : 0;
  if (PE) cudaFuncSetAttribute(hyg::sg_filter_kernel<R, PE>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, hyg::sg_filter_kernel<R, PE>, HYG_NT, smem) == cudaSuccess ? 0 : -1;
}
#define HYG_DISPATCH_R6(R_, EXPR)             \
  switch (R_) {                               \
    case 2: { constexpr int RR = 2; EXPR; } break; \
    case 3: { constexpr int RR = 3; EXPR; } break; \
    case 4: { constexpr int RR = 4; EXPR; } break; \
    case 5: { constexpr int RR = 5; EXPR; } break; \
    case 6: { constexpr int RR = 6; EXPR; } break; \
    default: break;                           \
  }

#define HYG_DISPATCH_R(R_, EXPR)              \
  switch (R_) {                               \
    case 2: { constexpr int RR = 2; EXPR; } break; \
    case 3: { constexpr int RR = 3; EXPR; } break; \
    case 4: { constexpr int RR = 4; EXPR; } break; \
    case 5: { constexpr int RR = 5; EXPR; } break; \
    case 6: { constexpr int RR = 6; EXPR; } break; \
    case 7: { constexpr int RR = 7; EXPR; } break; \
    default: break;                           \
  }

}  // namespace

extern "C" {

const char* hyg_version(void) { return HYG_VERSION_STR; }
const char* hyg_create_error(void) { return g_create_error.c_str(); }
const char* hyg_last_error(hyg_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
void* hyg_stream(hyg_ctx* ctx) { return ctx ? static_cast<void*>(ctx->stream) : nullptr; }

hyg_ctx* hyg_create(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                     " (this library has no CPU fallback)";
    return nullptr;
  }
  if (device < 0 || device >= n) { g_create_error = "device index out of range"; return nullptr; }
  if ((e = cudaSetDevice(device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return nullptr; }
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return nullptr; }
  if (prop.major < 10) {
    g_create_error = "device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) + "; this build targets sm_100a only";
    return nullptr;
  }
  hyg_ctx* c = new hyg_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    g_create_error = cudaGetErrorString(e);
    delete c;
    return nullptr;
  }
  cudaEventCreate(&c->ev_em0); cudaEventCreate(&c->ev_em1); cudaEventCreate(&c->ev_f0); cudaEventCreate(&c->ev_f1);
  cudaEventCreate(&c->ev_d0); cudaEventCreate(&c->ev_d1);
  cudaMalloc(&c->d_mdl, sizeof(hyg::SgModelDev));
  cudaMalloc(&c->d_queue, sizeof(unsigned int));
  return c;
}

void hyg_destroy(hyg_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  free_chains(c);
  free_datasets(c);
  pool_free(c, c->d_tab); pool_free(c, c->d_tabg); pool_free(c, c->d_emtab);
  pool_release(c);
  dfree(c->d_mdl); dfree(c->d_sets); dfree(c->d_psi); dfree(c->d_pe); dfree(c->d_theta0); dfree(c->d_queue); dfree(c->d_chains); dfree(c->d_fix); dfree(c->d_ovlchk); dfree(c->d_tg_mdl); dfree(c->d_tg_rho);
  cudaEventDestroy(c->ev_em0); cudaEventDestroy(c->ev_em1); cudaEventDestroy(c->ev_f0); cudaEventDestroy(c->ev_f1);
  cudaEventDestroy(c->ev_d0); cudaEventDestroy(c->ev_d1);
  cudaStreamDestroy(c->stream);
  delete c;
}

int hyg_sync(hyg_ctx* c) {
  if (!c) return HYG_ERR_ARG;
  HYG_CUDA(c, cudaSetDevice(c->device));
  HYG_CUDA(c, cudaStreamSynchronize(c->stream));
  return HYG_OK;
}

int hyg_sg_set_model(hyg_ctx* c, uint32_t R, uint32_t u, const double* alpha, const double* beta, int kappa_fixed, const double* kappa) {
  if (!c || !alpha || !beta) return fail(c, HYG_ERR_ARG, "null argument");
  // unchanged model (repeated sweeps re-stage it): keep the tables that are already on the device
  if (c->model_set && c->hm.R == static_cast<int>(R) && c->hm.u == static_cast<int>(u) && c->model_kappa_fixed == (kappa_fixed != 0) &&
      std::equal(alpha, alpha + R, c->hm.alpha.begin()) && std::equal(beta, beta + R, c->hm.beta.begin()) &&
      (!kappa || std::equal(kappa, kappa + R, c->hm.kappa.begin())) && (kappa || !c->model_had_kappa))
    return HYG_OK;
  int rc = c->hm.set_known(static_cast<int>(R), static_cast<int>(u), alpha, beta, kappa_fixed, kappa);
  if (rc) return fail(c, rc == -2 ? HYG_ERR_UNSUPPORTED : HYG_ERR_ARG, c->hm.err);
  c->model_set = true;
  c->theta_set = false;
  c->model_kappa_fixed = kappa_fixed != 0;
  c->model_had_kappa = kappa != nullptr;
  // emission table (misc.h:630-640 tabulated over the triangle n <= nmax_table)
  HYG_CUDA(c, cudaSetDevice(c->device));
  std::vector<double> tab;
  hyg::build_emission_table(c->hm.alpha.data(), c->hm.beta.data(), c->hm.R, c->nmax_table, tab);
  pool_free(c, c->d_emtab);   // pooled: cudaFree would synchronise the whole device, other contexts' kernels included
  HYG_CUDA(c, pool_alloc(c, reinterpret_cast<void**>(&c->d_emtab), tab.size() * sizeof(double)));
  HYG_CUDA(c, cudaMemcpyAsync(c->d_emtab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  HYG_CUDA(c, cudaStreamSynchronize(c->stream));
  return HYG_OK;
}

int hyg_sg_set_vartheta(hyg_ctx* c, const double* vt, uint32_t n) {
  if (!c || !vt || n < 2) return fail(c, HYG_ERR_ARG, "bad vartheta");
  const uint32_t R = static_cast<uint32_t>(vt[1]);
  if (R < 2 || R > 6 || n < 2 * R + 3) return fail(c, HYG_ERR_ARG, "bad vartheta (2 <= R <= 6)");
  const int kf = vt[2 * R + 2] != 0.0;
  if (kf && n < 3 * R + 3) return fail(c, HYG_ERR_ARG, "vartheta lacks kappa");
  return hyg_sg_set_model(c, R, static_cast<uint32_t>(vt[0]), vt + 2, vt + 2 + R, kf, kf ? vt + 2 * R + 3 : nullptr);
}

int hyg_sg_set_theta(hyg_ctx* c, const double* theta, uint32_t dim, uint64_t t_max) {
  if (!c || !theta) return fail(c, HYG_ERR_ARG, "null argument");
  if (!c->model_set) return fail(c, HYG_ERR_STATE, "hyg_sg_set_model first");
  // unchanged theta and a table that is already long enough: nothing to rebuild
  if (c->theta_set && dim == c->hm.theta.size() && std::equal(theta, theta + dim, c->hm.theta.begin()) && t_max <= c->theta_t_max) return HYG_OK;
  int rc = c->hm.set_theta(theta, dim, t_max);
  if (rc) return fail(c, HYG_ERR_ARG, c->hm.err);
  c->theta_t_max = t_max;
  HYG_CUDA(c, cudaSetDevice(c->device));
  pool_free(c, c->d_tab); pool_free(c, c->d_tabg);
  HYG_CUDA(c, pool_alloc(c, reinterpret_cast<void**>(&c->d_tab), c->hm.tab.size() * sizeof(double)));
  HYG_CUDA(c, pool_alloc(c, reinterpret_cast<void**>(&c->d_tabg), c->hm.tabg.size() * sizeof(double)));
  HYG_CUDA(c, cudaMemcpyAsync(c->d_tab, c->hm.tab.data(), c->hm.tab.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  HYG_CUDA(c, cudaMemcpyAsync(c->d_tabg, c->hm.tabg.data(), c->hm.tabg.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  HYG_CUDA(c, cudaStreamSynchronize(c->stream));
  c->theta_set = true;
  return HYG_OK;
}

int hyg_sg_get_tables(hyg_ctx* c, double* P, double* omega, uint32_t d_max, double* rho, uint8_t* exit_status) {
  if (!c) return HYG_ERR_ARG;
  if (!c->theta_set) return fail(c, HYG_ERR_STATE, "hyg_sg_set_theta first");
  const int R = c->hm.R;
  if (P) for (int i = 0; i < R; i++) for (int j = 0; j < R; j++) P[i * R + j] = c->hm.P[i][j];
  if (omega) for (int i = 0; i < R; i++) omega[i] = c->hm.omega[i];
  for (int r = 0; r < R; r++)
    for (uint32_t d = 0; d < d_max; d++) {
      const uint32_t i = d < c->hm.dcap ? d : c->hm.dcap - 1;
      const double cn = c->hm.tab[(static_cast<size_t>(r) * c->hm.dcap + i) * 2], lc = c->hm.tab[(static_cast<size_t>(r) * c->hm.dcap + i) * 2 + 1];
      const bool ex = std::isinf(lc) && lc < 0 && cn == 1.0;
      if (rho) rho[r * d_max + d] = ex ? 1.0 : (d + 1 < static_cast<uint32_t>(c->hm.u) ? 0.0 : cn);
      if (exit_status) exit_status[r * d_max + d] = ex ? 1 : 0;
    }
  return HYG_OK;
}

int hyg_sg_clear(hyg_ctx* c) {
  if (!c) return HYG_ERR_ARG;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  free_chains(c);
  free_datasets(c);
  return HYG_OK;
}

int hyg_sg_add_dataset(hyg_ctx* c, uint64_t T, uint32_t S, const uint16_t* n_total, const uint16_t* n_meth, int on_device, uint64_t pitch) {
  if (!c || !n_total || !n_meth || T == 0 || S == 0) return fail(c, HYG_ERR_ARG, "bad data set");
  if (!c->model_set) return fail(c, HYG_ERR_STATE, "hyg_sg_set_model first");
  HYG_CUDA(c, cudaSetDevice(c->device));
  Dataset d;
  d.T = T; d.S = S;
  if (on_device) {
    if (pitch < T || (pitch & 1)) return fail(c, HYG_ERR_ARG, "device pitch must be even and >= T");
    d.pitch = pitch; d.d_nt = n_total; d.d_nm = n_meth; d.owned = false;
  } else {
    d.pitch = (T + 7) / 8 * 8;
    uint16_t *a = nullptr, *b = nullptr;
    const size_t bytes = static_cast<size_t>(S) * d.pitch * sizeof(uint16_t);
    HYG_CUDA(c, pool_alloc(c, reinterpret_cast<void**>(&a), bytes));
    HYG_CUDA(c, pool_alloc(c, reinterpret_cast<void**>(&b), bytes));
    if (d.pitch != T) {  // zero only the pad columns (the kernels read whole site pairs)
      HYG_CUDA(c, cudaMemset2DAsync(a + T, d.pitch * 2, 0, (d.pitch - T) * 2, S, c->stream));
      HYG_CUDA(c, cudaMemset2DAsync(b + T, d.pitch * 2, 0, (d.pitch - T) * 2, S, c->stream));
    }
    // host rows may be longer than T (a window of a chromosome-wide matrix): `pitch` is then the host row pitch in elements
    if (pitch != 0 && pitch < T) return fail(c, HYG_ERR_ARG, "host pitch must be >= T (or 0 = T)");
    const size_t spitch = (pitch ? pitch : T) * sizeof(uint16_t);
    HYG_CUDA(c, cudaMemcpy2DAsync(a, d.pitch * 2, n_total, spitch, T * 2, S, cudaMemcpyHostToDevice, c->stream));
    HYG_CUDA(c, cudaMemcpy2DAsync(b, d.pitch * 2, n_meth, spitch, T * 2, S, cudaMemcpyHostToDevice, c->stream));
    d.d_nt = a; d.d_nm = b; d.owned = true;
  }
  HYG_CUDA(c, pool_alloc(c, reinterpret_cast<void**>(&d.d_logobs), static_cast<size_t>(T) * c->hm.R * sizeof(double)));
  c->ds.push_back(d);
  return static_cast<int>(c->ds.size()) - 1;
}

void hyg_sg_default_run_args(hyg_sg_run_args* a) {
  if (!a) return;
  std::memset(a, 0, sizeof(*a));
  a->n_particles_max = 250;             // bin/estimate_parameters_and_regimes:109-115
  a->smc_proposal_type = 1;             // :310
  a->smc_resample_type = 2;             // :311
  a->use_online_marginal_smoothing = 1;
  a->epsilon = 0.01;                    // :132-138
  a->use_online_parameter_estimation = 0;
  a->normalise_gradients = 0;
  a->use_adam = 1;
  a->n_steps_without_parameter_update = 200;
  a->learning_rate_exponent = 0.1;
  a->learning_rate_factor = 0.01;
  a->lag_capacity = 1024;
}

int hyg_sg_set_segmentation(hyg_ctx* c, uint64_t segment_sites, uint64_t halo_left, uint64_t halo_right) {
  if (!c) return HYG_ERR_ARG;
  if (segment_sites != HYG_SEGMENT_AUTO && segment_sites > 0 && segment_sites < 16) return fail(c, HYG_ERR_ARG, "segment_sites must be 0 (whole chains) or >= 16");
  c->seg_sites = segment_sites;
  c->seg_halo_left = halo_left;
  c->seg_halo_right = halo_right;
  return HYG_OK;
}

int hyg_sg_set_zero_copy_outputs(hyg_ctx* c, int enable) {
  if (!c) return HYG_ERR_ARG;
  c->zero_copy_out = enable != 0;
  return HYG_OK;
}

int hyg_sg_filter_units(hyg_ctx* c, uint32_t* n_units, uint64_t* segment_sites, uint32_t* resident_ctas) {
  if (!c) return HYG_ERR_ARG;
  if (n_units) *n_units = c->n_units_last;
  if (resident_ctas) *resident_ctas = c->grid_last;
  if (segment_sites) *segment_sites = c->seg_sites_last;
  return HYG_OK;
}

int hyg_sg_set_chains(hyg_ctx* c, const hyg_sg_chain* chains, uint32_t n) {
  if (!c || !chains || n == 0) return fail(c, HYG_ERR_ARG, "no chains");
  if (!c->model_set) return fail(c, HYG_ERR_STATE, "hyg_sg_set_model first");
  HYG_CUDA(c, cudaSetDevice(c->device));
  HYG_CUDA(c, cudaStreamSynchronize(c->stream));
  free_chains(c);
  const int R = c->hm.R;
  for (uint32_t i = 0; i < n; i++)
    if (chains[i].dataset >= c->ds.size()) return fail(c, HYG_ERR_ARG, "chain refers to an unknown data set");
  c->chains.resize(n);
  // a failing allocation below leaves no half-staged chains behind (hyg_sg_filter indexes c->order)
  struct Guard { hyg_ctx* c; bool ok = false; ~Guard() { if (!ok) free_chains(c); } } guard{c};
  for (uint32_t i = 0; i < n; i++) {
    ChainBuf& b = c->chains[i];
    b.host = chains[i];
    const uint64_t T = c->ds[chains[i].dataset].T;
    b.T = T;
    if (chains[i].uniforms) {
      HYG_CUDA(c, pool_alloc(c, reinterpret_cast<void**>(&b.d_unif), T * sizeof(double)));
      HYG_CUDA(c, cudaMemcpyAsync(b.d_unif, chains[i].uniforms, T * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    }
    if (chains[i].regime_probs) {
      // pinned (page-locked) caller memory is addressable from the device: K2 then writes every finalised row straight into
      // it (56 contiguous bytes per row, posted writes over PCIe at ~3 GB/s of the link's ~55) and there is no D2H stage
      cudaPointerAttributes pa;
      std::memset(&pa, 0, sizeof(pa));
      if (c->zero_copy_out && cudaPointerGetAttributes(&pa, chains[i].regime_probs) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer) {
        b.d_probs = static_cast<double*>(pa.devicePointer);
        b.probs_mapped = true;
      } else {
        cudaGetLastError();
        HYG_CUDA(c, pool_alloc(c, reinterpret_cast<void**>(&b.d_probs), T * (R + 1) * sizeof(double)));
      }
      if (chains[i].positions) {
        HYG_CUDA(c, pool_alloc(c, reinterpret_cast<void**>(&b.d_pos), T * sizeof(uint32_t)));
        HYG_CUDA(c, cudaMemcpyAsync(b.d_pos, chains[i].positions, T * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
      }
    }
    if (chains[i].logz) HYG_CUDA(c, pool_alloc(c, reinterpret_cast<void**>(&b.d_logz), T * sizeof(double)));
    if (chains[i].k_kept) HYG_CUDA(c, pool_alloc(c, reinterpret_cast<void**>(&b.d_k), T * sizeof(int)));
    if (chains[i].drew_uniform) HYG_CUDA(c, pool_alloc(c, reinterpret_cast<void**>(&b.d_drew), T));
    if (chains[i].n_pending) HYG_CUDA(c, pool_alloc(c, reinterpret_cast<void**>(&b.d_npend), T * sizeof(int)));
    if (chains[i].n_curr) HYG_CUDA(c, pool_alloc(c, reinterpret_cast<void**>(&b.d_ncurr), T * sizeof(int)));
    if (chains[i].finalised_at) HYG_CUDA(c, pool_alloc(c, reinterpret_cast<void**>(&b.d_fin), T * sizeof(int)));
    if (chains[i].support_hash) HYG_CUDA(c, pool_alloc(c, reinterpret_cast<void**>(&b.d_hash), T * sizeof(unsigned long long)));
    if (chains[i].tie_flags) HYG_CUDA(c, pool_alloc(c, reinterpret_cast<void**>(&b.d_tie), T));
    HYG_CUDA(c, pool_alloc(c, reinterpret_cast<void**>(&b.d_status), HYG_SG_STATUS_WORDS * sizeof(int)));
    HYG_CUDA(c, pool_alloc(c, reinterpret_cast<void**>(&b.d_ovlmax), sizeof(unsigned long long)));
  }
  // launch order: longest chain first (LPT), so the persistent CTAs finish together
  c->order.resize(n);
  std::iota(c->order.begin(), c->order.end(), 0u);
  std::stable_sort(c->order.begin(), c->order.end(), [&](uint32_t a, uint32_t b) { return c->chains[a].T > c->chains[b].T; });
  c->n_particles_staged = 0;
  guard.ok = true;
  return HYG_OK;
}

int hyg_sg_emission(hyg_ctx* c) {
  if (!c) return HYG_ERR_ARG;
  if (!c->model_set) return fail(c, HYG_ERR_STATE, "hyg_sg_set_model first");
  if (c->ds.empty()) return fail(c, HYG_ERR_STATE, "no data sets");
  HYG_CUDA(c, cudaSetDevice(c->device));
  const int R = c->hm.R;
  const size_t rows_max = static_cast<size_t>(HYG_EM_SMEM_DOUBLES - 2) / R;
  int nmax_smem = 0;
  while (static_cast<size_t>(nmax_smem + 2) * (nmax_smem + 3) / 2 <= rows_max && nmax_smem + 1 <= c->nmax_table) nmax_smem++;
  const size_t smem = static_cast<size_t>(HYG_EM_SMEM_DOUBLES + 2) * sizeof(double);
  // one persistent launch over the flattened (data set, tile of 1024 site pairs) space
  std::vector<hyg::SgEmissionSet> sets(c->ds.size());
  unsigned long long tiles = 0;
  for (size_t k = 0; k < c->ds.size(); k++) {
    const Dataset& d = c->ds[k];
    hyg::SgEmissionSet& e = sets[k];
    e.T = d.T; e.pitch = d.pitch; e.n_total = d.d_nt; e.n_meth = d.d_nm; e.logobs = d.d_logobs; e.tile0 = tiles; e.S = d.S; e.pad_ = 0;
    tiles += ((d.T + 1) / 2 + HYG_EM_TILE - 1) / HYG_EM_TILE;
  }
  if (c->d_sets_cap < sets.size()) {
    dfree(c->d_sets);
    HYG_CUDA(c, cudaMalloc(&c->d_sets, sets.size() * sizeof(hyg::SgEmissionSet)));
    c->d_sets_cap = sets.size();
  }
  HYG_CUDA(c, cudaMemcpyAsync(c->d_sets, sets.data(), sets.size() * sizeof(hyg::SgEmissionSet), cudaMemcpyHostToDevice, c->stream));
  HYG_CUDA(c, cudaStreamSynchronize(c->stream));  // `sets` is pageable and goes out of scope
  hyg::SgEmissionArgs a;
  a.sets = c->d_sets; a.n_sets = static_cast<uint32_t>(sets.size()); a.n_tiles = tiles;
  a.table = c->d_emtab; a.nmax_table = c->nmax_table; a.nmax_smem = nmax_smem;
  for (int r = 0; r < HYG_RMAX; r++) { a.alpha[r] = r < R ? c->hm.alpha[r] : 1.0; a.beta[r] = r < R ? c->hm.beta[r] : 1.0; }
  const int grid = static_cast<int>(std::min<unsigned long long>(tiles, static_cast<unsigned long long>(c->num_sms)));
  HYG_CUDA(c, cudaEventRecord(c->ev_em0, c->stream));
  int rc = HYG_ERR_UNSUPPORTED;
  HYG_DISPATCH_R(R, rc = launch_emission<RR>(c, a, smem, grid));
  if (rc) return rc;
  c->em_launches = 1;
  HYG_CUDA(c, cudaEventRecord(c->ev_em1, c->stream));
  c->timed_em = true;
  for (auto& d : c->ds) d.emitted = true;
  return HYG_OK;
}

int hyg_sg_filter(hyg_ctx* c, const hyg_sg_run_args* args) {
  if (!c || !args) return HYG_ERR_ARG;
  if (!c->theta_set) return fail(c, HYG_ERR_STATE, "hyg_sg_set_theta first");
  if (c->chains.empty()) return fail(c, HYG_ERR_STATE, "no chains staged");
  if (args->smc_proposal_type != 1 || args->smc_resample_type != 2)
    return fail(c, HYG_ERR_UNSUPPORTED, "only the change-point proposal (1) with optimal finite-state resampling (2) is implemented "
                                        "(the only combination the reference CLI uses)");
  const int R = c->hm.R;
  if (args->n_particles_max > HYG_NPMAX || args->n_particles_max < static_cast<uint32_t>(2 * R))
    return fail(c, HYG_ERR_UNSUPPORTED, "n_particles must be in [2R, 256]");
  for (auto& b : c->chains)
    if (!c->ds[b.host.dataset].emitted) return fail(c, HYG_ERR_STATE, "hyg_sg_emission has not run since a data set of these chains was added");
  const bool pe_mode = args->use_online_parameter_estimation != 0;
  if (pe_mode && R > 6) return fail(c, HYG_ERR_UNSUPPORTED, "online parameter estimation supports at most 6 regimes");
  if (pe_mode && args->n_steps_without_parameter_update == 0) return fail(c, HYG_ERR_ARG, "n_steps_without_parameter_update must be > 0");
  HYG_CUDA(c, cudaSetDevice(c->device));
  const uint32_t n = static_cast<uint32_t>(c->chains.size());
  const uint32_t Nmax = args->n_particles_max;

  c->n_particles_staged = Nmax;
  c->last_allow_forced = args->allow_forced_emission != 0;

  // model descriptor
  hyg::SgModelDev m;
  m.R = R; m.u = c->hm.u; m.n_particles = static_cast<int>(Nmax); m.dcap = c->hm.dcap;
  for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) { m.P[i][j] = c->hm.P[i][j]; m.logP[i][j] = c->hm.logP[i][j]; }
  m.tab = c->d_tab; m.tabg = c->d_tabg;
  HYG_CUDA(c, cudaMemcpyAsync(c->d_mdl, &m, sizeof(m), cudaMemcpyHostToDevice, c->stream));

  // parameter mode: initial theta on the device, theta traces, per-CTA table workspace
  uint64_t t_max = 2;
  for (auto& b : c->chains) t_max = std::max<uint64_t>(t_max, b.T);
  if (pe_mode) {
    if (!c->d_theta0) HYG_CUDA(c, cudaMalloc(&c->d_theta0, 64 * sizeof(double)));
    HYG_CUDA(c, cudaMemcpyAsync(c->d_theta0, c->hm.theta.data(), c->hm.D * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    for (auto& b : c->chains)
      if (b.host.theta_trace && !b.d_trace) HYG_CUDA(c, pool_alloc(c, reinterpret_cast<void**>(&b.d_trace), b.T * c->hm.D * sizeof(double)));
    HYG_CUDA(c, cudaStreamSynchronize(c->stream));
  }

  int occ = 1;
  if (pe_mode) { HYG_DISPATCH_R6(R, (filter_occupancy<RR, true>(&occ))); }
  else { HYG_DISPATCH_R6(R, (filter_occupancy<RR, false>(&occ))); }
  if (occ < 1) occ = 1;
  const int workers = c->num_sms * occ;
  uint64_t seg_sites = pe_mode ? 0 : c->seg_sites;
  if (seg_sites == HYG_SEGMENT_AUTO) {
    std::vector<uint64_t> lens;
    for (auto& b : c->chains) lens.push_back(b.T);
    seg_sites = choose_segment_sites(lens, c->seg_halo_left, workers);
  }
  c->seg_sites_last = seg_sites;

  // chain descriptors in launch order.  Whole-chain execution: one descriptor per chain.  Segmented execution
  // (hyg_sg_set_segmentation; not in parameter mode, where theta evolves along the chain): every chain is cut into
  // near-equal segments of <= seg_sites owned sites; a segment starts seg_halo_left sites early from the R-particle initial
  // system (the filter forgets it) and may run up to seg_halo_right sites past its end until its last owned site settles.
  const bool segmented = seg_sites > 0;
  std::vector<hyg::SgChainDev> cd;
  std::vector<hyg::SgLogzFix> fix;
  std::vector<hyg::SgOvlCheck> ovl;
  for (auto& b : c->chains) {
    b.n_seg = segmented ? static_cast<uint32_t>((b.T + seg_sites - 1) / seg_sites) : 1u;
    if (b.n_seg < 1) b.n_seg = 1;
    HYG_CUDA(c, cudaMemsetAsync(b.d_status, 0, HYG_SG_STATUS_WORDS * sizeof(int), c->stream));
    HYG_CUDA(c, cudaMemsetAsync(b.d_ovlmax, 0, sizeof(unsigned long long), c->stream));
    if (b.n_seg > 1 && b.d_probs && args->use_online_marginal_smoothing) {
      if (b.ovl_cap < b.n_seg) {
        pool_free(c, b.d_ovl);
        HYG_CUDA(c, pool_alloc(c, reinterpret_cast<void**>(&b.d_ovl), static_cast<size_t>(b.n_seg) * HYG_OVL_ROWS * R * sizeof(double)));
        b.ovl_cap = b.n_seg;
      }
      HYG_CUDA(c, cudaMemsetAsync(b.d_ovl, 0xFF, static_cast<size_t>(b.n_seg) * HYG_OVL_ROWS * R * sizeof(double), c->stream));   // NaN = not finalised
    }
    if (b.n_seg > 1 && b.seginc_cap < b.n_seg) {
      pool_free(c, b.d_seginc);
      HYG_CUDA(c, pool_alloc(c, reinterpret_cast<void**>(&b.d_seginc), b.n_seg * sizeof(double)));
      b.seginc_cap = b.n_seg;
    }
  }
  for (uint32_t k = 0; k < n; k++) {
    const ChainBuf& b = c->chains[c->order[k]];
    const uint64_t seg_len = (b.T + b.n_seg - 1) / b.n_seg;
    for (uint32_t j = 0; j < b.n_seg; j++) {
      const uint64_t t0 = j * seg_len, t1 = std::min<uint64_t>(b.T, t0 + seg_len);
      if (t0 >= t1) break;
      const bool last = (t1 == b.T);
      const uint64_t a = (b.n_seg == 1 || t0 < c->seg_halo_left) ? 0 : t0 - c->seg_halo_left;
      const uint64_t e = last ? b.T : std::min<uint64_t>(b.T, t1 + c->seg_halo_right);
      hyg::SgChainDev d;
      std::memset(&d, 0, sizeof(d));
      d.T = e - a;
      d.t_off = a; d.own_lo = t0 - a; d.own_hi = t1 - a; d.last_segment = (e == b.T) ? 1 : 0;
      d.logobs = c->ds[b.host.dataset].d_logobs + a * R;
      d.unif = b.d_unif ? b.d_unif + a : nullptr; d.seed = b.host.seed; d.chain_id = b.host.chain_id;
      d.probs = (args->use_online_marginal_smoothing && b.d_probs) ? b.d_probs + a * (R + 1) : nullptr;
      d.pos = b.d_pos ? b.d_pos + a : nullptr;
      d.logz = b.d_logz ? b.d_logz + a : nullptr;
      d.k_kept = b.d_k ? b.d_k + a : nullptr; d.drew = b.d_drew ? b.d_drew + a : nullptr;
      d.n_pending = b.d_npend ? b.d_npend + a : nullptr; d.n_curr = b.d_ncurr ? b.d_ncurr + a : nullptr;
      d.finalised_at = b.d_fin ? b.d_fin + a : nullptr;
      d.support_hash = b.d_hash ? b.d_hash + a : nullptr; d.tie_flags = b.d_tie ? b.d_tie + a : nullptr;
      d.status = b.d_status;
      if (b.n_seg > 1 && !last && b.d_ovl && d.probs) {
        d.ovl = b.d_ovl + static_cast<size_t>(j) * HYG_OVL_ROWS * R;
        hyg::SgOvlCheck oc;
        oc.ovl = d.ovl; oc.probs = b.d_probs + t1 * (R + 1); oc.max_bits = b.d_ovlmax; oc.status = b.d_status;
        oc.rows = static_cast<unsigned int>(std::min<uint64_t>(HYG_OVL_ROWS, e - t1)); oc.R = static_cast<unsigned int>(R);
        ovl.push_back(oc);
      }
      d.theta0 = c->d_theta0; d.theta_trace = pe_mode ? b.d_trace : nullptr;
      d.seg_inc = (b.n_seg > 1) ? b.d_seginc + j : nullptr;
      cd.push_back(d);
      if (j > 0 && b.d_logz) {
        hyg::SgLogzFix f;
        f.logz = b.d_logz + t0; f.seg_inc = b.d_seginc; f.len = t1 - t0; f.j = j; f.pad_ = 0;
        fix.push_back(f);
      }
    }
  }
  // longest first (LPT) over all segments of all chains
  std::stable_sort(cd.begin(), cd.end(), [](const hyg::SgChainDev& x, const hyg::SgChainDev& y) { return x.T > y.T; });
  const uint32_t n_units = static_cast<uint32_t>(cd.size());
  if (c->d_chains_cap < n_units) {
    dfree(c->d_chains);
    HYG_CUDA(c, cudaMalloc(&c->d_chains, n_units * sizeof(hyg::SgChainDev)));
    c->d_chains_cap = n_units;
  }
  HYG_CUDA(c, cudaMemcpyAsync(c->d_chains, cd.data(), n_units * sizeof(hyg::SgChainDev), cudaMemcpyHostToDevice, c->stream));
  if (!fix.empty()) {
    if (c->d_fix_cap < fix.size()) {
      dfree(c->d_fix);
      HYG_CUDA(c, cudaMalloc(&c->d_fix, fix.size() * sizeof(hyg::SgLogzFix)));
      c->d_fix_cap = fix.size();
    }
    HYG_CUDA(c, cudaMemcpyAsync(c->d_fix, fix.data(), fix.size() * sizeof(hyg::SgLogzFix), cudaMemcpyHostToDevice, c->stream));
  }
  if (!ovl.empty()) {
    if (c->d_ovlchk_cap < ovl.size()) {
      dfree(c->d_ovlchk);
      HYG_CUDA(c, cudaMalloc(&c->d_ovlchk, ovl.size() * sizeof(hyg::SgOvlCheck)));
      c->d_ovlchk_cap = ovl.size();
    }
    HYG_CUDA(c, cudaMemcpyAsync(c->d_ovlchk, ovl.data(), ovl.size() * sizeof(hyg::SgOvlCheck), cudaMemcpyHostToDevice, c->stream));
  }
  HYG_CUDA(c, cudaStreamSynchronize(c->stream));  // cd / fix / ovl are pageable
  c->n_units_last = n_units;
  const int grid = static_cast<int>(std::min<uint64_t>(n_units, static_cast<uint64_t>(workers)));
  c->grid_last = static_cast<uint32_t>(grid);

  hyg::SgRunDev run;
  run.use_smoothing = args->use_online_marginal_smoothing ? 1 : 0;
  run.epsilon = args->epsilon;
  run.lcap = args->lag_capacity ? static_cast<int>(args->lag_capacity) : 1024;
  run.psi_stride = static_cast<unsigned long long>(run.lcap) * R * HYG_NPMAX + (5ull * run.lcap + 1) / 2 + 8;
  run.force_full_sort = args->resample_full_sort ? 1 : 0;
  const size_t need = run.psi_stride * sizeof(double) * grid;
  if (need > c->psi_bytes) {
    dfree(c->d_psi);
    HYG_CUDA(c, cudaMalloc(&c->d_psi, need));
    c->psi_bytes = need;
  }
  run.psi_ws = c->d_psi;
  run.queue = c->d_queue;
  run.n_chains = static_cast<int>(n_units);
  run.use_param_est = pe_mode ? 1 : 0;
  run.normalise_gradients = args->normalise_gradients;
  run.use_adam = args->use_adam;
  run.n_steps_without_update = args->n_steps_without_parameter_update;
  run.lr_exponent = args->learning_rate_exponent;
  run.lr_factor = args->learning_rate_factor;
  for (int r = 0; r < HYG_RMAX; r++) run.kappa[r] = r < R ? c->hm.kappa[r] : 1.0;
  run.pe_ws = nullptr; run.pe_stride = 0; run.pe_dcap = 0;
  if (pe_mode) {
    run.pe_dcap = static_cast<uint32_t>(std::min<uint64_t>(std::max<uint64_t>(t_max + 8, 64), 65536));
    run.pe_stride = 5ull * R * run.pe_dcap;   // tab (2) + tabg + wh + wg
    const size_t pe_need = run.pe_stride * sizeof(double) * grid;
    if (pe_need > c->pe_bytes) {
      dfree(c->d_pe);
      HYG_CUDA(c, cudaMalloc(&c->d_pe, pe_need));
      c->pe_bytes = pe_need;
    }
    run.pe_ws = c->d_pe;
  }
  HYG_CUDA(c, cudaMemsetAsync(c->d_queue, 0, sizeof(unsigned int), c->stream));

  c->f_launches = 0;
  HYG_CUDA(c, cudaEventRecord(c->ev_f0, c->stream));
  int rc = HYG_ERR_UNSUPPORTED;
  if (pe_mode) { HYG_DISPATCH_R6(R, (rc = launch_filter<RR, true>(c, run, grid))); }
  else { HYG_DISPATCH_R6(R, (rc = launch_filter<RR, false>(c, run, grid))); }
  if (rc) return rc;
  c->f_launches++;
  if (!ovl.empty()) {
    sg_overlap_check_kernel<<<std::min<unsigned int>(static_cast<unsigned int>(ovl.size()), 1024u), HYG_OVL_ROWS, 0, c->stream>>>(c->d_ovlchk, static_cast<unsigned int>(ovl.size()));
    HYG_CUDA(c, cudaGetLastError());
    c->f_launches++;
  }
  if (!fix.empty()) {
    sg_logz_fix_kernel<<<c->num_sms * 4, 256, 0, c->stream>>>(c->d_fix, static_cast<unsigned int>(fix.size()));
    HYG_CUDA(c, cudaGetLastError());
    c->f_launches++;
  }
  HYG_CUDA(c, cudaEventRecord(c->ev_f1, c->stream));
  c->timed_f = true;
  return HYG_OK;
}

int hyg_sg_device_outputs(hyg_ctx* c, uint32_t chain, double** regime_probs, double** logz) {
  if (!c || chain >= c->chains.size()) return fail(c, HYG_ERR_ARG, "chain index out of range");
  const ChainBuf& b = c->chains[chain];
  if (regime_probs) *regime_probs = b.probs_mapped ? nullptr : b.d_probs;
  if (logz) *logz = b.d_logz;
  return HYG_OK;
}

int hyg_sg_download(hyg_ctx* c, hyg_sg_chain* chains, uint32_t n) {
  if (!c) return HYG_ERR_ARG;
  if (n != c->chains.size()) return fail(c, HYG_ERR_ARG, "chain count differs from hyg_sg_set_chains");
  HYG_CUDA(c, cudaSetDevice(c->device));
  const int R = c->hm.R;
  for (uint32_t i = 0; i < n; i++) {
    ChainBuf& b = c->chains[i];
    const hyg_sg_chain& h = b.host;
    const uint64_t T = b.T;
    if (h.regime_probs && b.d_probs && !b.probs_mapped) HYG_CUDA(c, cudaMemcpyAsync(h.regime_probs, b.d_probs, T * (R + 1) * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (h.logz) HYG_CUDA(c, cudaMemcpyAsync(h.logz, b.d_logz, T * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (h.k_kept) HYG_CUDA(c, cudaMemcpyAsync(h.k_kept, b.d_k, T * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (h.drew_uniform) HYG_CUDA(c, cudaMemcpyAsync(h.drew_uniform, b.d_drew, T, cudaMemcpyDeviceToHost, c->stream));
    if (h.n_pending) HYG_CUDA(c, cudaMemcpyAsync(h.n_pending, b.d_npend, T * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (h.n_curr) HYG_CUDA(c, cudaMemcpyAsync(h.n_curr, b.d_ncurr, T * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (h.finalised_at) HYG_CUDA(c, cudaMemcpyAsync(h.finalised_at, b.d_fin, T * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (h.theta_trace && b.d_trace) HYG_CUDA(c, cudaMemcpyAsync(h.theta_trace, b.d_trace, T * c->hm.D * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (h.support_hash && b.d_hash) HYG_CUDA(c, cudaMemcpyAsync(h.support_hash, b.d_hash, T * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    if (h.tie_flags && b.d_tie) HYG_CUDA(c, cudaMemcpyAsync(h.tie_flags, b.d_tie, T, cudaMemcpyDeviceToHost, c->stream));
    HYG_CUDA(c, cudaMemcpyAsync(b.host.status, b.d_status, HYG_SG_STATUS_WORDS * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    HYG_CUDA(c, cudaMemcpyAsync(&b.host.overlap_max_abs, b.d_ovlmax, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  }
  HYG_CUDA(c, cudaStreamSynchronize(c->stream));
  uint64_t forced = 0;
  for (uint32_t i = 0; i < n; i++) {
    const ChainBuf& b = c->chains[i];
    if (chains) { std::memcpy(chains[i].status, b.host.status, sizeof(b.host.status)); chains[i].overlap_max_abs = b.host.overlap_max_abs; }
    forced += static_cast<uint64_t>(b.host.status[0]);
  }
  // the results are in the caller's buffers either way; a forced emission means they are not the reference's estimator
  if (forced && !c->last_allow_forced)
    return fail(c, HYG_ERR_CAPACITY, std::to_string(forced) + " site(s) were emitted with their filtering estimate because the lag set was full: "
                                     "raise hyg_sg_run_args.lag_capacity (or set allow_forced_emission)");
  return HYG_OK;
}

int hyg_sg_timings(hyg_ctx* c, float* ms_em, float* ms_f, uint32_t* em_launches, uint32_t* f_launches) {
  if (!c) return HYG_ERR_ARG;
  HYG_CUDA(c, cudaSetDevice(c->device));
  HYG_CUDA(c, cudaStreamSynchronize(c->stream));
  if (ms_em) { *ms_em = 0.f; if (c->timed_em) HYG_CUDA(c, cudaEventElapsedTime(ms_em, c->ev_em0, c->ev_em1)); }
  if (ms_f) { *ms_f = 0.f; if (c->timed_f) HYG_CUDA(c, cudaEventElapsedTime(ms_f, c->ev_f0, c->ev_f1)); }
  if (em_launches) *em_launches = c->em_launches;
  if (f_launches) *f_launches = c->f_launches;
  return HYG_OK;
}

int hyg_sg_get_logobs(hyg_ctx* c, uint32_t dataset, double* logobs) {
  if (!c || !logobs || dataset >= c->ds.size()) return fail(c, HYG_ERR_ARG, "bad data set");
  HYG_CUDA(c, cudaSetDevice(c->device));
  const Dataset& d = c->ds[dataset];
  HYG_CUDA(c, cudaMemcpyAsync(logobs, d.d_logobs, d.T * c->hm.R * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  HYG_CUDA(c, cudaStreamSynchronize(c->stream));
  return HYG_OK;
}

int hyg_sg_run_online_combined_inference(hyg_ctx* c, const double* vartheta, uint32_t n_vartheta, const double* theta_init, uint32_t dim_theta,
                                         uint64_t T, uint32_t S, const uint32_t* positions, const uint16_t* n_total, const uint16_t* n_meth,
                                         const hyg_sg_run_args* args, uint64_t seed, const double* uniforms,
                                         double* regime_probs, double* theta_trace, double* logz, double* seconds, int32_t* status) {
  if (!c || !args) return HYG_ERR_ARG;
  const auto t0 = std::chrono::steady_clock::now();
  int rc;
  if ((rc = hyg_sg_clear(c))) return rc;
  if ((rc = hyg_sg_set_vartheta(c, vartheta, n_vartheta))) return rc;
  if ((rc = hyg_sg_set_theta(c, theta_init, dim_theta, T))) return rc;
  if ((rc = hyg_sg_add_dataset(c, T, S, n_total, n_meth, 0, T)) < 0) return rc;
  hyg_sg_chain ch;
  std::memset(&ch, 0, sizeof(ch));
  ch.dataset = 0; ch.seed = seed; ch.chain_id = 0; ch.uniforms = uniforms; ch.positions = positions;
  ch.regime_probs = args->use_online_marginal_smoothing ? regime_probs : nullptr;
  ch.logz = logz; ch.theta_trace = theta_trace;
  if ((rc = hyg_sg_set_chains(c, &ch, 1))) return rc;
  if ((rc = hyg_sg_emission(c))) return rc;
  // The reference's lag set is unbounded (OnlineMarginalSmoothing.h:119-255); here it lives in a workspace of lag_capacity
  // pending sites.  When it overflows the recursion is run again with four times the capacity (same draws, so the same
  // chain; the emission table is kept) until nothing overflowed or the capacity covers the whole chain.
  hyg_sg_run_args a = *args;
  if (!a.lag_capacity) a.lag_capacity = 1024;
  for (;;) {
    if ((rc = hyg_sg_filter(c, &a))) return rc;
    rc = hyg_sg_download(c, &ch, 1);
    if (status) std::memcpy(status, ch.status, sizeof(ch.status));
    if (rc != HYG_ERR_CAPACITY || a.lag_capacity >= T) break;
    a.lag_capacity = static_cast<uint32_t>(std::min<uint64_t>(4ull * a.lag_capacity, std::max<uint64_t>(T, 1)));
  }
  if (rc) return rc;
  if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return HYG_OK;
}


// ---- two-group path ------------------------------------------------------------------------------------------------

int hyg_tg_hazard_table(const double* omega, const double* kappa, uint32_t R, uint32_t u, uint32_t d_max, double* rho) {
  if (!omega || !kappa || !rho || R < 1 || R > HYG_RMAX || d_max < u) return HYG_ERR_ARG;
  std::vector<double> t;
  hyg::build_hazard_table(omega, kappa, static_cast<int>(R), static_cast<int>(u), d_max, t);
  std::memcpy(rho, t.data(), t.size() * sizeof(double));
  return HYG_OK;
}

int hyg_tg_reference_hazard_table(const double* omega, const double* kappa, uint32_t R, uint32_t u, uint32_t d_max, double* rho) {
  if (!omega || !kappa || !rho || R < 1 || R > HYG_RMAX || d_max < u) return HYG_ERR_ARG;
  std::vector<double> t;
  hyg::build_reference_hazard_table(omega, kappa, static_cast<int>(R), static_cast<int>(u), d_max, t);
  std::memcpy(rho, t.data(), t.size() * sizeof(double));
  return HYG_OK;
}

int hyg_tg_set_model(hyg_ctx* c, const hyg_tg_model* m, uint64_t t_max) {
  if (!c || !m) return HYG_ERR_ARG;
  if (m->R < 2 || m->R > HYG_RMAX) return fail(c, HYG_ERR_UNSUPPORTED, "two-group: R must be in 2..8");
  if (m->minimum_duration < 1) return fail(c, HYG_ERR_ARG, "two-group: minimum_duration must be >= 1");
  if (m->num_resampled < 1 || m->num_resampled > HYG_TG_MMAX) return fail(c, HYG_ERR_UNSUPPORTED, "two-group: num_resampled must be in 1..64");
  if (m->num_backward < 1 || m->num_backward > HYG_TG_BMAX) return fail(c, HYG_ERR_UNSUPPORTED, "two-group: num_backward must be in 1..32");
  const uint32_t R = m->R;
  if (static_cast<size_t>(m->num_resampled) * (2 * R + R * R) > HYG_TG_NPMAX)
    return fail(c, HYG_ERR_UNSUPPORTED, "two-group: num_resampled * (2R + R^2) exceeds the 2432 particle slots");
  if (!m->log_p_control) return fail(c, HYG_ERR_ARG, "two-group: log_p_control is required");
  if (!(m->merge_prob > 0.0 && m->merge_prob < 1.0 && m->split_prob > 0.0 && m->split_prob < 1.0))
    return fail(c, HYG_ERR_ARG, "two-group: merge_prob and split_prob must be in (0,1)");
  HYG_CUDA(c, cudaSetDevice(c->device));
  hyg::TgModelDev& h = c->tg_host;
  std::memset(&h, 0, sizeof(h));
  h.R = static_cast<int>(R); h.u = static_cast<int>(m->minimum_duration); h.M = static_cast<int>(m->num_resampled); h.B = static_cast<int>(m->num_backward);
  h.presel[0] = std::max(m->sort_preselect[0] ? static_cast<int>(m->sort_preselect[0]) : h.M + 110, h.M + 96);
  h.presel[1] = std::max(m->sort_preselect[1] ? static_cast<int>(m->sort_preselect[1]) : 3 * h.M + 250, h.presel[0]);
  h.big_from = m->sort_scratch_from ? static_cast<int>(std::min<uint32_t>(m->sort_scratch_from - 1, HYG_TG_SORTMAX)) : HYG_TG_SORTMAX;
  // tf.nn.softmax of the log-probabilities with a -inf diagonal (case_control_regime_model.py:90-95), in log space
  for (uint32_t i = 0; i < R; i++) {
    double mx = -HUGE_VAL;
    for (uint32_t j = 0; j < R; j++) if (j != i) mx = std::max(mx, m->log_p_control[i * R + j]);
    if (!std::isfinite(mx)) return fail(c, HYG_ERR_ARG, "two-group: a row of log_p_control has no finite off-diagonal entry");
    double sm = 0.0;
    for (uint32_t j = 0; j < R; j++) if (j != i) sm += std::exp(m->log_p_control[i * R + j] - mx);
    const double lz = mx + std::log(sm);
    for (uint32_t j = 0; j < R; j++) h.logP[i][j] = (j == i) ? -HUGE_VAL : m->log_p_control[i * R + j] - lz;
  }
  // rows = previous indicator (0 = split, 1 = merged); run_inference_two_groups.py:163-167
  h.logPm[0][0] = std::log(1.0 - m->merge_prob); h.logPm[0][1] = std::log(m->merge_prob);
  h.logPm[1][0] = std::log(m->split_prob);       h.logPm[1][1] = std::log(1.0 - m->split_prob);
  std::vector<double> rho_c, rho_k;
  uint32_t dmax;
  if (m->rho_control && m->rho_case) {
    dmax = m->d_max;
    if (dmax < m->minimum_duration) return fail(c, HYG_ERR_ARG, "two-group: d_max of the supplied hazard tables is below minimum_duration");
    rho_c.assign(m->rho_control, m->rho_control + static_cast<size_t>(R) * (dmax + 1));
    rho_k.assign(m->rho_case, m->rho_case + static_cast<size_t>(R) * (dmax + 1));
  } else {
    if (!m->omega_control || !m->omega_case) return fail(c, HYG_ERR_ARG, "two-group: omega_control / omega_case (or hazard tables) are required");
    for (uint32_t r = 0; r < R; r++)
      if (!(m->omega_control[r] > 0.0 && m->omega_control[r] < 1.0 && m->omega_case[r] > 0.0 && m->omega_case[r] < 1.0))
        return fail(c, HYG_ERR_ARG, "two-group: omega must be in (0,1)");
    const double two[HYG_RMAX] = {2, 2, 2, 2, 2, 2, 2, 2};
    // sojourn times beyond d_max reuse the last entry: the exact hazard of a negative binomial is flat by then, and the reference-mode
    // hazard is the constant 0.1 from the sojourn where the fp32 cdf rounds to 1 (4071 for omega = 0.995; 262144 covers omega up to 0.9999)
    dmax = static_cast<uint32_t>(std::min<uint64_t>(std::max<uint64_t>(t_max, m->minimum_duration + 1), 262144));
    if (m->hazard_mode > HYG_TG_HAZARD_EXACT) return fail(c, HYG_ERR_ARG, "two-group: hazard_mode must be HYG_TG_HAZARD_REFERENCE or HYG_TG_HAZARD_EXACT");
    auto* build = (m->hazard_mode == HYG_TG_HAZARD_EXACT) ? hyg::build_hazard_table : hyg::build_reference_hazard_table;
    build(m->omega_control, m->kappa_control ? m->kappa_control : two, h.R, h.u, dmax, rho_c);
    build(m->omega_case, m->kappa_case ? m->kappa_case : two, h.R, h.u, dmax, rho_k);
  }
  h.dmax = dmax;
  const size_t n = static_cast<size_t>(R) * (dmax + 1);
  dfree(c->d_tg_rho);
  // the kernels only ever need log rho and log(1 - rho): taken here with libm (as the oracle does), two per entry
  std::vector<double> lr(4 * n);
  for (size_t i = 0; i < n; i++) {
    lr[2 * i] = std::log(rho_c[i]); lr[2 * i + 1] = std::log(1.0 - rho_c[i]);
    lr[2 * n + 2 * i] = std::log(rho_k[i]); lr[2 * n + 2 * i + 1] = std::log(1.0 - rho_k[i]);
  }
  HYG_CUDA(c, cudaMalloc(&c->d_tg_rho, 4 * n * sizeof(double)));
  HYG_CUDA(c, cudaMemcpy(c->d_tg_rho, lr.data(), 4 * n * sizeof(double), cudaMemcpyHostToDevice));
  h.lrho_c = reinterpret_cast<const double2*>(c->d_tg_rho); h.lrho_k = reinterpret_cast<const double2*>(c->d_tg_rho + 2 * n);
  h.nl_rm1 = -std::log(static_cast<double>(R) - 1.0); h.nl_rm2 = -std::log(static_cast<double>(R) - 2.0);
  if (!c->d_tg_mdl) HYG_CUDA(c, cudaMalloc(&c->d_tg_mdl, sizeof(hyg::TgModelDev)));
  HYG_CUDA(c, cudaMemcpy(c->d_tg_mdl, &h, sizeof(h), cudaMemcpyHostToDevice));
  c->tg_set = true;
  return HYG_OK;
}

int hyg_tg_run(hyg_ctx* c, const hyg_tg_chain* chains, uint32_t n, float* ms_device) {
  if (!c || (!chains && n)) return HYG_ERR_ARG;
  if (!c->tg_set) return fail(c, HYG_ERR_STATE, "hyg_tg_set_model first");
  if (n == 0) return HYG_OK;
  HYG_CUDA(c, cudaSetDevice(c->device));
  const int R = c->tg_host.R, B = c->tg_host.B, M = c->tg_host.M;
  if (R != c->hm.R) return fail(c, HYG_ERR_STATE, "two-group: R differs from the emission model (hyg_sg_set_model)");
  uint64_t t_max = 0;
  for (uint32_t k = 0; k < n; k++) {
    const hyg_tg_chain& ch = chains[k];
    if (ch.control_dataset >= c->ds.size() || ch.case_dataset >= c->ds.size()) return fail(c, HYG_ERR_ARG, "two-group: data-set index out of range");
    const Dataset& a = c->ds[ch.control_dataset];
    const Dataset& b = c->ds[ch.case_dataset];
    if (a.T != b.T) return fail(c, HYG_ERR_ARG, "two-group: control and case data sets differ in length");
    if (!a.d_logobs || !b.d_logobs || !a.emitted || !b.emitted) return fail(c, HYG_ERR_STATE, "two-group: hyg_sg_emission has not run since the data sets were added");
    if (a.T == 0) return fail(c, HYG_ERR_ARG, "two-group: empty data set");
    if (!ch.trajectories || !ch.log_normalizing_constant) return fail(c, HYG_ERR_ARG, "two-group: output pointers are required");
    t_max = std::max<uint64_t>(t_max, a.T);
  }
  // longest chain first (one CTA per chain, dynamic queue)
  std::vector<uint32_t> order(n);
  std::iota(order.begin(), order.end(), 0u);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return c->ds[chains[x].control_dataset].T > c->ds[chains[y].control_dataset].T; });
  std::vector<hyg::TgChainDev> dev(n);
  std::vector<int*> d_traj(n, nullptr), d_taps(n, nullptr), traj_mapped(n, nullptr);
  double* d_ln = nullptr;
  hyg::TgChainDev* d_chains = nullptr;
  unsigned char* d_ws = nullptr;
  int rc = HYG_OK;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  auto cleanup = [&]() {
    for (auto& p : d_traj) pool_free(c, p);
    for (auto& p : d_taps) pool_free(c, p);
    pool_free(c, d_ln); pool_free(c, d_chains); pool_free(c, d_ws);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
  };
#define HYG_TG_CUDA(call)                                                                       \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess) { cleanup(); return fail(c, HYG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } \
  } while (0)
  HYG_TG_CUDA(pool_alloc(c, reinterpret_cast<void**>(&d_ln), n * sizeof(double)));
  for (uint32_t i = 0; i < n; i++) {
    const uint32_t k = order[i];
    const hyg_tg_chain& ch = chains[k];
    const uint64_t T = c->ds[ch.control_dataset].T;
    // pinned (page-locked) caller memory is addressable from the device: the backward pass then writes every site's 500 bytes
    // of trajectories straight into it (posted writes over PCIe, ~4 GB/s of the link) and there is no staging copy afterwards
    {
      cudaPointerAttributes pa;
      std::memset(&pa, 0, sizeof(pa));
      if (c->zero_copy_out && cudaPointerGetAttributes(&pa, ch.trajectories) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer) {
        traj_mapped[k] = static_cast<int*>(pa.devicePointer);
      } else {
        cudaGetLastError();
        HYG_TG_CUDA(pool_alloc(c, reinterpret_cast<void**>(&d_traj[k]), T * B * 5 * sizeof(int)));
      }
    }
    if (ch.taps) HYG_TG_CUDA(pool_alloc(c, reinterpret_cast<void**>(&d_taps[k]), T * 4 * sizeof(int)));
    hyg::TgChainDev& d = dev[i];
    d.T = T; d.lo_c = c->ds[ch.control_dataset].d_logobs; d.lo_k = c->ds[ch.case_dataset].d_logobs;
    d.seed = ch.seed; d.chain = ch.chain_id; d.traj = traj_mapped[k] ? traj_mapped[k] : d_traj[k]; d.log_norm = d_ln + k; d.taps = d_taps[k];
  }
  hyg::TgRunDev run;
  run.t_max = t_max;
  run.anc_pitch = static_cast<unsigned long long>(std::max(M, R * R));
  run.scratch_off = (sizeof(hyg::TgStepRec) * t_max + sizeof(hyg::TgAncRec) * t_max * run.anc_pitch + 255) & ~static_cast<size_t>(255);
  run.ws_stride = (run.scratch_off + static_cast<size_t>(HYG_TG_BIGMAX) * 18 + 255) & ~static_cast<size_t>(255);
  run.n_chains = static_cast<int>(n);
  run.queue = c->d_queue;
  // HYG_TG_CTAS CTAs per SM; fewer when the per-CTA ancestor history would not fit in what is free
  size_t free_b = 0, total_b = 0;
  HYG_TG_CUDA(cudaMemGetInfo(&free_b, &total_b));
  int grid = static_cast<int>(std::min<uint32_t>(n, static_cast<uint32_t>(c->num_sms) * HYG_TG_CTAS));
  const size_t fit = static_cast<size_t>(0.8 * static_cast<double>(free_b)) / run.ws_stride;
  if (fit < 1) { cleanup(); return fail(c, HYG_ERR_CUDA, "two-group: not enough device memory for one chain's ancestor history"); }
  grid = static_cast<int>(std::min<size_t>(static_cast<size_t>(grid), fit));
  HYG_TG_CUDA(pool_alloc(c, reinterpret_cast<void**>(&d_ws), run.ws_stride * grid));
  run.ws = d_ws;
  HYG_TG_CUDA(pool_alloc(c, reinterpret_cast<void**>(&d_chains), n * sizeof(hyg::TgChainDev)));
  HYG_TG_CUDA(cudaMemcpyAsync(d_chains, dev.data(), n * sizeof(hyg::TgChainDev), cudaMemcpyHostToDevice, c->stream));
  HYG_TG_CUDA(cudaMemsetAsync(c->d_queue, 0, sizeof(unsigned int), c->stream));
  HYG_TG_CUDA(cudaFuncSetAttribute(hyg::tg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(hyg::TgSmem))));
  HYG_TG_CUDA(cudaEventCreate(&e0));
  HYG_TG_CUDA(cudaEventCreate(&e1));
  HYG_TG_CUDA(cudaEventRecord(e0, c->stream));
  hyg::tg_kernel<<<grid, HYG_TG_NT, sizeof(hyg::TgSmem), c->stream>>>(c->d_tg_mdl, d_chains, run);
  HYG_TG_CUDA(cudaGetLastError());
  HYG_TG_CUDA(cudaEventRecord(e1, c->stream));
  for (uint32_t k = 0; k < n; k++) {
    const hyg_tg_chain& ch = chains[k];
    const uint64_t T = c->ds[ch.control_dataset].T;
    if (!traj_mapped[k]) HYG_TG_CUDA(cudaMemcpyAsync(ch.trajectories, d_traj[k], T * B * 5 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    HYG_TG_CUDA(cudaMemcpyAsync(ch.log_normalizing_constant, d_ln + k, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (ch.taps) HYG_TG_CUDA(cudaMemcpyAsync(ch.taps, d_taps[k], T * 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  }
  HYG_TG_CUDA(cudaStreamSynchronize(c->stream));
  if (ms_device) HYG_TG_CUDA(cudaEventElapsedTime(ms_device, e0, e1));
#undef HYG_TG_CUDA
  cleanup();
  return rc;
}

double hyg_philox_uniform(uint64_t seed, uint32_t chain_id, uint64_t t) { return hyg::philox_uniform(seed, chain_id, t); }

// ---- DMP calling: per-site statistics of the aggregated trajectories (K6) and the two FDR procedures ------------------------

int hyg_tg_site_statistics(hyg_ctx* c, uint64_t T, uint32_t P, uint32_t R, const int8_t* merged, const int8_t* control_regimes,
                           const int8_t* case_regimes, int on_device, double* split_prob, double* null_stat, double* control_freq,
                           double* case_freq, double* pair_stat, float* ms_device) {
  if (!c || !merged || !control_regimes || !case_regimes || !split_prob || !null_stat) return fail(c, HYG_ERR_ARG, "null argument");
  if (T == 0 || P == 0 || R == 0 || R > 8) return fail(c, HYG_ERR_ARG, "need T > 0, P > 0, 1 <= R <= 8");
  const size_t smem = 16 + 3 * ((static_cast<size_t>(HYG_DMP_TILE) * P + 15) / 16 * 16) + (static_cast<size_t>(P) + 2) * sizeof(double);
  if (smem > 227 * 1024) return fail(c, HYG_ERR_UNSUPPORTED, "too many particles per site for the shared-memory tile");
  HYG_CUDA(c, cudaSetDevice(c->device));
  const size_t nb = static_cast<size_t>(T) * P;
  const int8_t* d_in[3] = {merged, control_regimes, case_regimes};
  int8_t* own_in[3] = {nullptr, nullptr, nullptr};
  double *d_split = nullptr, *d_null = nullptr, *d_cf = nullptr, *d_kf = nullptr, *d_pair = nullptr;
  auto cleanup = [&]() {
    for (auto& p : own_in) pool_free(c, p);
    pool_free(c, d_split); pool_free(c, d_null); pool_free(c, d_cf); pool_free(c, d_kf); pool_free(c, d_pair);
  };
#define HYG_DMP_CUDA(call)                                                                                          \
  do {                                                                                                              \
    cudaError_t e_ = (call);                                                                                        \
    if (e_ != cudaSuccess) { cleanup(); return fail(c, HYG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } \
  } while (0)
  if (!on_device) {
    for (int k = 0; k < 3; k++) {
      HYG_DMP_CUDA(pool_alloc(c, reinterpret_cast<void**>(&own_in[k]), nb + 16));
      HYG_DMP_CUDA(cudaMemcpyAsync(own_in[k], d_in[k], nb, cudaMemcpyHostToDevice, c->stream));
      d_in[k] = own_in[k];
    }
  }
  // outputs: device pointers when the inputs are device resident, else staged
  double* o_split = split_prob; double* o_null = null_stat; double* o_cf = control_freq; double* o_kf = case_freq; double* o_pair = pair_stat;
  if (!on_device) {
    HYG_DMP_CUDA(pool_alloc(c, reinterpret_cast<void**>(&d_split), T * sizeof(double))); o_split = d_split;
    HYG_DMP_CUDA(pool_alloc(c, reinterpret_cast<void**>(&d_null), T * sizeof(double))); o_null = d_null;
    if (control_freq) { HYG_DMP_CUDA(pool_alloc(c, reinterpret_cast<void**>(&d_cf), T * R * sizeof(double))); o_cf = d_cf; }
    if (case_freq) { HYG_DMP_CUDA(pool_alloc(c, reinterpret_cast<void**>(&d_kf), T * R * sizeof(double))); o_kf = d_kf; }
    if (pair_stat) { HYG_DMP_CUDA(pool_alloc(c, reinterpret_cast<void**>(&d_pair), T * R * R * sizeof(double))); o_pair = d_pair; }
  }
  hyg::DmpArgs a;
  a.T = T; a.P = P; a.R = R; a.merged = reinterpret_cast<const signed char*>(d_in[0]);
  a.control = reinterpret_cast<const signed char*>(d_in[1]); a.cse = reinterpret_cast<const signed char*>(d_in[2]);
  a.split_prob = o_split; a.null_stat = o_null; a.control_freq = o_cf; a.case_freq = o_kf; a.pair_stat = o_pair;
  a.n_tiles = (T + HYG_DMP_TILE - 1) / HYG_DMP_TILE;
  HYG_DMP_CUDA(cudaFuncSetAttribute(hyg::dmp_site_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hyg::dmp_site_stats_kernel, HYG_DMP_NT, smem);
  if (occ < 1) occ = 1;
  const int grid = static_cast<int>(std::min<unsigned long long>(a.n_tiles, static_cast<unsigned long long>(c->num_sms) * occ));
  HYG_DMP_CUDA(cudaEventRecord(c->ev_d0, c->stream));
  hyg::dmp_site_stats_kernel<<<grid, HYG_DMP_NT, smem, c->stream>>>(a);
  HYG_DMP_CUDA(cudaGetLastError());
  HYG_DMP_CUDA(cudaEventRecord(c->ev_d1, c->stream));
  if (!on_device) {
    HYG_DMP_CUDA(cudaMemcpyAsync(split_prob, d_split, T * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    HYG_DMP_CUDA(cudaMemcpyAsync(null_stat, d_null, T * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (control_freq) HYG_DMP_CUDA(cudaMemcpyAsync(control_freq, d_cf, T * R * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (case_freq) HYG_DMP_CUDA(cudaMemcpyAsync(case_freq, d_kf, T * R * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (pair_stat) HYG_DMP_CUDA(cudaMemcpyAsync(pair_stat, d_pair, T * R * R * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  }
  HYG_DMP_CUDA(cudaStreamSynchronize(c->stream));
  if (ms_device) HYG_DMP_CUDA(cudaEventElapsedTime(ms_device, c->ev_d0, c->ev_d1));
  cleanup();
  return HYG_OK;
}

// ---- device primitives of the two procedures: hand-written kernels in hyg_dmp.cuh (radix sort, scan, count) ----
extern "C++" {
namespace {
struct DmpScratch {
  hyg_ctx* c;
  std::vector<void*> bufs;
  explicit DmpScratch(hyg_ctx* ctx) : c(ctx) {}
  ~DmpScratch() { for (void* p : bufs) pool_free(c, p); }
  template <class T> cudaError_t get(T** p, size_t count) {
    void* q = nullptr;
    const cudaError_t e = pool_alloc(c, &q, std::max<size_t>(count, 1) * sizeof(T));
    if (e == cudaSuccess) { bufs.push_back(q); *p = static_cast<T*>(q); }
    return e;
  }
};

// stable ascending sort of the n keys (and, if pay != nullptr, their payloads) in place; k2 / p2 are scratch of the same size
cudaError_t dmp_radix_sort(hyg_ctx* c, DmpScratch& sc, unsigned long long* keys, unsigned long long* k2, unsigned long long* pay, unsigned long long* p2, uint64_t n) {
  const unsigned int n_tiles = static_cast<unsigned int>(std::min<uint64_t>((n + 2047) / 2048, static_cast<uint64_t>(c->num_sms) * 4));
  const unsigned long long tile = (n + n_tiles - 1) / n_tiles;
  unsigned int* d_counts = nullptr;
  unsigned long long* d_offs = nullptr;
  cudaError_t e;
  if ((e = sc.get(&d_counts, 256ull * n_tiles)) != cudaSuccess) return e;
  if ((e = sc.get(&d_offs, 256ull * n_tiles)) != cudaSuccess) return e;
  unsigned long long *src_k = keys, *dst_k = k2, *src_p = pay, *dst_p = p2;
  for (int pass = 0; pass < 8; pass++) {
    const int shift = 8 * pass;
    hyg::dmp_radix_hist_kernel<<<n_tiles, HYG_RS_NT, 0, c->stream>>>(src_k, n, tile, shift, d_counts, n_tiles);
    hyg::dmp_scan_counts_kernel<<<1, 1024, 0, c->stream>>>(d_counts, d_offs, 256u * n_tiles);
    if (pay) hyg::dmp_radix_scatter_kernel<true><<<n_tiles, HYG_RS_NT, 0, c->stream>>>(src_k, src_p, dst_k, dst_p, n, tile, shift, d_offs, n_tiles);
    else hyg::dmp_radix_scatter_kernel<false><<<n_tiles, HYG_RS_NT, 0, c->stream>>>(src_k, nullptr, dst_k, nullptr, n, tile, shift, d_offs, n_tiles);
    std::swap(src_k, dst_k);
    std::swap(src_p, dst_p);
  }
  return cudaGetLastError();   // eight passes: the result is back in keys / pay
}

cudaError_t dmp_inclusive_scan(hyg_ctx* c, DmpScratch& sc, const double* x, double* out, uint64_t n) {
  const unsigned int n_tiles = static_cast<unsigned int>(std::min<uint64_t>((n + 4095) / 4096, static_cast<uint64_t>(c->num_sms) * 4));
  unsigned long long tile = (n + n_tiles - 1) / n_tiles;
  tile = (tile + HYG_RS_NT * HYG_SCAN_ITEMS - 1) / (HYG_RS_NT * HYG_SCAN_ITEMS) * (HYG_RS_NT * HYG_SCAN_ITEMS);   // whole chunks per tile
  const unsigned int used = static_cast<unsigned int>((n + tile - 1) / tile);
  double* d_sums = nullptr;
  cudaError_t e;
  if ((e = sc.get(&d_sums, used)) != cudaSuccess) return e;
  hyg::dmp_tile_sum_kernel<<<used, HYG_RS_NT, 0, c->stream>>>(x, n, tile, d_sums);
  hyg::dmp_scan_tile_sums_kernel<<<1, 32, 0, c->stream>>>(d_sums, used);
  hyg::dmp_tile_scan_kernel<<<used, HYG_RS_NT, 0, c->stream>>>(x, out, n, tile, d_sums);
  return cudaGetLastError();
}

cudaError_t dmp_count_le(hyg_ctx* c, DmpScratch& sc, const double* x, uint64_t n, double bound, uint64_t* out) {
  unsigned long long* d_cnt = nullptr;
  cudaError_t e;
  if ((e = sc.get(&d_cnt, 1)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), c->stream)) != cudaSuccess) return e;
  hyg::dmp_count_le_kernel<<<c->num_sms * 4, 256, 0, c->stream>>>(x, n, bound, d_cnt);
  unsigned long long h = 0;
  if ((e = cudaMemcpyAsync(&h, d_cnt, sizeof(h), cudaMemcpyDeviceToHost, c->stream)) != cudaSuccess) return e;
  if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) return e;
  *out = h;
  return cudaGetLastError();
}
}  // namespace
}  // extern "C++"

#define HYG_FDR_CUDA(call)                                                                                          \
  do {                                                                                                              \
    cudaError_t e_ = (call);                                                                                        \
    if (e_ != cudaSuccess) return fail(c, HYG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));        \
  } while (0)

int hyg_fdr_procedure(hyg_ctx* c, uint64_t n, const double* test_statistics, double fdr_threshold, uint64_t* k, double* Qk, double* threshold) {
  if (!c || !test_statistics || !k || !Qk || !threshold || n == 0) return fail(c, HYG_ERR_ARG, "null argument / empty input");
  HYG_CUDA(c, cudaSetDevice(c->device));
  DmpScratch sc(c);
  double *d_t = nullptr, *d_q = nullptr;
  unsigned long long *d_k = nullptr, *d_k2 = nullptr;
  HYG_FDR_CUDA(sc.get(&d_t, n)); HYG_FDR_CUDA(sc.get(&d_q, n)); HYG_FDR_CUDA(sc.get(&d_k, n)); HYG_FDR_CUDA(sc.get(&d_k2, n));
  HYG_FDR_CUDA(cudaMemcpyAsync(d_t, test_statistics, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  hyg::dmp_make_keys_kernel<<<c->num_sms * 4, 256, 0, c->stream>>>(d_t, d_k, n);
  HYG_FDR_CUDA(dmp_radix_sort(c, sc, d_k, d_k2, nullptr, nullptr, n));                       // np.sort (multiple_testing.py:4)
  hyg::dmp_keys_to_doubles_kernel<<<c->num_sms * 4, 256, 0, c->stream>>>(d_k, d_t, n);
  HYG_FDR_CUDA(dmp_inclusive_scan(c, sc, d_t, d_q, n));                                      // np.cumsum
  hyg::dmp_running_mean_kernel<<<c->num_sms * 4, 256, 0, c->stream>>>(d_q, d_q, n);          // Qs (:5-6)
  uint64_t s = 0;
  HYG_FDR_CUDA(dmp_count_le(c, sc, d_q, n, fdr_threshold, &s));                              // :7
  double first = 0.0;
  HYG_FDR_CUDA(cudaMemcpyAsync(&first, d_t, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  HYG_FDR_CUDA(cudaStreamSynchronize(c->stream));
  if (fdr_threshold < first) {                                   // :8-9
    *k = 0; *Qk = 0.0; *threshold = 0.0;
  } else {
    // s >= 1 here (Qs[0] = sorted[0] <= fdr_threshold)
    HYG_FDR_CUDA(cudaMemcpyAsync(Qk, d_q + (s - 1), sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (s == n) *threshold = 1.01;                               // :10-11
    else HYG_FDR_CUDA(cudaMemcpyAsync(threshold, d_t + s, sizeof(double), cudaMemcpyDeviceToHost, c->stream));   // :12
    HYG_FDR_CUDA(cudaStreamSynchronize(c->stream));
    *k = s;
  }
  return HYG_OK;
}

int hyg_weighted_fdr_procedure(hyg_ctx* c, uint64_t n, const double* test_statistics, double fdr_threshold, const double* weights_false_positives,
                               const double* weights_false_negatives, uint64_t* n_selected, uint64_t* indices, double* Nk) {
  if (!c || !test_statistics || !weights_false_positives || !weights_false_negatives || !n_selected || !indices || !Nk || n == 0)
    return fail(c, HYG_ERR_ARG, "null argument / empty input");
  HYG_CUDA(c, cudaSetDevice(c->device));
  DmpScratch sc(c);
  double *d_t = nullptr, *d_fp = nullptr, *d_fn = nullptr, *d_rank = nullptr, *d_ex = nullptr, *d_sum = nullptr;
  unsigned long long *d_idx = nullptr, *d_key = nullptr, *d_k2 = nullptr, *d_p2 = nullptr;
  for (double** p : {&d_t, &d_fp, &d_fn, &d_rank, &d_ex, &d_sum}) HYG_FDR_CUDA(sc.get(p, n));
  for (unsigned long long** p : {&d_idx, &d_key, &d_k2, &d_p2}) HYG_FDR_CUDA(sc.get(p, n));
  HYG_FDR_CUDA(cudaMemcpyAsync(d_t, test_statistics, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  HYG_FDR_CUDA(cudaMemcpyAsync(d_fp, weights_false_positives, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  HYG_FDR_CUDA(cudaMemcpyAsync(d_fn, weights_false_negatives, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  hyg::dmp_weighted_rank_kernel<<<c->num_sms * 4, 256, 0, c->stream>>>(d_t, d_fp, d_fn, fdr_threshold, d_rank, d_ex, d_idx, n);
  // np.argsort(ranking) (:15); ties keep the smaller index first (numpy's default sort leaves tie order unspecified)
  hyg::dmp_make_keys_kernel<<<c->num_sms * 4, 256, 0, c->stream>>>(d_rank, d_key, n);
  HYG_FDR_CUDA(dmp_radix_sort(c, sc, d_key, d_k2, d_idx, d_p2, n));
  hyg::dmp_gather_kernel<<<c->num_sms * 4, 256, 0, c->stream>>>(d_ex, d_idx, d_rank, n);     // ranked excessive error rates (:17), d_rank reused
  HYG_FDR_CUDA(dmp_inclusive_scan(c, sc, d_rank, d_sum, n));                                 // Nsums (:18)
  uint64_t s = 0;
  HYG_FDR_CUDA(dmp_count_le(c, sc, d_sum, n, 0.0, &s));                                      // :19
  *n_selected = s;
  if (s > 0) HYG_FDR_CUDA(cudaMemcpyAsync(indices, d_idx, s * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  // Nsums[s-1]; for s = 0 Python's index -1 is the last element (:20)
  HYG_FDR_CUDA(cudaMemcpyAsync(Nk, d_sum + (s > 0 ? s - 1 : n - 1), sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  HYG_FDR_CUDA(cudaStreamSynchronize(c->stream));
  return HYG_OK;
}
#undef HYG_FDR_CUDA
#undef HYG_DMP_CUDA

int hyg_sg_sample_theta_prior(uint32_t dim, uint64_t seed, double* theta) {
  if (!theta) return HYG_ERR_ARG;
  // theta ~ N(0, I) (singleGroup.h:479-483), Box-Muller on Philox uniforms (stream tag 0xFFFFFFFF)
  for (uint32_t i = 0; i < dim; i += 2) {
    double u1 = hyg::philox_uniform(seed, 0xFFFFFFFFu, i), u2 = hyg::philox_uniform(seed, 0xFFFFFFFFu, i + 1);
    if (u1 < 1e-300) u1 = 1e-300;
    const double r = std::sqrt(-2.0 * std::log(u1)), a = 6.283185307179586476925 * u2;
    theta[i] = r * std::cos(a);
    if (i + 1 < dim) theta[i + 1] = r * std::sin(a);
  }
  return HYG_OK;
}

}  // extern "C"
