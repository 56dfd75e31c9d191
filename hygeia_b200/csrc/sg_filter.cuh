// hygeia_b200/csrc/sg_filter.cuh -- K2: the single-group recursion, one CTA per chain.
//
// What it computes (reference = /root/reference/src/single_group/src/cpp):
//   * discrete particle filter over (sojourn d, regime r), <= 256 support points
//       Smc::initialise / iterate                      algorithms/Smc.h:114-286
//       resampleCp + resample::optimalFiniteState      algorithms/Smc.h:406-450, misc/resample.h:289-409
//       systematic resampling, ONE uniform per site    misc/resample.h:85-127
//       sampleParticlesCp / computeWeightsCp           algorithms/Smc.h:504-574
//       selfNormaliseWeights (running log Z_t)         algorithms/Smc.h:576-579
//   * backward kernels of the R new-segment particles  algorithms/Smc.h:288-326
//   * forward-only adaptive fixed-lag smoothing        algorithms/OnlineMarginalSmoothing.h:40-63,119-255
//   * the outer loop over sites                        algorithms/OnlineCombinedInference.h:48-118
//
// How (B200-first, not a translation):
//   * the emission term logObs[t][r] is read from the T x R table K1 produced (48 B per site) -- the
//     reference re-evaluates it 1744 times per site;
//   * transition terms come from a host-built table {c_new(d,r), log(1-rho(d,r))} that each particle
//     carries one step ahead (entry for d+1 is gathered from the ancestor, entry for d+2 is loaded now and
//     consumed a whole step later), so no L2 latency sits on the per-site critical path;
//   * the R x N_prev new-segment log-sum-exps and backward kernels collapse to R class sums
//     E[r'] = sum_{n in class r'} W_n c_new(d_n, r') because logTrans((1,r) <- (d,r')) = log c_new(d,r') + log P[r'][r]
//     factorises: 6 block reductions instead of 3000 exp() per site (exact log-domain fallback when a class
//     underflows);
//   * sort = register/shuffle bitonic network (strides < 32) + 6 shared-memory exchange stages;
//     prefix sums, the K fixed point and systematic resampling are warp scans/ballots;
//   * the uniform of site t is Philox(seed, chain, t) or an injected per-site array.
// All arithmetic fp64.  Rounding differs from the reference at the 1e-16 level (tree sums vs sequential
// sums); decisions (K, ancestors) are discontinuous in the weights, so parity is stated on log Z_t and the
// posteriors (1e-6 relative) and on the argmax regime calls, and checked step by step against the oracle.
#ifndef HYG_SG_FILTER_CUH
#define HYG_SG_FILTER_CUH

#include "hyg_common.cuh"
#include "hyg_dev_structs.h"
#include "sg_param.cuh"

namespace hyg {

struct SgSmem {
  // previous particle system, storage order (thread n owns slot n)
  double W[HYG_NPMAX];
  double lw[HYG_NPMAX];
  double2 cur[HYG_NPMAX];  // {c_new(d,r), log(1-rho(d,r))}
  double2 nxt[HYG_NPMAX];  // same for d+1
  uint32_t d[HYG_NPMAX];
  unsigned char r[HYG_NPMAX];
  // resampling scratch
  unsigned long long key[6][HYG_NPMAX];   // one exchange buffer per cross-warp sort stage (no reuse inside a site)
  double Q[HYG_NPMAX + 1];
  unsigned short idx[HYG_NPMAX];
  unsigned short anc[HYG_NPMAX];
  int iscan[2][HYG_NW];
  BlockScratch sc;
  double lo[2][HYG_RMAX];
  double part[2][HYG_NW][8];   // per-warp partials of the 8-wide transposed reductions (double-buffered)
  double partG[2][HYG_NW][8];  // parameter mode: partials of Eg[r'] = sum e_n dlogrho_n
  unsigned vmask[2][HYG_NW];
  int slow[HYG_RMAX];
  unsigned slowmask;           // bit r: regime r takes the log-domain path at this site
  double lomax[2];             // max_r logObs(t, r), [t & 1] (published by the service warp with the prefetched row)
  double new_lw[HYG_RMAX];     // log-weight of the new-segment particle (1, r)
  double new_invE[HYG_RMAX];   // 1 / sumE[r]
  double u;                    // resampling uniform of the current site
  double res_lw;               // log-weight given to resampled particles: lsum_prev - log C
  double lsum[2];              // running log Z_t, [t & 1]
};

__device__ __forceinline__ bool hyg_isfinite(double x) {
  const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(x));
  return ((b >> 52) & 0x7ffull) != 0x7ffull;
}

// #{ j in [0,L) : (j + u) / L <= c } = floor(c L - u) + 1 clamped to [0, L]  (T_j = (j+u)/L, resample.h:95); x = c L - u.
// The closed form and the reference's compare-by-division agree unless |x - integer| is below rounding (~1e-13).
__device__ __forceinline__ int sys_count_x(double x, int L) {
  return (x < 0.0) ? 0 : ((x >= static_cast<double>(L)) ? L : static_cast<int>(x) + 1);
}

#ifndef HYG_SORT_RANKMERGE
#define HYG_SORT_RANKMERGE 0
#endif
#if HYG_SORT_RANKMERGE
#define HYG_SORT_BARRIERS 2
// Descending sort of one unique 64-bit key per worker thread (256 keys).  Each warp sorts its 32 keys with a
// register/shuffle bitonic network (15 stages, no barrier); every key then finds its rank in the other seven sorted
// runs by binary search in shared memory (7 independent 5-step searches) and is scattered to its final position:
// 2 barriers instead of the 6 exchange barriers + 36 dependent stages of a full block-wide bitonic sort.
// Returns the key of sorted position threadIdx.x.  Worker warps only.
__device__ __forceinline__ unsigned long long block_sort_desc(unsigned long long key, SgSmem& s) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const unsigned long long other = __shfl_xor_sync(HYG_FULL, key, j);
      const bool desc_block = ((lane & k) == 0);
      const bool lower = ((lane & j) == 0);
      const bool take_max = (lower == desc_block);
      const bool other_gt = other > key;
      key = (take_max == other_gt) ? other : key;
    }
  }
  s.key[0][tid] = key;   // eight descending runs of 32
  __syncthreads();
  int rank = lane;       // position inside the own run
#pragma unroll
  for (int w = 0; w < HYG_WORKER_WARPS; w++) {
    if (w == warp) continue;
    const unsigned long long* run = s.key[0] + 32 * w;
    // number of keys in `run` (descending) that are greater than `key`
    int lo = 0;
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) lo += (run[lo + step - 1] > key) ? step : 0;
    lo += (run[lo] > key) ? 1 : 0;   // lo <= 31 here
    rank += lo;
  }
  s.key[1][rank] = key;
  __syncthreads();
  return s.key[1][tid];
}

#else
#define HYG_SORT_BARRIERS 1
// Descending bitonic sort of one 64-bit key per worker thread (256 keys): strides < 32 by warp shuffles, strides
// 32/64/128 through shared memory (6 exchange stages, each with its own buffer).  The first exchange is a full block barrier
// (it also publishes the class sums to every warp, service warp included); in the other five a warp only needs its partner
// warp's keys, so they are 64-thread named barriers: id = 1 + 4 log2(stride / 32) + pair index, always the same two warps per id.
__device__ __forceinline__ unsigned long long block_sort_desc(unsigned long long key, SgSmem& s) {
  const int tid = threadIdx.x, warp = tid >> 5;
  int kbuf = 0;
#pragma unroll
  for (int k = 2; k <= HYG_NPMAX; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      unsigned long long other;
      if (j < 32) {
        other = __shfl_xor_sync(HYG_FULL, key, j);
      } else {
        s.key[kbuf][tid] = key;
        if (kbuf == 0) {
          __syncthreads();
        } else {
          // pair index: the warp number with the bit of the partner stride removed
          const int jw = j >> 5;   // 1, 2, 4
          const int pair = (warp & (jw - 1)) | ((warp & ~(2 * jw - 1)) >> 1);
          named_barrier(1 + 4 * (jw == 1 ? 0 : (jw == 2 ? 1 : 2)) + pair, 64);
        }
        other = s.key[kbuf][tid ^ j];
        kbuf++;
      }
      const bool desc_block = ((tid & k) == 0);
      const bool lower = ((tid & j) == 0);
      const bool take_max = (lower == desc_block);
      const bool other_gt = other > key;
      key = (take_max == other_gt) ? other : key;
    }
  }
  return key;
}
#endif

struct SgChainState {
  // per-thread particle (slot = threadIdx.x)
  double lw, W;
  double2 cur, nxt;
  double gcur, gnxt;   // parameter mode: d log rho / d theta_omega for d and d+1
  uint32_t d;
  int r;
};

template <int RT> __device__ __forceinline__ double pick(const double (&v)[RT], int i) {
  double o = 0.0;
#pragma unroll
  for (int q = 0; q < RT; q++) o = (i == q) ? v[q] : o;
  return o;
}

// Transposed warp reduction of eight values per lane: 9 shuffle steps instead of 40.  On return every lane holds the
// warp total of value index (lane >> 2) & 7.
__device__ __forceinline__ double warp_reduce8(const double (&v)[8]) {
  const int lane = threadIdx.x & 31;
  double w4[4], w2[2], w1;
  {
    const bool hi = (lane & 16) != 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const double send = hi ? v[i] : v[i + 4];
      const double keep = hi ? v[i + 4] : v[i];
      w4[i] = keep + __shfl_xor_sync(HYG_FULL, send, 16);
    }
  }
  {
    const bool hi = (lane & 8) != 0;
#pragma unroll
    for (int i = 0; i < 2; i++) {
      const double send = hi ? w4[i] : w4[i + 2];
      const double keep = hi ? w4[i + 2] : w4[i];
      w2[i] = keep + __shfl_xor_sync(HYG_FULL, send, 8);
    }
  }
  {
    const bool hi = (lane & 4) != 0;
    const double send = hi ? w2[0] : w2[1];
    const double keep = hi ? w2[1] : w2[0];
    w1 = keep + __shfl_xor_sync(HYG_FULL, send, 4);
  }
  w1 += __shfl_xor_sync(HYG_FULL, w1, 2);
  w1 += __shfl_xor_sync(HYG_FULL, w1, 1);
  return w1;
}
// The same reduction for the one-hot input v[i] = (i == r ? e : 0), i < 7, v[7] = x7 (r < 7): the first exchange stage needs
// only the class index relative to the half the lane sends / keeps, so the eight values are never materialised.  Bit-identical
// to warp_reduce8 on that input (same additions in the same order).
__device__ __forceinline__ double warp_reduce_onehot(double e, int r, double x7) {
  const int lane = threadIdx.x & 31;
  double w4[4], w2[2], w1;
  {
    const bool hi = (lane & 16) != 0;
    const int rs = hi ? r : r - 4;   // index (0..3) of the non-zero among the four values this lane sends
    const int rk = hi ? r - 4 : r;   // ... and among the four it keeps
#pragma unroll
    for (int i = 0; i < 4; i++) {
      double send = (rs == i) ? e : 0.0;
      double keep = (rk == i) ? e : 0.0;
      if (i == 3) { send = hi ? send : x7; keep = hi ? x7 : keep; }   // v[7] travels with the upper four
      w4[i] = keep + __shfl_xor_sync(HYG_FULL, send, 16);
    }
  }
  {
    const bool hi = (lane & 8) != 0;
#pragma unroll
    for (int i = 0; i < 2; i++) {
      const double send = hi ? w4[i] : w4[i + 2];
      const double keep = hi ? w4[i + 2] : w4[i];
      w2[i] = keep + __shfl_xor_sync(HYG_FULL, send, 8);
    }
  }
  {
    const bool hi = (lane & 4) != 0;
    const double send = hi ? w2[0] : w2[1];
    const double keep = hi ? w2[1] : w2[0];
    w1 = keep + __shfl_xor_sync(HYG_FULL, send, 4);
  }
  w1 += __shfl_xor_sync(HYG_FULL, w1, 2);
  w1 += __shfl_xor_sync(HYG_FULL, w1, 1);
  return w1;
}
// Lanes 0,4,..,28 publish the warp totals; any warp then folds the eight rows with two loads and two shuffle steps.
__device__ __forceinline__ void publish8(double (*part)[8], double wtot) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if ((lane & 3) == 0) part[warp][lane >> 2] = wtot;
}
// returns in every lane the block total of value index lane & 7
__device__ __forceinline__ double combine8(const double (*part)[8]) {
  const int lane = threadIdx.x & 31;
  double a = part[lane >> 3][lane & 7] + part[(lane >> 3) + 4][lane & 7];
  a += __shfl_xor_sync(HYG_FULL, a, 8);
  a += __shfl_xor_sync(HYG_FULL, a, 16);
  return a;
}

#define HYG_LINEAR_FLOOR 1e-250

// Service-warp job: new-segment particles (1, r), r = lane < R.  sumE[r] = sum_{r' != r} P[r'][r] E[r'] is the linear-domain
// mass flowing into regime r (relative to exp(lsum_prev)); its log-weight is lsum_prev + logObs_r + log(sumE[r])
// (computeWeightsCp, Smc.h:562-573, after factorising logTrans((1,r) <- (d,r')) = log c_new(d,r') + log P[r'][r]).
// Returns sumE[lane]; the log is taken later, in one evaluation shared with log C.
template <int R>
__device__ __forceinline__ double sg_service_new_segments(const SgModelDev& mdl, SgSmem& s, double totA, int pA) {
  const int lane = threadIdx.x & 31;
  unsigned vm = 0;
#pragma unroll
  for (int w = 0; w < HYG_WORKER_WARPS; w++) vm |= s.vmask[pA][w];
  double a = 0.0;
  bool could = false;
#pragma unroll
  for (int rp = 0; rp < R; rp++) {
    const double Erp = __shfl_sync(HYG_FULL, totA, rp);
    const double Pv = (lane < R) ? mdl.P[rp][lane] : 0.0;   // zero diagonal
    a += Pv * Erp;
    could = could || (((vm >> rp) & 1u) && Pv > 0.0);
  }
  if (lane < R) {
    // Below HYG_LINEAR_FLOOR the linear-domain sum is made of subnormal terms (or 1/a overflows): such regimes take the
    // exact log-domain path, like the ones whose sum underflowed to zero.
    const bool lin = a > HYG_LINEAR_FLOOR;
    s.new_invE[lane] = lin ? 1.0 / a : 0.0;
    s.slow[lane] = (!lin && could) ? 1 : 0;
  }
  const unsigned smask = __ballot_sync(HYG_FULL, (lane < R) && !(a > HYG_LINEAR_FLOOR) && could);
  if (lane == 0) s.slowmask = smask;
  return a;
}

// plain load for tables that this kernel rewrites (parameter mode), read-only path otherwise
template <bool PE, class T> __device__ __forceinline__ T tab_load(const T* p) {
  if (PE) return *p;
  return __ldg(p);
}

// Lag-set update of one site: psi of every pending site is propagated to the new particle system and sites whose R filtered
// variances dropped below epsilon are emitted (OnlineMarginalSmoothing.h:148-255).  Returns the number of sites still pending.
template <int R> struct SgLagArgs {
  const double* pp;   // psi written at step t-1
  double* pc;         // psi of step t
  int* pend_t;
  int n_pend, N_prev, M, N_curr, anc, pr;
  double e_prev, cW, my_invE;
  unsigned slowmask;
  double bk_slow[R];
  unsigned long long t, T, own_lo, own_hi, t_off;
  bool last_seg, worker;
  double epsilon;
};

template <int R>
__device__ __noinline__ int sg_lag_update(const SgLagArgs<R>& a, const SgModelDev& mdl, const SgChainDev& ch, SgSmem& s, int& flip, int& n_halo_forced) {
  const int tid = threadIdx.x;
  const int N_prev = a.N_prev, M = a.M, N_curr = a.N_curr, anc = a.anc;
  int kept = 0;
  for (int i = 0; i < a.n_pend; i++) {
    const double* src = a.pp + static_cast<size_t>(i) * R * HYG_NPMAX;
    double own[R], val[R];
#pragma unroll
    for (int q = 0; q < R; q++) {
      own[q] = (tid < N_prev) ? src[q * HYG_NPMAX + tid] : 0.0;
      val[q] = (tid < M) ? src[q * HYG_NPMAX + anc] : 0.0;
    }
    // class sums G[q][r'] = sum_{n in class r'} e_n psi_q[n]; new particle (1,r): sum_{r'} P[r'][r] G[q][r'] / sumE[r]
#pragma unroll
    for (int q = 0; q < R; q++) {
      double g[R];
#pragma unroll
      for (int rp = 0; rp < R; rp++) g[rp] = (a.pr == rp) ? a.e_prev * own[q] : 0.0;
      block_sum<R>(g, s.sc, flip);
      if (tid >= M && tid < N_curr) {
        const int r = tid - M;
        double acc = 0.0;
#pragma unroll
        for (int rp = 0; rp < R; rp++) acc += (rp != r) ? mdl.P[rp][r] * g[rp] : 0.0;
        val[q] = acc * a.my_invE;
      }
    }
    if (a.slowmask) {
#pragma unroll
      for (int r = 0; r < R; r++) {
        if (!((a.slowmask >> r) & 1u)) continue;
        double g[R];
#pragma unroll
        for (int q = 0; q < R; q++) g[q] = a.bk_slow[r] * own[q];
        block_sum<R>(g, s.sc, flip);
        if (tid == M + r) {
#pragma unroll
          for (int q = 0; q < R; q++) val[q] = g[q];
        }
      }
    }
    // storeEstimates (OnlineMarginalSmoothing.h:197-255): emit when all R filtered variances < epsilon
    double mv[2 * R];
#pragma unroll
    for (int q = 0; q < R; q++) { mv[q] = a.cW * val[q]; mv[R + q] = a.cW * val[q] * val[q]; }
    block_sum<2 * R>(mv, s.sc, flip);
    bool settled = true;
#pragma unroll
    for (int q = 0; q < R; q++) {
      const double var = mv[R + q] - mv[q] * mv[q];  // sum W (x-m)^2 with sum W = 1
      if (!(var < a.epsilon)) settled = false;
    }
    const bool emit = settled || (a.t == a.T - 1);
    const int ts = a.pend_t[i];
    if (emit) {
      const bool own_s = (static_cast<unsigned long long>(ts) >= a.own_lo) && (static_cast<unsigned long long>(ts) < a.own_hi);
      if (own_s) {
        // whole row (position, p_1..p_R) in one store instruction: 56 contiguous bytes, also when the row goes to mapped host memory
        if (tid <= R && ch.probs) {
          const double outv = (tid == 0) ? (ch.pos ? static_cast<double>(ch.pos[ts]) : static_cast<double>(static_cast<unsigned long long>(ts) + a.t_off))
                                         : pick<2 * R>(mv, tid - 1);
          ch.probs[static_cast<size_t>(ts) * (R + 1) + tid] = outv;
        }
        if (tid == 0 && ch.finalised_at) ch.finalised_at[ts] = static_cast<int>(a.t + a.t_off);
        if (!settled && !a.last_seg) n_halo_forced++;   // the segment's right halo ended before this site settled
      }
    } else {
      double* dst = a.pc + static_cast<size_t>(kept) * R * HYG_NPMAX;
      if (a.worker) {
#pragma unroll
        for (int q = 0; q < R; q++) dst[q * HYG_NPMAX + tid] = val[q];
      }
      __syncthreads();  // pend_t[i] has been read by every thread before slot `kept` (<= i) is overwritten
      if (tid == 0) a.pend_t[kept] = ts;
      kept++;
    }
  }
  return kept;
}

template <int RT, bool PE>
__device__ void sg_filter_chain(SgModelDev& mdl, const SgChainDev& ch, const SgRunDev& run, double* psi_ws, SgSmem& s, SgPeSmem<RT>* pe) {
  static_assert(RT <= 7, "class sums share an 8-wide reduction with the finite-weight count");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // Warps 0..7 own the particles (one slot per thread).  Warp 8 is a SERVICE warp: it owns no particle, follows the same
  // barrier sequence, and evaluates every scalar exp/log of the step (regime factors, new-segment weights, log C, log Z_t,
  // the Philox draw, the emission-row prefetch) while the workers sort and scan -- so no transcendental latency chain
  // sits on the workers' critical path.
  const bool worker = warp < HYG_WORKER_WARPS;
  const bool service = !worker;
  constexpr int R = RT;
  const int Nmax = mdl.n_particles;
  uint32_t dcap = mdl.dcap;
  const unsigned int T = static_cast<unsigned int>(ch.T);   // sites of this unit (< 2^32): 32-bit counters in the hot loop
  const int lcap = run.lcap;
  int flip = 0, ibuf = 0, pbuf = 0;
  constexpr int D = R * R;
  // parameter mode: per-CTA table workspace (tables are rebuilt on the device whenever theta moves)
  double2* pe_tab = nullptr; double* pe_tabg = nullptr; double* pe_wh = nullptr; double* pe_wg = nullptr;
  if (PE) {
    double* base = run.pe_ws + static_cast<size_t>(blockIdx.x) * run.pe_stride;
    pe_tab = reinterpret_cast<double2*>(base);
    pe_tabg = base + 2ull * R * run.pe_dcap;
    pe_wh = pe_tabg + static_cast<size_t>(R) * run.pe_dcap;
    pe_wg = pe_wh + static_cast<size_t>(R) * run.pe_dcap;
    if (tid < D) {
      pe->theta[tid] = ch.theta0[tid];
      pe->adam_m[tid] = 0.0; pe->adam_v[tid] = 0.0; pe->grad_cur[tid] = 0.0; pe->grad_prev[tid] = 0.0;
    }
    if (tid < R) pe->kappa[tid] = run.kappa[tid];
    if (tid == 0) { pe->iter = 0; mdl.tab = pe_tab; mdl.tabg = pe_tabg; }
    for (int i = tid; i < 2 * D * HYG_PHI_PITCH; i += HYG_NT) (&pe->phi[0][0][0])[i] = 0.0;
    __syncthreads();
    pe_set_theta<R>(mdl, *pe);
    uint32_t dn = run.n_steps_without_update + 12;
    dn = dn > run.pe_dcap ? run.pe_dcap : dn;
    pe_rebuild_tables<R>(mdl, *pe, pe_tab, pe_tabg, pe_wh, pe_wg, run.pe_dcap, dn);
    dcap = run.pe_dcap;   // row pitch of the workspace tables; valid entries: mdl.dcap
  }
  const SgModelDev& cm = mdl;

  // psi workspace (global, L2-resident): [2][lcap][R][256] doubles followed by lcap ints of site indices
  double* psi[2] = {psi_ws, psi_ws + static_cast<size_t>(lcap) * R * HYG_NPMAX};
  int* pend_t = reinterpret_cast<int*>(psi_ws + 2 * static_cast<size_t>(lcap) * R * HYG_NPMAX);
  int n_pend = 0, n_forced = 0, max_pend = 0;
  // segmented execution: local sites [own_lo, own_hi) are written, the rest is warm-up / run-out (hyg_dev_structs.h)
  const unsigned int own_lo = static_cast<unsigned int>(ch.own_lo), own_hi = static_cast<unsigned int>(ch.own_hi);
  const unsigned long long t_off = ch.t_off;
  const bool last_seg = ch.last_segment != 0;
  int n_halo_forced = 0;
  int n_steps = 0;
  double lz_base = 0.0;   // log Z (local) of site own_lo - 1: owned rows of logz are written relative to it

  SgChainState p;
  p.lw = -HYG_INF; p.W = 0.0; p.cur = make_double2(0.0, 0.0); p.nxt = p.cur; p.gcur = 0.0; p.gnxt = 0.0; p.d = 0; p.r = 0;

  // ---- t = 0 : Smc::initialise (Smc.h:114-188) ----
  if (tid < R) s.lo[0][tid] = __ldg(ch.logobs + tid);
  if (T > 1 && tid < R) s.lo[1][tid] = __ldg(ch.logobs + R + tid);
  __syncthreads();
  if (tid < 2 && static_cast<unsigned int>(tid) < T) {
    double m = s.lo[tid][0];
#pragma unroll
    for (int r = 1; r < R; r++) m = s.lo[tid][r] > m ? s.lo[tid][r] : m;
    s.lomax[tid] = m;
  }
  int N = R;
  double pend_shift = 0.0, pend_S = 1.0;   // log Z of the last completed site = pend_shift + log(pend_S)
  {
    double lomax = s.lo[0][0];
#pragma unroll
    for (int r = 1; r < R; r++) lomax = s.lo[0][r] > lomax ? s.lo[0][r] : lomax;
    const double lsum_prev = -log(static_cast<double>(R));  // evaluateLogInitialDensity, singleGroup.h:559-566
    if (tid < R) {
      p.d = 1; p.r = tid;
      p.lw = lsum_prev + s.lo[0][tid];
      p.cur = tab_load<PE>(mdl.tab + static_cast<size_t>(tid) * dcap + 0);
      p.nxt = tab_load<PE>(mdl.tab + static_cast<size_t>(tid) * dcap + 1);
      if (PE) { p.gcur = mdl.tabg[static_cast<size_t>(tid) * dcap + 0]; p.gnxt = mdl.tabg[static_cast<size_t>(tid) * dcap + 1]; }
    }
    const double c = lsum_prev + lomax;
    double e[1] = {(tid < N) ? exp(p.lw - c) : 0.0};
    const double mine = e[0];
    block_sum<1>(e, s.sc, flip);
    pend_shift = c; pend_S = e[0];
    if (tid == 0) s.lsum[0] = c + log(e[0]);
    p.W = mine / e[0];
  }

  double pos_nxt = (tid == 0) ? (ch.pos ? static_cast<double>(__ldg(ch.pos)) : static_cast<double>(t_off)) : 0.0;
  for (unsigned int t = 0; t < T; t++) {
    const double* lo = s.lo[t & 1];
    // emission row of site t+2: issued now by the service warp, stored at the end of this step
    double lo_pref = 0.0;
    if (service) {   // warp-uniform: the eight worker warps skip the address arithmetic
      if (lane < R && t + 2 < T) lo_pref = __ldg(ch.logobs + static_cast<size_t>(t + 2) * R + lane);
    }
    // genomic position of the next site, loaded a step ahead by the thread that writes column 0 of the posterior rows
    const double pos_cur = pos_nxt;
    if (warp == 0 && tid == 0 && t + 1 < T) pos_nxt = ch.pos ? static_cast<double>(__ldg(ch.pos + t + 1)) : static_cast<double>(static_cast<unsigned long long>(t) + 1ull + t_off);
    int k_kept = -1;
    bool drew = false;
    bool emit_now = false;       // current site finalised at this step
    const bool own_t = (t >= own_lo) && (t < own_hi);
    double cw_lane = 0.0;        // regime mass of index lane & 7 (current site)

    if (t > 0) {
      // =========================== Smc::iterate (Smc.h:190-286) ===========================
      const int N_prev = N;
      const int N_curr = (N_prev + R > Nmax) ? Nmax : N_prev + R;
      const int M = N_curr - R;
      const bool capped = (N_curr < N_prev + R);
      int anc = tid;
      const double lomax = s.lomax[t & 1];

      if (service) {
        // the uniform of this site, and log Z_{t-1}: the log of last step's normaliser is evaluated now, off the
        // workers' critical path (they read s.lsum only after the sort's barriers)
        if (lane == 8) s.u = ch.unif ? __ldg(ch.unif + t) : philox_uniform(ch.seed, ch.chain_id, t + t_off);
        if (lane == 0) s.lsum[(t + 1) & 1] = pend_shift + log(pend_S);
        __syncwarp();
      }

      // ---- class sums over the previous particles (replace the R x N_prev log-sum-exps of Smc.h:562-573) ----
      const double e_prev = (tid < N_prev) ? p.W * p.cur.x : 0.0;  // W_n * c_new(d_n, r_n)
      const bool finite_prev = (tid < N_prev) && hyg_isfinite(p.lw);
      const bool valid = finite_prev && (p.cur.x > 0.0);
      if (worker) {
        // class sums of e_prev; slot 7: F = #finite(logw_prev), Smc.h:413 (exact in fp64).  e_prev is 0 beyond N_prev.
        publish8(s.part[pbuf], warp_reduce_onehot(e_prev, p.r, finite_prev ? 1.0 : 0.0));
        const unsigned vm = __reduce_or_sync(HYG_FULL, valid ? (1u << p.r) : 0u);
        if (lane == 0) s.vmask[pbuf][warp] = vm;
        if (PE) {
          pe->eprev[tid] = e_prev;
          publish8(s.partG[pbuf], warp_reduce_onehot(e_prev * p.gcur, p.r, 0.0));
        }
      }
      const int pA = pbuf;
      pbuf ^= 1;

      // ---- ancestors: Smc::resampleCp (Smc.h:406-450) ----
      bool own_weight = true;   // child keeps its ancestor's own weight (top-K / keep-largest / growth)
      int sidx = tid;
      double totA;
      double sumE_lane = 0.0;   // service warp, lane r < R: new-segment mass of regime r
      if (capped) {
        if (worker) {
          // sort by log-weight, descending (ties by slot); W is a monotone map of logw
          // (dead slots get distinct tiny keys so that every key is unique)
          unsigned long long key = (tid < N_prev) ? order_key(p.lw, tid) : static_cast<unsigned long long>(255 - tid);
          key = block_sort_desc(key, s);   // s.part[pA] / s.vmask[pA] are visible after its first barrier
          sidx = 255 - static_cast<int>(key & 0xFFull);
          s.idx[tid] = static_cast<unsigned short>(sidx);
          totA = combine8(s.part[pA]);
        } else {
          __syncthreads();
          totA = combine8(s.part[pA]);
          sumE_lane = sg_service_new_segments<R>(mdl, s, totA, pA);
          if (PE) {
            const double totG = combine8(s.partG[pA]);
            if (lane < 8) { pe->Etot[lane] = totA; pe->Egtot[lane] = totG; }
          }
#pragma unroll
          for (int b = 1; b < HYG_SORT_BARRIERS; b++) __syncthreads();
        }
      } else {
        __syncthreads();
        totA = combine8(s.part[pA]);           // lane & 7 -> E[0..R-1], [7] = F
        if (service) {
          sumE_lane = sg_service_new_segments<R>(mdl, s, totA, pA);
          if (PE) {
            const double totG = combine8(s.partG[pA]);
            if (lane < 8) { pe->Etot[lane] = totA; pe->Egtot[lane] = totG; }
          }
          if (lane < R) s.new_lw[lane] = (sumE_lane > 0.0) ? s.lsum[(t + 1) & 1] + lo[lane] + log(sumE_lane) : -HYG_INF;
        }
        __syncthreads();                       // growth phase has no later barrier before the new-segment values are read
      }
      const int F = static_cast<int>(__shfl_sync(HYG_FULL, totA, 7) + 0.5);
      if (capped) {
        bool keep_largest = (F <= M);
        int K = 0;
        double Qk = 0.0;
        if (!keep_largest) {
          // ---- resample::optimalFiniteState (resample.h:289-409) ----
          const double qv = (tid < N_prev) ? s.W[sidx] : 0.0;
          double v = qv;  // Q[p] = sum_{j >= p} q_j
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const double tt = __shfl_down_sync(HYG_FULL, v, o);
            if (lane + o < 32) v += tt;
          }
          if (lane == 0) s.sc.d[flip][warp][0] = v;
          __syncthreads();
          double tail = 0.0;
#pragma unroll
          for (int w = HYG_NW - 1; w > 0; w--)
            if (w > warp) tail += s.sc.d[flip][w][0];   // warp-uniform predicate: the additions of the warps below are skipped
          flip ^= 1;
          const double Qp = v + tail;
          if (worker) s.Q[tid] = Qp;
          if (tid == 0) s.Q[HYG_NPMAX] = 0.0;
          // Fixed point for K (resample.h:333-342).  The reference iterates K <- K + #{i >= K : log q_i > -log C(K)},
          // log C(K) = log(M-K) - log Q[K], from K = 0.  Along that iteration the threshold Q[K]/(M-K) only decreases, so
          // it stops at the FIRST sorted position p whose own weight is not above its own threshold:
          //   K* = min{ p : !(q_p (M-p) > Q[p]) }   (one parallel pass instead of up to ~N sequential ones).
          const bool stop = (tid >= M) || (tid >= N_prev) || !(qv * static_cast<double>(M - tid) > Qp);
          const unsigned sb = __ballot_sync(HYG_FULL, stop);
          if (lane == 0) s.iscan[ibuf][warp] = sb ? (warp * 32 + __ffs(static_cast<int>(sb)) - 1) : HYG_NPMAX;
          __syncthreads();
          K = __reduce_min_sync(HYG_FULL, (lane < HYG_NW) ? s.iscan[ibuf][lane] : HYG_NPMAX);   // one load + one warp reduction
          ibuf ^= 1;
          if (K >= M) {
            keep_largest = true;  // log C not finite (resample.h:345,366)
          } else {
            Qk = s.Q[K];
            if (!(Qk > 0.0) || !hyg_isfinite(Qk)) keep_largest = true;
          }
        }
        if (keep_largest) {
          // keep the M largest by log-weight (Smc.h:432-441; resample.h:366-375)
          if (service && lane < R) s.new_lw[lane] = (sumE_lane > 0.0) ? s.lsum[(t + 1) & 1] + lo[lane] + log(sumE_lane) : -HYG_INF;
          __syncthreads();  // s.idx and s.new_lw visible
          anc = (tid < M) ? s.idx[tid] : tid;
          k_kept = -2;
        } else {
          const int L = M - K;
          if (service) {
            // ONE log evaluation: lanes 0..R-1 the new-segment particles, lane 7 the weight lsum_prev - log C of a
            // resampled particle (resample.h:361-364); consumed by the workers two barriers later
            const double arg = (lane == 7) ? Qk / static_cast<double>(L) : sumE_lane;
            const double lg = (arg > 0.0) ? log(arg) : -HYG_INF;
            const double lsp = s.lsum[(t + 1) & 1];
            if (lane < R) s.new_lw[lane] = (arg > 0.0) ? lsp + lo[lane] + lg : -HYG_INF;
            if (lane == 7) s.res_lw = lsp + lg;
          }
          k_kept = K;
          drew = true;
          // systematic resampling of L offspring among the sorted residual particles (resample.h:85-127,354-359).
          // C_p = #{ j < L : (j+u)/L <= cumulative residual weight up to p }; forced monotone, C_last = L, so the
          // offspring counts o_p = C_p - C_{p-1} are >= 0 and sum to L whatever the rounding of the suffix sums.
          const double u = s.u;
          int C = 0;
          if (tid >= K && tid < N_prev) C = (tid == N_prev - 1) ? L : sys_count_x((Qk - s.Q[tid + 1]) * (static_cast<double>(L) / Qk) - u, L);
          if (tid >= N_prev) C = L;
#pragma unroll
          for (int dlt = 1; dlt < 32; dlt <<= 1) {
            const int tt = __shfl_up_sync(HYG_FULL, C, dlt);
            if (lane >= dlt) C = tt > C ? tt : C;
          }
          if (lane == 31) s.iscan[ibuf][warp] = C;
          __syncthreads();
          int before = __reduce_max_sync(HYG_FULL, (lane < warp && lane < HYG_NW - 1) ? s.iscan[ibuf][lane] : 0);
          ibuf ^= 1;
          C = before > C ? before : C;
          int Cprev = __shfl_up_sync(HYG_FULL, C, 1);
          if (lane == 0) Cprev = before;
          if (tid <= K) Cprev = 0;
          if (tid < K) s.anc[tid] = static_cast<unsigned short>(sidx);
          if (tid >= K && tid < N_prev)
            for (int slot = K + Cprev; slot < K + C && slot < M; slot++) s.anc[slot] = static_cast<unsigned short>(sidx);
          __syncthreads();
          anc = (tid < M) ? s.anc[tid] : tid;
          own_weight = (tid < K);
        }
      }
      if (ch.ancestors && own_t && tid < Nmax - R)
        ch.ancestors[static_cast<unsigned long long>(t) * static_cast<unsigned long long>(Nmax - R) + tid] = (tid < M) ? static_cast<short>(anc) : static_cast<short>(-1);

      // ---- propose + weight: sampleParticlesCp / computeWeightsCp (Smc.h:504-574) ----
      const double lsum_prev = s.lsum[(t + 1) & 1];
      SgChainState c;
      c.lw = -HYG_INF; c.W = 0.0; c.cur = make_double2(0.0, 0.0); c.nxt = c.cur; c.gcur = 0.0; c.gnxt = 0.0; c.d = 0; c.r = 0;
      const uint32_t vcap = PE ? cm.dcap : dcap;   // valid table entries per regime (parameter mode: last rebuild)
      double grad_c = 0.0;   // parameter mode: d log f / d theta_omega(r) of the continuation (singleGroup.h:679-693)
      if (tid < M) {
        c.r = s.r[anc];
        c.d = s.d[anc] + 1;
        c.cur = s.nxt[anc];
        const double2 pc = s.cur[anc];
        const double lc = pc.y;                         // log(1 - rho(d_prev, r)) or -inf (singleGroup.h:597-605)
        c.lw = (own_weight ? s.lw[anc] : s.res_lw) + (lc + lo[c.r]);
        const uint32_t di = (c.d + 1 <= vcap) ? c.d : vcap - 1;  // 0-based index of d+1, clamped to the terminal entry
        c.nxt = tab_load<PE>(mdl.tab + static_cast<size_t>(c.r) * dcap + di);
        if (PE) {
          c.gcur = pe->gnxt[anc];
          c.gnxt = mdl.tabg[static_cast<size_t>(c.r) * dcap + di];
          const double rho = pc.x;                      // c_new = rho for d >= u (0 below u)
          grad_c = (lc > -HYG_INF && rho < 1.0) ? -pe->gcur[anc] * rho / (1.0 - rho) : 0.0;
        }
      } else if (tid < N_curr) {
        const int r = tid - M;
        c.r = r; c.d = 1;
        c.cur = tab_load<PE>(mdl.tab + static_cast<size_t>(r) * dcap + 0);
        c.nxt = tab_load<PE>(mdl.tab + static_cast<size_t>(r) * dcap + 1);
        if (PE) { c.gcur = mdl.tabg[static_cast<size_t>(r) * dcap + 0]; c.gnxt = mdl.tabg[static_cast<size_t>(r) * dcap + 1]; }
        c.lw = s.new_lw[r];
      }
      const double my_invE = (tid >= M && tid < N_curr) ? s.new_invE[tid - M] : 0.0;
      // exact log-domain path for regimes whose linear-domain sum underflowed (rare)
      const unsigned slowmask = s.slowmask;
      double bk_slow[R];
#pragma unroll
      for (int r = 0; r < R; r++) bk_slow[r] = 0.0;
      if (slowmask) {
#pragma unroll
        for (int r = 0; r < R; r++) {
          if (!((slowmask >> r) & 1u)) continue;
          const bool ok = valid && p.r != r && mdl.P[p.r][r] > 0.0;
          const double x = ok ? p.lw + log(p.cur.x) + mdl.logP[p.r][r] : -HYG_INF;
          const double mx = block_max(x, s.sc, flip);
          double ex[1] = {ok ? exp(x - mx) : 0.0};
          const double mine = ex[0];
          block_sum<1>(ex, s.sc, flip);
          if (tid == M + r) c.lw = lo[r] + (mx + log(ex[0]));
          bk_slow[r] = (ex[0] > 0.0) ? mine / ex[0] : 0.0;
        }
      }

      // ---- selfNormaliseWeights (Smc.h:576-579), fused with the regime masses of the new site ----
      {
        double shift = lsum_prev + lomax;   // upper bound of every logw instead of the exact max
        c.W = (tid < N_curr) ? exp(c.lw - shift) : 0.0;
        if (worker) publish8(s.part[pbuf], warp_reduce_onehot(c.W, c.r, 0.0));   // c.W is 0 beyond N_curr
        __syncthreads();
        double tot = combine8(s.part[pbuf]);   // lane & 7 -> class sum of the relative weights
        pbuf ^= 1;
        double S = tot;
        S += __shfl_xor_sync(HYG_FULL, S, 1);
        S += __shfl_xor_sync(HYG_FULL, S, 2);
        S += __shfl_xor_sync(HYG_FULL, S, 4);
        if (!(S > 0.0) || !hyg_isfinite(S)) {
          // the linear-domain weights underflowed against the bound: renormalise from the log-weights with the exact max
          shift = block_max((tid < N_curr) ? c.lw : -HYG_INF, s.sc, flip);
          c.W = (tid < N_curr && c.lw > -HYG_INF) ? exp(c.lw - shift) : 0.0;
          if (worker) publish8(s.part[pbuf], warp_reduce_onehot(c.W, c.r, 0.0));
          __syncthreads();
          tot = combine8(s.part[pbuf]);
          pbuf ^= 1;
          S = tot;
          S += __shfl_xor_sync(HYG_FULL, S, 1);
          S += __shfl_xor_sync(HYG_FULL, S, 2);
          S += __shfl_xor_sync(HYG_FULL, S, 4);
        }
        const double invS = 1.0 / S;
        c.W *= invS;
        cw_lane = tot * invS;
        pend_shift = shift; pend_S = S;   // log Z_t = shift + log S is evaluated by the service warp at the next step
      }

      // ---- fixed-lag smoothing: updatePsi (OnlineMarginalSmoothing.h:148-177) ----
      if (run.use_smoothing && n_pend > 0) {
        // out of line: the lag set is empty at most sites of informative data, and its 30 live doubles per thread would
        // otherwise set the register budget (and the spills) of the whole recursion
        SgLagArgs<R> la;
        la.pp = psi[(t + 1) & 1]; la.pc = psi[t & 1]; la.pend_t = pend_t; la.n_pend = n_pend;
        la.N_prev = N_prev; la.M = M; la.N_curr = N_curr; la.anc = anc; la.pr = p.r;
        la.e_prev = e_prev; la.cW = c.W; la.my_invE = my_invE; la.slowmask = slowmask;
#pragma unroll
        for (int r = 0; r < R; r++) la.bk_slow[r] = bk_slow[r];
        la.t = t; la.T = T; la.own_lo = own_lo; la.own_hi = own_hi; la.t_off = t_off; la.last_seg = last_seg; la.worker = worker;
        la.epsilon = run.epsilon;
        n_pend = sg_lag_update<R>(la, mdl, ch, s, flip, n_halo_forced);
      }
      // ---- K3: score recursion (OnlineParameterEstimation.h:135-158), see sg_param.cuh ----
      if (PE) {
        const int pb = (t + 1) & 1, cb = t & 1;   // phi buffers: [pb] was written at site t-1
        if (tid < 8 * D) {
          // class sums G[k][r'] = sum_{n in class r'} e_n phi_n[k]: thread = (component k, chunk of 32 previous particles)
          const int k = tid % D, chunk = tid / D;
          double acc[R];
#pragma unroll
          for (int q = 0; q < R; q++) acc[q] = 0.0;
          for (int i = 0; i < 32; i++) {
            const int n = chunk * 32 + i;
            const double ev = pe->eprev[n] * pe->phi[pb][k][n];
            const int rn = s.r[n];
#pragma unroll
            for (int q = 0; q < R; q++) acc[q] += (rn == q) ? ev : 0.0;
          }
#pragma unroll
          for (int q = 0; q < R; q++) pe->part[chunk][k][q] = acc[q];
        }
        __syncthreads();
        if (tid < D * R) {
          const int k = tid / R, q = tid % R;
          double a = 0.0;
#pragma unroll
          for (int chn = 0; chn < 8; chn++) a += pe->part[chn][k][q];
          pe->Gs[k][q] = a;
        }
        __syncthreads();
        if (tid < R * D) {
          // new particle (1, r): phi' = sum_n bk_r[n] (phi_n + grad_n), bk_r[n] = e_n P[r_n][r] / sumE[r];
          // grad_n = dlogrho(d_n, r_n) on the omega slot of r_n and (1[c == r] - P[r_n][c]) on the P block of r_n
          // (singleGroup.h:657-678)
          const int r = tid / D, k = tid % D;
          double a = 0.0;
          for (int rp = 0; rp < R; rp++) {
            if (rp == r) continue;
            double term = pe->Gs[k][rp];
            if (k == R * (R - 1) + rp) {
              term += pe->Egtot[rp];
            } else if (k >= rp * (R - 1) && k < (rp + 1) * (R - 1)) {
              const int j = k - rp * (R - 1);
              const int col = (j < rp) ? j : j + 1;
              term += pe->Etot[rp] * (((col == r) ? 1.0 : 0.0) - cm.P[rp][col]);
            }
            a += cm.P[rp][r] * term;
          }
          pe->phi[cb][k][M + r] = a * s.new_invE[r];
        }
        if (slowmask) {
          // regimes on the log-domain path: phi' = sum_n bk_r[n] (phi_n + grad_n) with the per-particle normalised kernel
          __syncthreads();
#pragma unroll
          for (int r = 0; r < R; r++) {
            if (!((slowmask >> r) & 1u)) continue;
            for (int k = 0; k < D; k++) {
              double v[1] = {0.0};
              if (bk_slow[r] > 0.0) {
                const int rn = p.r;
                double g = pe->phi[pb][k][tid];
                if (k == R * (R - 1) + rn) {
                  g += p.gcur;
                } else if (k >= rn * (R - 1) && k < (rn + 1) * (R - 1)) {
                  const int j = k - rn * (R - 1);
                  const int col = (j < rn) ? j : j + 1;
                  g += ((col == r) ? 1.0 : 0.0) - cm.P[rn][col];
                }
                v[0] = bk_slow[r] * g;
              }
              block_sum<1>(v, s.sc, flip);
              if (tid == 0) pe->phi[cb][k][M + r] = v[0];
            }
          }
        }
        if (tid < M) {
          // continuing particle: phi' = phi_anc + grad, grad touches only the omega slot of its regime
          const int ko = R * (R - 1) + c.r;
          for (int k = 0; k < D; k++) {
            double v = pe->phi[pb][k][anc];
            if (k == ko) v += grad_c;
            pe->phi[cb][k][tid] = v;
          }
        }
      }
      p = c;
      N = N_curr;
    } else {
      // t = 0: regime masses of the initial particle system
      double v8[8];
      if (worker) {
#pragma unroll
        for (int q = 0; q < 8; q++) v8[q] = (q < R && tid < N && p.r == q) ? p.W : 0.0;
        publish8(s.part[pbuf], warp_reduce8(v8));
      }
      __syncthreads();
      cw_lane = combine8(s.part[pbuf]);
      pbuf ^= 1;
    }

    // ---- initialisePsi + storeEstimates for the current site (OnlineMarginalSmoothing.h:119-146,197-255) ----
    if (run.use_smoothing) {
      // every lane holds the mass of regime (lane & 7); psi is the 0/1 regime indicator, so
      // Var = m (1-m)^2 + (sum W - m) m^2
      double sw = cw_lane;
      sw += __shfl_xor_sync(HYG_FULL, sw, 1);
      sw += __shfl_xor_sync(HYG_FULL, sw, 2);
      sw += __shfl_xor_sync(HYG_FULL, sw, 4);
      const double m = cw_lane;
      const double var = m * (1.0 - m) * (1.0 - m) + (sw - m) * m * m;
      const bool ok = ((lane & 7) >= R) || (var < run.epsilon);
      const bool settled = __all_sync(HYG_FULL, ok);
      emit_now = (t == T - 1) || settled;
      if (!emit_now && n_pend >= lcap) { emit_now = true; n_forced += own_t ? 1 : 0; }  // lag set full: emit the filtering estimate now (reported)
      if (emit_now) {
        if (own_t) {
          if (warp == 0 && ch.probs) {
            const double left = __shfl_up_sync(HYG_FULL, cw_lane, 1);   // lane j >= 1: mass of regime j-1
            if (lane <= R) ch.probs[static_cast<size_t>(t) * (R + 1) + lane] = (lane == 0) ? pos_cur : left;
          }
          if (tid == 0 && ch.finalised_at) ch.finalised_at[t] = static_cast<int>(t + t_off);
          if (t == T - 1 && !settled && !last_seg) n_halo_forced++;
        }
      } else {
        double* dst = psi[t & 1] + static_cast<size_t>(n_pend) * R * HYG_NPMAX;
        if (worker) {
#pragma unroll
          for (int q = 0; q < R; q++) dst[q * HYG_NPMAX + tid] = (tid < N && p.r == q) ? 1.0 : 0.0;
        }
        if (tid == 0) pend_t[n_pend] = static_cast<int>(t);
        n_pend++;
      }
      max_pend = n_pend > max_pend ? n_pend : max_pend;
    }

    // ---- publish the particle system for the next site (all gathers of this step precede the normaliser barrier) ----
    if (worker) {
      s.W[tid] = p.W; s.lw[tid] = p.lw; s.cur[tid] = p.cur; s.nxt[tid] = p.nxt; s.d[tid] = p.d; s.r[tid] = static_cast<unsigned char>(p.r);
      if (PE) { pe->gcur[tid] = p.gcur; pe->gnxt[tid] = p.gnxt; }
    }
    if (service && t + 2 < T) {
      if (lane < R) s.lo[t & 1][lane] = lo_pref;
      double mx = (lane < R) ? lo_pref : -HYG_INF;
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) { const double tt = __shfl_xor_sync(HYG_FULL, mx, o); mx = tt > mx ? tt : mx; }
      if (lane == 0) s.lomax[t & 1] = mx;
    }
    __syncthreads();
    // segmented execution: stop as soon as the owned range is stepped through and none of its sites is still pending
    // (the lag set is ordered by site, oldest first)
    bool exit_now = false;
    if (!PE && t + 1 >= own_hi && t + 1 < T) exit_now = (n_pend == 0) || (static_cast<unsigned int>(pend_t[0]) >= own_hi);
    const bool last_step = (t == T - 1) || exit_now;
    n_steps++;

    // ---- K3: parameter update every n_steps sites (OnlineParameterEstimation.h:51-61) ----
    if (PE) {
      if (t > 0 && (t % run.n_steps_without_update) == 0) {
        if (tid < D) {
          // g = sum_n W_n phi_n over the current particles (computeFilteredMean, Smc.h:340-349)
          double g = 0.0;
          for (int n = 0; n < N; n++) g = g + s.W[n] * pe->phi[t & 1][tid][n];
          pe->grad_prev[tid] = pe->grad_cur[tid];
          pe->grad_cur[tid] = g;
        }
        __syncthreads();
        pe_ascent<R>(*pe, run);
        __syncthreads();
        pe_set_theta<R>(mdl, *pe);
        // rebuild the sojourn tables up to the largest sojourn any particle can reach before the next rebuild
        const double dmax = block_max((tid < N) ? static_cast<double>(p.d) : 0.0, s.sc, flip);
        unsigned long long dn = static_cast<unsigned long long>(dmax) + run.n_steps_without_update + 4;
        dn = dn > run.pe_dcap ? run.pe_dcap : dn;
        pe_rebuild_tables<R>(mdl, *pe, pe_tab, pe_tabg, pe_wh, pe_wg, run.pe_dcap, static_cast<uint32_t>(dn));
        if (worker && tid < N) {
          const uint32_t vc = cm.dcap;
          const uint32_t i0 = (p.d <= vc ? p.d : vc) - 1, i1 = (p.d + 1 <= vc ? p.d + 1 : vc) - 1;
          p.cur = pe_tab[static_cast<size_t>(p.r) * dcap + i0];
          p.nxt = pe_tab[static_cast<size_t>(p.r) * dcap + i1];
          p.gcur = pe_tabg[static_cast<size_t>(p.r) * dcap + i0];
          p.gnxt = pe_tabg[static_cast<size_t>(p.r) * dcap + i1];
          s.cur[tid] = p.cur; s.nxt[tid] = p.nxt; pe->gcur[tid] = p.gcur; pe->gnxt[tid] = p.gnxt;
        }
        __syncthreads();
      }
      if (ch.theta_trace && tid < D) ch.theta_trace[static_cast<size_t>(t) * D + tid] = pe->theta[tid];
    }

    // ---- taps ----
    if (service && lane == 0) {
      if (t > 0) {
        const double lz_prev = s.lsum[(t + 1) & 1];   // log Z (local) of site t-1
        if (t == own_lo) lz_base = lz_prev;
        if (t - 1 >= own_lo && t - 1 < own_hi) {
          if (ch.logz) ch.logz[t - 1] = lz_prev - lz_base;
          if (t == own_hi && ch.seg_inc) *ch.seg_inc = lz_prev - lz_base;
        }
      }
      if (last_step && own_t) {
        const double lz = pend_shift + log(pend_S) - lz_base;
        if (ch.logz) ch.logz[t] = lz;
        if (t + 1 == own_hi && ch.seg_inc) *ch.seg_inc = lz;
      }
      if (own_t) {
        if (ch.k_kept) ch.k_kept[t] = k_kept;
        if (ch.drew) ch.drew[t] = drew ? 1 : 0;
        if (ch.n_pending) ch.n_pending[t] = n_pend;
        if (ch.n_curr) ch.n_curr[t] = N;
      }
    }
    if (exit_now) break;
  }
  if (tid == 0 && ch.status) {
    atomicAdd(ch.status + 0, n_forced);
    atomicMax(ch.status + 1, max_pend);
    atomicAdd(ch.status + 2, n_halo_forced);
    atomicAdd(ch.status + 3, n_steps);
  }
  __syncthreads();
}

// Persistent launch: CTAs pull chains (pre-sorted longest first by the host) from an atomic queue.
#ifdef HYG_EMU
static double hyg_pe_smem_storage[(sizeof(SgPeSmem<6>) + 7) / 8 + 8];
#define HYG_PE_SMEM reinterpret_cast<void*>(hyg_pe_smem_storage)
#else
extern __shared__ __align__(16) unsigned char hyg_pe_smem_dyn[];
#define HYG_PE_SMEM reinterpret_cast<void*>(hyg_pe_smem_dyn)
#endif

template <int RT, bool PE>
__device__ __forceinline__ void sg_filter_entry(const SgModelDev* mdl, const SgChainDev* chains, SgRunDev run) {
  __shared__ SgSmem s;
  __shared__ int s_next;
  __shared__ SgModelDev s_mdl;
  __shared__ SgChainDev s_ch;
  SgPeSmem<RT>* pe = PE ? reinterpret_cast<SgPeSmem<RT>*>(HYG_PE_SMEM) : nullptr;
  double* psi_ws = run.psi_ws + static_cast<size_t>(blockIdx.x) * run.psi_stride;
  for (;;) {
    if (threadIdx.x == 0) { s_next = static_cast<int>(atomicAdd(run.queue, 1u)); s_mdl = *mdl; }
    __syncthreads();
    const int c = s_next;
    __syncthreads();
    if (c >= run.n_chains) break;
    // the descriptor lives in shared memory: its dozen pointers are re-read where they are used (after a barrier the compiler
    // must reload them) instead of occupying two registers each for the whole recursion
    if (threadIdx.x == 0) s_ch = chains[c];
    __syncthreads();
    sg_filter_chain<RT, PE>(s_mdl, s_ch, run, psi_ws, s, pe);
  }
}

#ifndef HYG_K2_MIN_CTAS
#define HYG_K2_MIN_CTAS 2   // resident CTAs per SM the register budget is held to (K2 is latency-bound: a second chain fills idle issue slots)
#endif
#ifndef HYG_EMU
template <int RT, bool PE>
__global__ void __launch_bounds__(HYG_NT, PE ? 1 : HYG_K2_MIN_CTAS) sg_filter_kernel(const SgModelDev* mdl, const SgChainDev* chains, SgRunDev run) {
  sg_filter_entry<RT, PE>(mdl, chains, run);
}
// The same recursion compiled for ONE CTA per SM (168 registers, no spills): used when there are no more units than SMs
// (whole-chain execution of a few chains, the C ABI's and the CLI's default), where per-site latency is all that counts.
template <int RT>
__global__ void __launch_bounds__(HYG_NT, 1) sg_filter_kernel_sparse(const SgModelDev* mdl, const SgChainDev* chains, SgRunDev run) {
  sg_filter_entry<RT, false>(mdl, chains, run);
}
#endif

}  // namespace hyg
#endif
