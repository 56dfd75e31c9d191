// hygeia_b200/csrc/sg_filter.cuh -- K2: the single-group recursion, one CTA per chain (or chain segment).
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
//   * STATIONARY SLOTS.  The reference rebuilds its particle arrays in sorted order at every site (resample.h:347-364).
//     Nothing in the model depends on where a particle is stored, and at a typical site only R of the ~250 particles die
//     (the K of optimal resampling averages ~235 of M = 244).  So a particle stays in the thread (slot) it was born in, with
//     its state in registers; a site only rewrites the <= R slots whose particle died with the R new-segment particles.
//     No ancestor gather, no per-site publication of the particle system, and the lag-set / score vectors stay in place.
//   * The sort that optimal resampling needs runs only where the reference resamples optimally (at ~40 % of the sites of
//     32-sample data the weights of >= R particles have underflowed to -inf and "keep the M largest" just drops R of those);
//     it sorts one packed 64-bit word per particle and falls back to an exact sort of the full words only where two weights
//     agree in their top 56 bits, so exact ties are ordered canonically (below) without slowing the common site.
//   * the emission term logObs[t][r] is read from the T x R table K1 produced (48 B per site);
//   * transition terms come from a host-built table {c_new(d,r), log(1-rho(d,r))}; every particle holds the entries for d
//     and d+1 and loads the one for d+2 a whole site ahead, so no L2 latency sits on the per-site critical path;
//   * the R x N_prev new-segment log-sum-exps and backward kernels collapse to R class sums
//     E[r'] = sum_{n in class r'} W_n c_new(d_n, r') because logTrans((1,r) <- (d,r')) = log c_new(d,r') + log P[r'][r]
//     factorises (exact log-domain fallback when a class underflows);
//   * a SERVICE warp (no particles) draws the uniform, evaluates the scalar logs (log Z_t, the new-segment weights) and
//     prefetches the emission rows while the eight worker warps reduce, sort and scan;
//   * the uniform of site t is Philox(seed, chain, t) or an injected per-site array.
// Order of exactly equal weights: canonical (log-weight, then regime, then sojourn), see hyg_common.cuh and DESIGN.md C-14;
// sites where such a tie decided a particle's fate are counted (status[5]) and flagged (tie_flags tap).
// All arithmetic fp64.  Rounding differs from the reference at the 1e-16 level (tree sums vs sequential sums).
#ifndef HYG_SG_FILTER_CUH
#define HYG_SG_FILTER_CUH

#include "hyg_common.cuh"
#include "hyg_dev_structs.h"
#include "sg_param.cuh"

#define HYG_FATE_KEEP 0        // continues with its own weight
#define HYG_FATE_SURV 1        // drawn by the systematic resampling: continues with the common weight lsum - log C
#define HYG_FATE_DEAD 2        // + rank among the dead: the slot is reused by the new-segment particle of that regime
#define HYG_RES_DREW 1
#define HYG_RES_KEEP_LARGEST 2
#define HYG_RES_EXACT_SORT 4
#define HYG_WORKER_BAR 1       // named barrier of the 256 worker threads
#define HYG_SVC_BAR 14         // workers arrive, the service warp waits: the workers' partial sums are published
#define HYG_RES_BAR 15         // the service warp arrives, workers wait: the new-segment weights are published

namespace hyg {

struct SgResOut {
  int K;                          // k_kept tap
  int flags;
  int tie;                        // bit 0: equal keys among the sorted particles; bit 1: a tie decided a fate
  int n_dup;                      // systematic draws that hit one particle twice (rounding); resolved, counted
  double res_lw;                  // log-weight of a particle drawn by the systematic resampling, relative to log Z_{t-1}
};

struct SgSmem {
  double part[2][HYG_NW][8];      // per-warp partials of the 8-wide transposed reductions (double-buffered)
  double partG[2][HYG_NW][8];     // parameter mode: partials of Eg[r'] = sum e_n dlogrho_n
  unsigned vmask[2][HYG_NW];
  int infcnt[2][HYG_NW];          // per-warp counts of particles with log-weight -inf
  double W[HYG_NPMAX];            // previous self-normalised weights by slot (resampling input)
  unsigned short fate[HYG_NPMAX];
  unsigned short ofs[HYG_NPMAX];  // offspring counts (repair of double draws)
  SgResOut res;
  short new_slot[HYG_RMAX];       // slot of the new-segment particle (1, r) of this site
  unsigned long long xk[6][HYG_NPMAX];   // one exchange buffer per cross-warp sort stage (no reuse inside a site)
  unsigned long long srt[HYG_NPMAX + 1]; // sorted words, for the neighbour tests
  double Q[HYG_NPMAX + 1];
  double qtail[HYG_NW];
  int iscan[2][HYG_NW];
  int iflag[2][HYG_NW];
  BlockScratch sc;
  double lo[2][HYG_RMAX];
  double lomax[2];                // max_r logObs(t, r), [t & 1]
  int slow[HYG_RMAX];
  unsigned slowmask;              // bit r: regime r takes the log-domain path at this site
  double new_lw[HYG_RMAX];        // log-weight of the new-segment particle (1, r)
  double new_invE[HYG_RMAX];      // 1 / sumE[r]
  double u[2];                    // resampling uniform of site t, [t & 1]
  double lsum[2];                 // running log Z_t, [t & 1]
  // lag set (fixed-lag smoother)
  double bk[HYG_NPMAX][HYG_RMAX - 2];   // backward kernels bk_r[n] of the new-segment particles, [slot][r], r < 6
  double Wc[HYG_NPMAX];           // current weight of a continuing particle, 0 for a slot that was rewritten
  double Wnew[HYG_RMAX];          // current weight of the new-segment particle (1, r)
  double lag_m[8][HYG_RMAX];      // per pending site of the current batch: filtered means
  double lag_val[8][HYG_RMAX - 2][HYG_RMAX - 2];   // psi of the new-segment particles [q][r]
  int lag_ok[8][2];
  unsigned char r[HYG_NPMAX];     // regime by slot (lag set, parameter mode)
  unsigned long long hsum[HYG_NW];
  int cnt[8];                     // the chain's status words, kept by thread 0 (eight live counters per thread do not fit the register budget)
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

template <int RT> __device__ __forceinline__ double pick(const double (&v)[RT], int i) {
  double o = 0.0;
#pragma unroll
  for (int q = 0; q < RT; q++) o = (i == q) ? v[q] : o;
  return o;
}

// Transposed warp reduction of eight values per lane: 9 shuffle steps instead of 40.  On return every lane holds the
// warp total of value index (lane >> 2) & 7.
__device__ __forceinline__ double warp_reduce8(const double (&v)[8], int lane) {
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
__device__ __forceinline__ double warp_reduce_onehot(double e, int r, double x7, int lane) {
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
__device__ __forceinline__ void publish8(double (*part)[8], double wtot, int lane, int warp) {
  if ((lane & 3) == 0) part[warp][lane >> 2] = wtot;
}
// returns in every lane the block total of value index lane & 7
__device__ __forceinline__ double combine8(const double (*part)[8], int lane) {
  double a = part[lane >> 3][lane & 7] + part[(lane >> 3) + 4][lane & 7];
  a += __shfl_xor_sync(HYG_FULL, a, 8);
  a += __shfl_xor_sync(HYG_FULL, a, 16);
  return a;
}

#define HYG_LINEAR_FLOOR 1e-250

// Service-warp job: new-segment particles (1, r), r = lane < R.  sumE[r] = sum_{r' != r} P[r'][r] E[r'] is the linear-domain
// mass flowing into regime r (relative to exp(lsum_prev)); its log-weight is lsum_prev + logObs_r + log(sumE[r])
// (computeWeightsCp, Smc.h:562-573, after factorising logTrans((1,r) <- (d,r')) = log c_new(d,r') + log P[r'][r]).
template <int R>
__device__ __forceinline__ double sg_service_new_segments(const SgModelDev& mdl, SgSmem& s, double totA, int pA, int lane) {
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

// ... and their log-weights: one log per regime, evaluated by the service warp while the workers resample
template <int R>
__device__ __forceinline__ void sg_service_new_weights(const SgModelDev& mdl, SgSmem& s, double totA, int pA, int lane, double lsum_prev, const double* lo) {
  const double sumE_lane = sg_service_new_segments<R>(mdl, s, totA, pA, lane);
  if (lane < R) s.new_lw[lane] = (sumE_lane > 0.0) ? lsum_prev + lo[lane] + log(sumE_lane) : -HYG_INF;
  __syncwarp();   // the service warp's own lanes read s.slowmask next, without a block barrier in between
}

// plain load for tables that this kernel rewrites (parameter mode), read-only path otherwise
template <bool PE, class T> __device__ __forceinline__ T tab_load(const T* p) {
  if (PE) return *p;
  return __ldg(p);
}

// ------------------------------------------------------------------------------------------------------------------
// Resampling (Smc::resampleCp -> resample::optimalFiniteState, Smc.h:406-450, resample.h:289-409) by the 256 worker threads
// (barriers among the workers only; the service warp evaluates the new-segment weights meanwhile).
//   * Sort: one 64-bit word per particle -- the top 56 bits of the order-preserving image of its log-weight, its slot in the
//     low eight -- through a bitonic network: strides < 32 by warp shuffles, strides 32/64/128 through shared memory with
//     64-thread pair barriers (a warp only needs its partner warp's words).  If two neighbours of the result agree in the 56
//     bits (an exact tie of weights, or weights within 2^-44 of each other: ~5 % of the sites of one-sample data, ~0.05 % at 32
//     samples) the particles are sorted AGAIN on the full (key, regime, sojourn) words, so the canonical order (hyg_common.cuh,
//     DESIGN.md quirk C-14) is exact.
//   * K = the first sorted position whose weight is not above its own threshold Q[p]/(M-p): the reference iterates
//     K <- K + #{i >= K : log q_i > -log C(K)} from K = 0 (resample.h:333-342); along that iteration the threshold only
//     decreases, so it stops at that first position -- one parallel pass instead of ~100 sequential ones.
//   * Systematic resampling of the tail (resample.h:85-127,354-359) as a monotone cumulative-count scan.
//   * Results are scattered to the particles' slots as FATES (keep / drawn / dead + rank); nothing moves.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long sg_block_sort_packed(unsigned long long key, SgSmem& s, int tid, int warp) {
  int kbuf = 0;
#pragma unroll
  for (int k = 2; k <= HYG_NPMAX; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      unsigned long long other;
      if (j < 32) {
        other = __shfl_xor_sync(HYG_FULL, key, j);
      } else {
        s.xk[kbuf][tid] = key;
        if (kbuf == 0) {
          named_barrier(HYG_WORKER_BAR, HYG_NPMAX);   // also orders s.W (written by the caller) before its first read
        } else {
          // pair index: the warp number with the bit of the partner stride removed
          const int jw = j >> 5;   // 1, 2, 4
          const int pair = (warp & (jw - 1)) | ((warp & ~(2 * jw - 1)) >> 1);
          named_barrier(2 + 4 * (jw == 1 ? 0 : (jw == 2 ? 1 : 2)) + pair, 64);
        }
        other = s.xk[kbuf][tid ^ j];
        kbuf++;
      }
      const bool take_max = (((tid & j) == 0) == ((tid & k) == 0));
      key = (take_max == (other > key)) ? other : key;
    }
  }
  return key;
}

// exact canonical order: full 64-bit keys, ties by (regime, sojourn); rare, compact code instead of an unrolled network
__device__ __noinline__ void sg_block_sort_exact(unsigned long long& key, unsigned long long& pay, SgSmem& s, int tid) {
  int xb = 0;
#pragma unroll 1
  for (int k2 = 2; k2 <= HYG_NPMAX; k2 <<= 1) {
#pragma unroll 1
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      unsigned long long ok, op;
      if (j < 32) {
        ok = __shfl_xor_sync(HYG_FULL, key, j);
        op = __shfl_xor_sync(HYG_FULL, pay, j);
      } else {
        named_barrier(HYG_WORKER_BAR, HYG_NPMAX);   // the buffer's previous readers are done
        s.xk[xb][tid] = key; s.xk[xb + 2][tid] = pay;
        named_barrier(HYG_WORKER_BAR, HYG_NPMAX);
        ok = s.xk[xb][tid ^ j]; op = s.xk[xb + 2][tid ^ j];
        xb ^= 1;
      }
      const bool take_first = (((tid & j) == 0) == ((tid & k2) == 0));
      const bool tk = (take_first == order_before(ok, op, key, pay));
      key = tk ? ok : key;
      pay = tk ? op : pay;
    }
  }
  named_barrier(HYG_WORKER_BAR, HYG_NPMAX);
}

// key/pay: this slot's particle (key 0 = none); Wprev its self-normalised weight.  Writes s.fate[slot] for every particle and s.res.
template <int R>
__device__ __forceinline__ void sg_resample_block(SgSmem& s, unsigned long long key, unsigned long long pay, double Wprev, int N_prev, int M, double u,
                                                  bool force_exact, int tid, int lane, int warp) {
  int ibuf = 0;
  s.W[tid] = Wprev;
  const bool real = tid < N_prev;   // after the sort: sorted position tid holds a particle (empty slots sort last)
  unsigned long long pk = sg_block_sort_packed((key != 0ull) ? ((key & ~0xFFull) | static_cast<unsigned long long>(255 - tid)) : static_cast<unsigned long long>(255 - tid), s, tid, warp);
  int sidx = 255 - static_cast<int>(pk & 0xFFull);
  unsigned long long xkey = 0ull;   // exact key of this sorted position (only after the exact re-sort)
  bool exact = false;
  int K = 0;
  double qv = 0.0, Qk = 0.0;
  bool keep_largest = false;
  for (int pass = 0; pass < 2; pass++) {
    if (pass == 1) {
      // weights equal in their top 56 bits somewhere: the canonical order needs the full words
      unsigned long long k2 = key, p2 = (key != 0ull) ? pay : (~0ull - static_cast<unsigned long long>(tid));
      sg_block_sort_exact(k2, p2, s, tid);
      sidx = static_cast<int>(p2 & 0xFFull);
      xkey = k2; pk = k2;
      exact = true;
    }
    qv = real ? s.W[sidx] : 0.0;
    s.srt[tid] = pk;   // neighbours' words, for the tie tests
    double v = qv;     // Q[p] = sum_{j >= p} q_j (resample.h:306)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double tt = __shfl_down_sync(HYG_FULL, v, o);
      if (lane + o < 32) v += tt;
    }
    if (lane == 0) s.qtail[warp] = v;   // (not s.sc: its double-buffer index must stay in step with the service warp's)
    named_barrier(HYG_WORKER_BAR, HYG_NPMAX);
    double tail = 0.0;
#pragma unroll
    for (int w = HYG_WORKER_WARPS - 1; w > 0; w--)
      if (w > warp) tail += s.qtail[w];
    const double Qp = v + tail;
    s.Q[tid] = Qp;
    if (tid == 0) { s.Q[HYG_NPMAX] = 0.0; s.srt[HYG_NPMAX] = 0ull; }
    const unsigned long long nx = s.srt[tid + 1];
    // first pass: do two neighbours agree in the 56 bits the packed sort compared?  (zero weights / empty slots do not matter)
    const bool close = !exact && real && (tid + 1 < N_prev) && ((pk ^ nx) >> 8) == 0ull && (pk >> 8) != (HYG_KEY_NEGINF >> 8);
    const bool stop = (tid >= M) || !real || !(qv * static_cast<double>(M - tid) > Qp);
    const unsigned sb = __ballot_sync(HYG_FULL, stop);
    const unsigned cb = __ballot_sync(HYG_FULL, close);
    if (lane == 0) { s.iscan[ibuf][warp] = sb ? (warp * 32 + __ffs(static_cast<int>(sb)) - 1) : HYG_NPMAX; s.iflag[ibuf][warp] = cb ? 1 : 0; }
    named_barrier(HYG_WORKER_BAR, HYG_NPMAX);
    K = __reduce_min_sync(HYG_FULL, (lane < HYG_WORKER_WARPS) ? s.iscan[ibuf][lane] : HYG_NPMAX);
    const int anyclose = __reduce_max_sync(HYG_FULL, (lane < HYG_WORKER_WARPS) ? s.iflag[ibuf][lane] : 0);
    ibuf ^= 1;
    if (!anyclose && !(force_exact && pass == 0)) break;
  }
  keep_largest = (K >= M);
  if (!keep_largest) {
    Qk = s.Q[K];
    if (!(Qk > 0.0) || !hyg_isfinite(Qk)) keep_largest = true;
  }
  int fate = HYG_FATE_KEEP;
  int o = 0;
  int tie = 0, n_dup = 0;
  const unsigned long long nxk = exact ? s.srt[tid + 1] : 0ull;
  const bool eqnext = exact && real && (tid + 1 < N_prev) && xkey == nxk && xkey > HYG_KEY_NEGINF;
  if (exact && __any_sync(HYG_FULL, eqnext)) tie |= 1;
  if (keep_largest) {
    // keep the M largest by log-weight (Smc.h:432-441; resample.h:366-375): positions p >= M die
    fate = (real && tid >= M) ? HYG_FATE_DEAD : HYG_FATE_KEEP;
    if (tid == M - 1 && eqnext) tie |= 2;   // a tie across the cut decides who is kept
  } else {
    const int L = M - K;
    // C_p = #{ i < L : (i+u)/L <= cumulative residual weight up to p }; forced monotone, C_last = L, so the offspring counts
    // o_p = C_p - C_{p-1} are >= 0 and sum to L whatever the rounding of the suffix sums.  Residual fraction first, then x L:
    // L / Qk would overflow when the tail mass is subnormal (informative data).
    int C = 0;
    if (tid >= K && real) C = (tid == N_prev - 1) ? L : sys_count_x(((Qk - s.Q[tid + 1]) / Qk) * static_cast<double>(L) - u, L);
    if (!real) C = L;
#pragma unroll
    for (int dlt = 1; dlt < 32; dlt <<= 1) {
      const int tt = __shfl_up_sync(HYG_FULL, C, dlt);
      if (lane >= dlt) C = tt > C ? tt : C;
    }
    if (lane == 31) s.iscan[ibuf][warp] = C;
    named_barrier(HYG_WORKER_BAR, HYG_NPMAX);
    const int before = __reduce_max_sync(HYG_FULL, (lane < warp && lane < HYG_WORKER_WARPS) ? s.iscan[ibuf][lane] : 0);
    ibuf ^= 1;
    C = before > C ? before : C;
    int Cprev = __shfl_up_sync(HYG_FULL, C, 1);
    if (lane == 0) Cprev = before;
    if (tid <= K) Cprev = 0;
    o = (tid >= K && real) ? C - Cprev : 0;
    // A particle drawn twice: only possible when rounding puts a residual weight above the step Qk/L.  The reference would
    // duplicate the support point; here the extra draw goes to the next undrawn tail particle (repaired by one thread, counted).
    const bool anydup = __any_sync(HYG_FULL, o > 1);
    s.ofs[tid] = static_cast<unsigned short>(o);
    if (lane == 0) s.iflag[ibuf][warp] = anydup ? 1 : 0;
    named_barrier(HYG_WORKER_BAR, HYG_NPMAX);
    const int dupany = __reduce_max_sync(HYG_FULL, (lane < HYG_WORKER_WARPS) ? s.iflag[ibuf][lane] : 0);
    if (dupany) {
      if (tid == 0) {
        int extra = 0;
        for (int j = K; j < N_prev; j++) { const int oj = s.ofs[j]; if (oj > 1) { extra += oj - 1; s.ofs[j] = 1; } }
        s.iflag[ibuf][0] = extra;
        for (int j = K; j < N_prev && extra > 0; j++) if (s.ofs[j] == 0) { s.ofs[j] = 1; extra--; }
      }
      named_barrier(HYG_WORKER_BAR, HYG_NPMAX);
      o = s.ofs[tid];
      n_dup = s.iflag[ibuf][0];
      named_barrier(HYG_WORKER_BAR, HYG_NPMAX);
    }
    ibuf ^= 1;
    if (exact) {
      // two equal residual weights, one drawn and one not: the tie decided
      const int nxo = (tid + 1 < HYG_NPMAX) ? static_cast<int>(s.ofs[tid + 1]) : 0;
      if (__any_sync(HYG_FULL, eqnext && tid >= K && qv != 0.0 && ((o > 0) != (nxo > 0)))) tie |= 2;
    }
    fate = (!real || tid < K) ? HYG_FATE_KEEP : (o > 0 ? HYG_FATE_SURV : HYG_FATE_DEAD);
  }
  // ranks of the dead in sorted order: the slot of the k-th dead particle is reused by the new-segment particle (1, k)
  const unsigned dm = __ballot_sync(HYG_FULL, fate == HYG_FATE_DEAD);
  if (lane == 0) { s.iscan[ibuf][warp] = __popc(dm); s.iflag[ibuf][warp] = tie; }
  named_barrier(HYG_WORKER_BAR, HYG_NPMAX);
  int drank = __popc(dm & ((1u << lane) - 1u));
#pragma unroll
  for (int w = 0; w < HYG_WORKER_WARPS; w++) {
    if (w < warp) drank += s.iscan[ibuf][w];
    tie |= s.iflag[ibuf][w];
  }
  if (fate == HYG_FATE_DEAD) fate = HYG_FATE_DEAD + drank;
  if (real) s.fate[sidx] = static_cast<unsigned short>(fate);
  if (tid == 0) {
    s.res.K = keep_largest ? -2 : K;
    s.res.flags = (keep_largest ? HYG_RES_KEEP_LARGEST : HYG_RES_DREW) | (exact ? HYG_RES_EXACT_SORT : 0);
    s.res.tie = tie;
    s.res.n_dup = n_dup;
    if (!keep_largest) s.res.res_lw = log(Qk / static_cast<double>(M - K));   // + lsum_prev: resample.h:361-364
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Fixed-lag smoother: updatePsi + storeEstimates of every pending site (OnlineMarginalSmoothing.h:148-255).
// psi rows live in a per-CTA global workspace (L2-resident) and never move: rows[row][q][slot]; a continuing particle's
// entry is untouched, only the <= R rewritten slots get sum_n bk_r[n] psi[n].  One warp per (pending site, three regime
// indicators): lanes stride over the slots, so the 256-term sums are serial per lane and finish with one 8-wide transposed
// warp reduction -- no block barrier per pending site.
// ------------------------------------------------------------------------------------------------------------------
struct SgLagState {
  // One pointer, the capacity and a flip bit instead of six pointers: these live in every thread for the whole chain, and at
  // 72 registers per thread each pointer that is not needed at a site is a spill.
  double* rows;        // [lcap][R][256], then five int lists of lcap entries each:
  int lcap;            //   pend_t / pend_row (site and row of pending entry i, oldest first), their alternates (the lists are
  int flip;            //   rebuilt into the alternate copy while the current one is being read), the stack of free rows
  int row_doubles;     // R x 256
  int n_pend, n_free;
  __device__ __forceinline__ int* lists(int k) const { return reinterpret_cast<int*>(rows + static_cast<size_t>(lcap) * row_doubles) + static_cast<size_t>(k) * lcap; }
  __device__ __forceinline__ int* pend_t() const { return lists(flip ? 2 : 0); }
  __device__ __forceinline__ int* pend_row() const { return lists(flip ? 3 : 1); }
  __device__ __forceinline__ int* pend_t_alt() const { return lists(flip ? 0 : 2); }
  __device__ __forceinline__ int* pend_row_alt() const { return lists(flip ? 1 : 3); }
  __device__ __forceinline__ int* free_row() const { return lists(4); }
};

template <int R>
__device__ __noinline__ int sg_lag_update(SgLagState& lag, const SgChainDev& ch, SgSmem& s, unsigned int t, unsigned int T, unsigned int own_lo,
                                          unsigned int own_hi, unsigned long long t_off, bool last_seg, double epsilon) {
  static_assert(R <= HYG_RMAX - 2, "lag-set tasks handle three regime indicators each");
  const int tid = hyg_tid(), lane = tid & 31, warp = tid >> 5;
  // s.bk[slot][r], s.Wc[slot], s.new_slot[r] and s.Wnew[r] were published before the caller's barrier
  constexpr int HALF = (R + 1) / 2;
  int kept = 0;
  const int n_pend = lag.n_pend;
  for (int i0 = 0; i0 < n_pend; i0 += 8) {
    const int nb = (n_pend - i0) < 8 ? (n_pend - i0) : 8;
    // ---- tasks (pending site, half): all nine warps take their share ----
    for (int task = warp; task < 2 * nb; task += HYG_NW) {
      const int bi = task >> 1, h = task & 1;
      const int q0 = h * HALF, nq = (h == 0) ? HALF : (R - HALF);
      double* row = lag.rows + static_cast<size_t>(lag.pend_row()[i0 + bi]) * R * HYG_NPMAX;
      double acc[HALF][8];
#pragma unroll
      for (int j = 0; j < HALF; j++)
#pragma unroll
        for (int c = 0; c < 8; c++) acc[j][c] = 0.0;
#pragma unroll 2
      for (int k = 0; k < HYG_NPMAX / 32; k++) {
        const int n = lane + 32 * k;
        double b[R];
#pragma unroll
        for (int r = 0; r < R; r++) b[r] = s.bk[n][r];
        const double w = s.Wc[n];
#pragma unroll
        for (int j = 0; j < HALF; j++) {
          if (j < nq) {
            const double ps = row[(q0 + j) * HYG_NPMAX + n];
#pragma unroll
            for (int r = 0; r < R; r++) acc[j][r] += b[r] * ps;
            const double wp = w * ps;
            acc[j][6] += wp;
            acc[j][7] += wp * ps;
          }
        }
      }
      bool ok = true;
#pragma unroll
      for (int j = 0; j < HALF; j++) {
        if (j < nq) {
          const double tot = warp_reduce8(acc[j], lane);        // lane group g = (lane >> 2) & 7 holds the total of index g
          const int g = (lane >> 2) & 7;
          // psi of the new-segment particle (1, g) is sum_n bk_g[n] psi[n] = tot; it carries the weight Wnew[g]
          const double wg = (g < R) ? s.Wnew[g] : 0.0;
          double mterm = (g < R) ? wg * tot : ((g == 6) ? tot : 0.0);
          double vterm = (g < R) ? wg * tot * tot : ((g == 7) ? tot : 0.0);
          mterm += __shfl_xor_sync(HYG_FULL, mterm, 4); vterm += __shfl_xor_sync(HYG_FULL, vterm, 4);
          mterm += __shfl_xor_sync(HYG_FULL, mterm, 8); vterm += __shfl_xor_sync(HYG_FULL, vterm, 8);
          mterm += __shfl_xor_sync(HYG_FULL, mterm, 16); vterm += __shfl_xor_sync(HYG_FULL, vterm, 16);
          const double var = vterm - mterm * mterm;       // sum W (x - m)^2 with sum W = 1
          if (!(var < epsilon)) ok = false;
          if (lane == 0) s.lag_m[bi][q0 + j] = mterm;
          if ((lane & 3) == 0 && g < R) s.lag_val[bi][q0 + j][g] = tot;
        }
      }
      if (lane == 0) s.lag_ok[bi][h] = ok ? 1 : 0;
    }
    __syncthreads();
    // ---- storeEstimates (OnlineMarginalSmoothing.h:197-255): emit when all R filtered variances < epsilon ----
    for (int bi = 0; bi < nb; bi++) {
      const int i = i0 + bi;
      const bool settled = s.lag_ok[bi][0] && s.lag_ok[bi][1];
      const bool emit = settled || (t == T - 1);
      const int ts = lag.pend_t()[i];
      const int rw = lag.pend_row()[i];
      if (emit) {
        const bool own_s = (static_cast<unsigned int>(ts) >= own_lo) && (static_cast<unsigned int>(ts) < own_hi);
        if (own_s) {
          // whole row (position, p_1..p_R): 56 contiguous bytes, also when the row goes to mapped host memory
          if (tid <= R && ch.probs) {
            const double outv = (tid == 0) ? (ch.pos ? static_cast<double>(ch.pos[ts]) : static_cast<double>(static_cast<unsigned long long>(ts) + t_off))
                                           : s.lag_m[bi][tid - 1];
            ch.probs[static_cast<size_t>(ts) * (R + 1) + tid] = outv;
          }
          if (tid == 0 && ch.finalised_at) ch.finalised_at[ts] = static_cast<int>(t + t_off);
          if (!settled && !last_seg && tid == 0) s.cnt[2]++;   // the segment's right halo ended before this site settled
        } else if (ch.ovl && tid < R && static_cast<unsigned int>(ts) >= own_hi && static_cast<unsigned int>(ts) < own_hi + HYG_OVL_ROWS) {
          ch.ovl[(static_cast<unsigned int>(ts) - own_hi) * R + tid] = s.lag_m[bi][tid];   // left-halo check (hyg_dev_structs.h)
        }
        if (tid == 0) lag.free_row()[lag.n_free] = rw;
        lag.n_free++;
      } else {
        // psi of the rewritten slots; every other entry of the row stays as it is
        if (tid < R * R) {
          const int q = tid / R, r = tid % R;
          lag.rows[static_cast<size_t>(rw) * R * HYG_NPMAX + q * HYG_NPMAX + s.new_slot[r]] = s.lag_val[bi][q][r];
        }
        if (tid == 0) { lag.pend_t_alt()[kept] = ts; lag.pend_row_alt()[kept] = rw; }
        kept++;
      }
    }
    __syncthreads();
  }
  lag.flip ^= 1;
  lag.n_pend = kept;
  return kept;
}

struct SgChainState {
  // per-thread particle (slot = threadIdx.x)
  double lw, W;
  double2 cur, nxt;    // table entries {c_new, log(1 - rho)} for d and d + 1
  double gcur, gnxt;   // parameter mode: d log rho / d theta_omega for d and d+1
  uint32_t d;
  int r;
};

template <int RT, bool PE>
__device__ void sg_filter_chain(SgModelDev& mdl, const SgChainDev& ch, const SgRunDev& run, double* psi_ws, SgSmem& s, SgPeSmem<RT>* pe) {
  static_assert(RT <= 7, "class sums share an 8-wide reduction with the finite-weight count");
  const int tid = hyg_tid(), lane = tid & 31, warp = tid >> 5;
  // Warps 0..7 own the particles (one slot per thread).  Warp 8 is a SERVICE warp: it owns no particle, follows the same
  // block barriers, and evaluates every scalar exp/log of the step (new-segment weights, log C, log Z_t, the Philox draw,
  // the emission-row prefetch).
  const bool worker = warp < HYG_WORKER_WARPS;
  const bool service = !worker;
  constexpr int R = RT;
  const int Nmax = mdl.n_particles;
  uint32_t dcap = mdl.dcap;
  const unsigned int T = static_cast<unsigned int>(ch.T);   // sites of this unit (< 2^32): 32-bit counters in the hot loop
  const int lcap = run.lcap;
  int flip = 0, pbuf = 0;
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
    for (int i = tid; i < D * HYG_PHI_PITCH; i += HYG_NT) (&pe->phi[0][0])[i] = 0.0;
    __syncthreads();
    pe_set_theta<R>(mdl, *pe);
    uint32_t dn = run.n_steps_without_update + 12;
    dn = dn > run.pe_dcap ? run.pe_dcap : dn;
    pe_rebuild_tables<R>(mdl, *pe, pe_tab, pe_tabg, pe_wh, pe_wg, run.pe_dcap, dn);
    dcap = run.pe_dcap;   // row pitch of the workspace tables; valid entries: mdl.dcap
  }
  const SgModelDev& cm = mdl;

  // lag-set workspace (global, L2-resident): [lcap][R][256] doubles, then five int arrays of lcap entries
  SgLagState lag;
  lag.rows = psi_ws; lag.lcap = lcap; lag.flip = 0; lag.row_doubles = R * HYG_NPMAX;
  lag.n_pend = 0; lag.n_free = lcap;
  for (int i = tid; i < lcap; i += HYG_NT) lag.free_row()[i] = lcap - 1 - i;
  if (tid < 8) s.cnt[tid] = 0;   // read and written by thread 0 only from here on
  // segmented execution: local sites [own_lo, own_hi) are written, the rest is warm-up / run-out (hyg_dev_structs.h)
  const unsigned int own_lo = static_cast<unsigned int>(ch.own_lo), own_hi = static_cast<unsigned int>(ch.own_hi);
  const unsigned long long t_off = ch.t_off;
  const bool last_seg = ch.last_segment != 0;
  const unsigned int run_to = own_hi + (ch.ovl ? HYG_OVL_ROWS : 0u);
  // which optional outputs exist, in a register: the descriptor lives in shared memory and a pointer test per site and output
  // is a load + compare on the critical path
  const unsigned taps = (ch.probs ? 1u : 0u) | (ch.finalised_at ? 2u : 0u) | (ch.ovl ? 4u : 0u) | (ch.support_hash ? 8u : 0u) |
                        ((ch.k_kept || ch.drew || ch.n_pending || ch.n_curr || ch.tie_flags) ? 16u : 0u) | (ch.logz ? 32u : 0u) | (ch.seg_inc ? 64u : 0u);
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
    bool drew = false, resampled = false;
    int tie_site = 0;
    bool emit_now = false;       // current site finalised at this step
    const bool own_t = (t >= own_lo) && (t < own_hi);
    double cw_lane = 0.0;        // regime mass of index lane & 7 (current site)

    if (t > 0) {
      // =========================== Smc::iterate (Smc.h:190-286) ===========================
      const int N_prev = N;
      const int N_curr = (N_prev + R > Nmax) ? Nmax : N_prev + R;
      const int M = N_curr - R;
      const bool capped = (N_curr < N_prev + R);
      const int nd = N_prev - M;   // particles that die at this site (0 in the growth phase)
      const double lomax = s.lomax[t & 1];
      const bool alive = tid < N_prev;   // service warp: never

      // (the uniform of this site and log Z_{t-1} were evaluated by the service warp behind the normaliser barrier of site t-1)
      // ---- class sums over the previous particles (replace the R x N_prev log-sum-exps of Smc.h:562-573) ----
      const double e_prev = alive ? p.W * p.cur.x : 0.0;  // W_n * c_new(d_n, r_n)
      const bool finite_prev = alive && hyg_isfinite(p.lw);
      const bool valid = finite_prev && (p.cur.x > 0.0);
      unsigned infmask = 0;
      if (worker) {
        // class sums of e_prev; slot 7: F = #finite(logw_prev), Smc.h:413 (exact in fp64).  e_prev is 0 beyond N_prev.
        publish8(s.part[pbuf], warp_reduce_onehot(e_prev, p.r, finite_prev ? 1.0 : 0.0, lane), lane, warp);
        const unsigned vm = __reduce_or_sync(HYG_FULL, valid ? (1u << p.r) : 0u);
        if (lane == 0) s.vmask[pbuf][warp] = vm;
        if (PE) {
          pe->eprev[tid] = e_prev;
          publish8(s.partG[pbuf], warp_reduce_onehot(e_prev * p.gcur, p.r, 0.0, lane), lane, warp);
        }
        if (capped) {
          infmask = __ballot_sync(HYG_FULL, alive && !finite_prev);
          if (lane == 0) s.infcnt[pbuf][warp] = __popc(infmask);
        }
      }
      const int pA = pbuf;
      pbuf ^= 1;
      // ---- B1 ---- the workers synchronise among themselves and signal the service warp; they never wait for it here
      if (worker) { named_arrive(HYG_SVC_BAR, HYG_NT); named_barrier(HYG_WORKER_BAR, HYG_NPMAX); }
      else named_barrier(HYG_SVC_BAR, HYG_NT);
      const double totA = combine8(s.part[pA], lane);           // lane & 7 -> E[0..R-1], [7] = F
      const int F = static_cast<int>(__shfl_sync(HYG_FULL, totA, 7) + 0.5);
      if (service) {
        const double lsum_prev = s.lsum[(t + 1) & 1];
        if (PE) {
          const double totG = combine8(s.partG[pA], lane);
          if (lane < 8) { pe->Etot[lane] = totA; pe->Egtot[lane] = totG; }
        }
        sg_service_new_weights<R>(mdl, s, totA, pA, lane, lsum_prev, lo);   // while the workers resample
      }

      // ---- who dies: Smc::resampleCp (Smc.h:406-450) ----
      int fate = HYG_FATE_KEEP;
      if (capped) {
        if (F <= M) {
          // keep the M largest (Smc.h:432-441): at least nd particles have weight zero, and which of them go is immaterial
          int irank = __popc(infmask & ((1u << lane) - 1u));
          if (worker) {
#pragma unroll
            for (int w = 0; w < HYG_WORKER_WARPS; w++)
              if (w < warp) irank += s.infcnt[pA][w];
          }
          if (alive && !finite_prev && irank < nd) fate = HYG_FATE_DEAD + irank;
          k_kept = -2;
        } else {
          // ---- resample::optimalFiniteState (resample.h:289-409) ----
          if (worker) sg_resample_block<R>(s, alive ? order_key(p.lw) : 0ull, order_pay(p.r, p.d, tid), alive ? p.W : 0.0, N_prev, M, s.u[t & 1], run.force_full_sort != 0, tid, lane, warp);
          resampled = true;
        }
      }
      // ---- B3 ---- the new-segment weights (service warp) and the fates (workers) are published; the service warp does not wait
      if (worker) named_barrier(HYG_RES_BAR, HYG_NT);
      else named_arrive(HYG_RES_BAR, HYG_NT);
      const double lsum_prev = s.lsum[(t + 1) & 1];   // written by the service warp behind the normaliser barrier of site t-1
      if (worker && resampled) {
        if (alive) fate = s.fate[tid];
        k_kept = s.res.K;
        drew = (s.res.flags & HYG_RES_DREW) != 0;
        tie_site = s.res.tie;
        if (tid == 0) {
          s.cnt[4] += (s.res.flags & HYG_RES_EXACT_SORT) ? 1 : 0;
          s.cnt[6] += s.res.n_dup;
          s.cnt[5] += (tie_site & 2) ? 1 : 0;
        }
      }

      // ---- propose + weight: sampleParticlesCp / computeWeightsCp (Smc.h:504-574) ----
      // the k-th dead slot (sorted order) takes the new-segment particle (1, k); the others open the slots N_prev, N_prev+1, ..
      int newreg = -1;
      if (fate >= HYG_FATE_DEAD) newreg = fate - HYG_FATE_DEAD;
      else if (tid >= N_prev && tid < N_curr) newreg = nd + (tid - N_prev);
      HYG_CHECK(newreg < R, 2, newreg, fate);
      if (newreg >= 0) s.new_slot[newreg] = static_cast<short>(tid);
      const double my_e = e_prev;            // backward-kernel numerator of the particle that WAS in this slot
      const int my_r_prev = p.r;
      const bool was_valid = valid;
      const double lw_prev = p.lw, cur_x_prev = p.cur.x, gcur_prev = p.gcur;
      const uint32_t vcap = PE ? cm.dcap : dcap;   // valid table entries per regime (parameter mode: last rebuild)
      double grad_c = 0.0;   // parameter mode: d log f / d theta_omega(r) of the continuation (singleGroup.h:679-693)
      const bool cont = alive && newreg < 0;
      if (cont) {
        const double2 pc = p.cur;
        const double lc = pc.y;                         // log(1 - rho(d_prev, r)) or -inf (singleGroup.h:597-605)
        p.lw = ((fate == HYG_FATE_SURV) ? lsum_prev + s.res.res_lw : p.lw) + (lc + lo[p.r]);
        p.d = p.d + 1;
        p.cur = p.nxt;
        const uint32_t di = (p.d + 1 <= vcap) ? p.d : vcap - 1;  // 0-based index of d+1, clamped to the terminal entry
        HYG_CHECK(p.r >= 0 && p.r < R && di < dcap, 3, p.r, di);
        p.nxt = tab_load<PE>(mdl.tab + static_cast<size_t>(p.r) * dcap + di);
        if (PE) {
          const double rho = pc.x;                      // c_new = rho for d >= u (0 below u)
          grad_c = (lc > -HYG_INF && rho < 1.0) ? -p.gcur * rho / (1.0 - rho) : 0.0;
          p.gcur = p.gnxt;
          p.gnxt = mdl.tabg[static_cast<size_t>(p.r) * dcap + di];
        }
      } else if (newreg >= 0) {
        const int r = newreg;
        p.r = r; p.d = 1;
        p.cur = tab_load<PE>(mdl.tab + static_cast<size_t>(r) * dcap + 0);
        p.nxt = tab_load<PE>(mdl.tab + static_cast<size_t>(r) * dcap + 1);
        if (PE) { p.gcur = mdl.tabg[static_cast<size_t>(r) * dcap + 0]; p.gnxt = mdl.tabg[static_cast<size_t>(r) * dcap + 1]; }
        p.lw = s.new_lw[r];
      }
      const bool now_alive = tid < N_curr;
      // exact log-domain path for regimes whose linear-domain sum underflowed (rare)
      const unsigned slowmask = s.slowmask;
      double bk_slow[R];
#pragma unroll
      for (int r = 0; r < R; r++) bk_slow[r] = 0.0;
      if (slowmask) {
#pragma unroll
        for (int r = 0; r < R; r++) {
          if (!((slowmask >> r) & 1u)) continue;
          const bool ok = was_valid && my_r_prev != r && mdl.P[my_r_prev][r] > 0.0;
          const double x = ok ? lw_prev + log(cur_x_prev) + mdl.logP[my_r_prev][r] : -HYG_INF;
          const double mx = block_max(x, s.sc, flip);
          double ex[1] = {ok ? exp(x - mx) : 0.0};
          const double mine = ex[0];
          block_sum<1>(ex, s.sc, flip);
          if (newreg == r) p.lw = lo[r] + (mx + log(ex[0]));
          bk_slow[r] = (ex[0] > 0.0) ? mine / ex[0] : 0.0;
        }
      }

      // ---- selfNormaliseWeights (Smc.h:576-579), fused with the regime masses of the new site ----
      {
        double shift = lsum_prev + lomax;   // upper bound of every logw instead of the exact max
        p.W = now_alive ? exp(p.lw - shift) : 0.0;
        if (worker) publish8(s.part[pbuf], warp_reduce_onehot(p.W, p.r, 0.0, lane), lane, warp);   // p.W is 0 beyond N_curr
        __syncthreads();   // ---- B4 ---- (a full barrier: two worker arrivals on HYG_SVC_BAR must never be outstanding at once)
        double tot = combine8(s.part[pbuf], lane);   // lane & 7 -> class sum of the relative weights
        pbuf ^= 1;
        double S = tot;
        S += __shfl_xor_sync(HYG_FULL, S, 1);
        S += __shfl_xor_sync(HYG_FULL, S, 2);
        S += __shfl_xor_sync(HYG_FULL, S, 4);
        if (!(S > 0.0) || !hyg_isfinite(S)) {
          // the linear-domain weights underflowed against the bound: renormalise from the log-weights with the exact max
          shift = block_max(now_alive ? p.lw : -HYG_INF, s.sc, flip);
          p.W = (now_alive && p.lw > -HYG_INF) ? exp(p.lw - shift) : 0.0;
          if (worker) publish8(s.part[pbuf], warp_reduce_onehot(p.W, p.r, 0.0, lane), lane, warp);
          __syncthreads();
          tot = combine8(s.part[pbuf], lane);
          pbuf ^= 1;
          S = tot;
          S += __shfl_xor_sync(HYG_FULL, S, 1);
          S += __shfl_xor_sync(HYG_FULL, S, 2);
          S += __shfl_xor_sync(HYG_FULL, S, 4);
        }
        const double invS = 1.0 / S;
        p.W *= invS;
        cw_lane = tot * invS;
        pend_shift = shift; pend_S = S;
        if (service) {
          // log Z_t and the uniform of site t+1, while the workers finish the site: nothing waits for them before the next
          // site's first barrier
          if (lane == 0) s.lsum[t & 1] = shift + log(S);
          // (two sites ahead: the workers read it before the next barrier the service warp takes part in)
          if (lane == 8 && t + 2 < T) s.u[t & 1] = ch.unif ? __ldg(ch.unif + t + 2) : philox_uniform(ch.seed, ch.chain_id, static_cast<unsigned long long>(t) + 2ull + t_off);
          // the resampling taps of this site, for the end-of-step taps below (stable until the next site's resampling)
          if (resampled) { k_kept = s.res.K; drew = (s.res.flags & HYG_RES_DREW) != 0; tie_site = s.res.tie; }
        }
      }

      // ---- fixed-lag smoothing: updatePsi (OnlineMarginalSmoothing.h:148-177) ----
      if (run.use_smoothing && lag.n_pend > 0) {
        if (worker) {
          // backward kernels of the new-segment particles, bk_r[n] = e_n P[r_n][r] / sumE[r] (Smc.h:288-326, factorised)
#pragma unroll
          for (int r = 0; r < R; r++) {
            const double lin = my_e * mdl.P[my_r_prev][r] * s.new_invE[r];
            s.bk[tid][r] = ((slowmask >> r) & 1u) ? bk_slow[r] : lin;
          }
          s.Wc[tid] = cont ? p.W : 0.0;
          if (newreg >= 0) s.Wnew[newreg] = p.W;       // weights of the rewritten slots
        }
        __syncthreads();
        sg_lag_update<R>(lag, ch, s, t, T, own_lo, own_hi, t_off, last_seg, run.epsilon);
      }
      // ---- K3: score recursion (OnlineParameterEstimation.h:135-158), see sg_param.cuh ----
      if (PE) {
        if (tid < 8 * D) {
          // class sums G[k][r'] = sum_{n in class r'} e_n phi_n[k]: thread = (component k, chunk of 32 previous particles)
          const int k = tid % D, chunk = tid / D;
          double acc[R];
#pragma unroll
          for (int q = 0; q < R; q++) acc[q] = 0.0;
          for (int i = 0; i < 32; i++) {
            const int n = chunk * 32 + i;
            const double ev = pe->eprev[n] * pe->phi[k][n];
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
        if (slowmask) {
          // regimes on the log-domain path: phi' = sum_n bk_r[n] (phi_n + grad_n) with the per-particle normalised kernel;
          // staged in pe->part (free again) because phi is updated in place afterwards
#pragma unroll
          for (int r = 0; r < R; r++) {
            if (!((slowmask >> r) & 1u)) continue;
            for (int k = 0; k < D; k++) {
              double v[1] = {0.0};
              if (bk_slow[r] > 0.0) {
                const int rn = my_r_prev;
                double g = pe->phi[k][tid];
                if (k == R * (R - 1) + rn) {
                  g += gcur_prev;
                } else if (k >= rn * (R - 1) && k < (rn + 1) * (R - 1)) {
                  const int j = k - rn * (R - 1);
                  const int col = (j < rn) ? j : j + 1;
                  g += ((col == r) ? 1.0 : 0.0) - cm.P[rn][col];
                }
                v[0] = bk_slow[r] * g;
              }
              block_sum<1>(v, s.sc, flip);
              if (tid == 0) pe->part[0][k][r] = v[0];
            }
          }
          __syncthreads();
        }
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
          pe->phi[k][s.new_slot[r]] = ((slowmask >> r) & 1u) ? pe->part[0][k][r] : a * s.new_invE[r];
        }
        // continuing particle: phi' = phi + grad, grad touches only the omega slot of its regime
        if (cont) pe->phi[R * (R - 1) + p.r][tid] += grad_c;
      }
      N = N_curr;
    } else {
      // t = 0: regime masses of the initial particle system
      double v8[8];
      if (worker) {
#pragma unroll
        for (int q = 0; q < 8; q++) v8[q] = (q < R && tid < N && p.r == q) ? p.W : 0.0;
        publish8(s.part[pbuf], warp_reduce8(v8, lane), lane, warp);
      }
      if (service && lane >= 8 && lane <= 9 && lane - 7 < static_cast<int>(T)) {   // uniforms of sites 1 and 2
        const unsigned long long tt = static_cast<unsigned long long>(lane - 7);
        s.u[tt & 1] = ch.unif ? __ldg(ch.unif + tt) : philox_uniform(ch.seed, ch.chain_id, tt + t_off);
      }
      __syncthreads();
      cw_lane = combine8(s.part[pbuf], lane);
      pbuf ^= 1;
    }
    if (worker && (PE || run.use_smoothing)) s.r[tid] = static_cast<unsigned char>(p.r);

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
      if (!emit_now && lag.n_pend >= lcap) { emit_now = true; if (tid == 0 && own_t) s.cnt[0]++; }  // lag set full: emit the filtering estimate now (reported)
      if (emit_now) {
        if (own_t) {
          if (warp == 0 && (taps & 1u)) {
            const double left = __shfl_up_sync(HYG_FULL, cw_lane, 1);   // lane j >= 1: mass of regime j-1
            if (lane <= R) ch.probs[static_cast<size_t>(t) * (R + 1) + lane] = (lane == 0) ? pos_cur : left;
          }
          if (tid == 0 && (taps & 2u)) ch.finalised_at[t] = static_cast<int>(t + t_off);
          if (t == T - 1 && !settled && !last_seg && tid == 0) s.cnt[2]++;
        } else if ((taps & 4u) && warp == 0 && t >= own_hi && t < own_hi + HYG_OVL_ROWS) {
          const double left = __shfl_up_sync(HYG_FULL, cw_lane, 1);
          if (lane >= 1 && lane <= R) ch.ovl[(t - own_hi) * R + (lane - 1)] = left;
        }
      } else {
        // free_row / pend_* were last written before a block barrier of this step (or at initialisation)
        HYG_CHECK(lag.n_free >= 1 && lag.n_free <= lcap, 6, lag.n_free, lag.n_pend);
        const int rw = lag.free_row()[lag.n_free - 1];
        HYG_CHECK(rw >= 0 && rw < lcap, 7, rw, lag.n_free);
        double* dst = lag.rows + static_cast<size_t>(rw) * R * HYG_NPMAX;
        if (worker) {
#pragma unroll
          for (int q = 0; q < R; q++) dst[q * HYG_NPMAX + tid] = (tid < N && p.r == q) ? 1.0 : 0.0;
        }
        if (tid == 0) { lag.pend_t()[lag.n_pend] = static_cast<int>(t); lag.pend_row()[lag.n_pend] = rw; }
        lag.n_pend++; lag.n_free--;
      }
      if (tid == 0 && lag.n_pend > s.cnt[1]) s.cnt[1] = lag.n_pend;
    }

    if (service && t + 2 < T) {
      if (lane < R) s.lo[t & 1][lane] = lo_pref;
      double mx = (lane < R) ? lo_pref : -HYG_INF;
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) { const double tt = __shfl_xor_sync(HYG_FULL, mx, o); mx = tt > mx ? tt : mx; }
      if (lane == 0) s.lomax[t & 1] = mx;
    }
    // Everything written above is next read behind one of the block barriers of the following site.  The one exception
    // is the lag-set list head the exit test below reads, so only that test pays for a barrier.
    // Segmented execution: stop as soon as the owned range is stepped through and none of its sites is still pending
    // (the lag set is ordered by site, oldest first); with the left-halo check on, a segment first steps through
    // HYG_OVL_ROWS sites of the next one.
    bool exit_now = false;
    if (!PE && t + 1 >= run_to && t + 1 < T) {
      __syncthreads();
      exit_now = (lag.n_pend == 0) || (static_cast<unsigned int>(lag.pend_t()[0]) >= own_hi);
    }
    const bool last_step = (t == T - 1) || exit_now;
    if (tid == 0) s.cnt[3]++;

    // ---- K3: parameter update every n_steps sites (OnlineParameterEstimation.h:51-61) ----
    if (PE) {
      if (t > 0 && (t % run.n_steps_without_update) == 0) {
        if (worker) s.W[tid] = p.W;
        __syncthreads();
        if (tid < D) {
          // g = sum_n W_n phi_n over the current particles (computeFilteredMean, Smc.h:340-349)
          double g = 0.0;
          for (int n = 0; n < N; n++) g = g + s.W[n] * pe->phi[tid][n];
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
        }
        __syncthreads();
      }
      if (ch.theta_trace && tid < D) ch.theta_trace[static_cast<size_t>(t) * D + tid] = pe->theta[tid];
    }

    // ---- taps ----
    if (taps & 8u) {
      // order-independent hash of the finite-weight support {(d, r)} (parity tests only)
      unsigned long long h = (worker && tid < N && p.lw > -HYG_INF) ? mix64((static_cast<unsigned long long>(p.r) << 28) | p.d) : 0ull;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(HYG_FULL, h, o);
      if (lane == 0) s.hsum[warp] = h;
      __syncthreads();
      if (tid == 0 && own_t) {
        unsigned long long tot = 0ull;
        for (int w = 0; w < HYG_NW; w++) tot += s.hsum[w];
        ch.support_hash[t] = tot;
      }
      __syncthreads();
    }
    if (service && lane == 0) {
      if (t > 0) {
        const double lz_prev = s.lsum[(t + 1) & 1];   // log Z (local) of site t-1
        if (t == own_lo) lz_base = lz_prev;
        if (t - 1 >= own_lo && t - 1 < own_hi) {
          if (taps & 32u) ch.logz[t - 1] = lz_prev - lz_base;
          if (t == own_hi && (taps & 64u)) *ch.seg_inc = lz_prev - lz_base;
        }
      }
      if (last_step && own_t) {
        const double lz = s.lsum[t & 1] - lz_base;
        if (taps & 32u) ch.logz[t] = lz;
        if (t + 1 == own_hi && (taps & 64u)) *ch.seg_inc = lz;
      }
      if (own_t && (taps & 16u)) {
        if (ch.k_kept) ch.k_kept[t] = k_kept;
        if (ch.drew) ch.drew[t] = drew ? 1 : 0;
        if (ch.n_pending) ch.n_pending[t] = lag.n_pend;
        if (ch.n_curr) ch.n_curr[t] = N;
        if (ch.tie_flags) ch.tie_flags[t] = static_cast<unsigned char>(tie_site);
      }
    }
    if (exit_now) break;
  }
  if (tid == 0 && ch.status) {
    atomicAdd(ch.status + 0, s.cnt[0]);
    atomicMax(ch.status + 1, s.cnt[1]);
    atomicAdd(ch.status + 2, s.cnt[2]);
    atomicAdd(ch.status + 3, s.cnt[3]);
    atomicAdd(ch.status + 4, s.cnt[4]);
    atomicAdd(ch.status + 5, s.cnt[5]);
    atomicAdd(ch.status + 6, s.cnt[6]);
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
#define HYG_K2_MIN_CTAS 3   // resident CTAs per SM the register budget is held to (K2 is latency- and barrier-bound: more chains per SM fill idle issue slots; measured 3.23 / 2.91 / 2.95 us per site and SM at 2 / 3 / 4)
#endif
#ifndef HYG_EMU
template <int RT, bool PE>
__global__ void __launch_bounds__(HYG_NT, PE ? 1 : HYG_K2_MIN_CTAS) sg_filter_kernel(const SgModelDev* mdl, const SgChainDev* chains, SgRunDev run) {
  sg_filter_entry<RT, PE>(mdl, chains, run);
}
// The same recursion compiled for ONE CTA per SM (no register cap): used when there are no more units than SMs
// (whole-chain execution of a few chains, the C ABI's and the CLI's default), where per-site latency is all that counts.
template <int RT>
__global__ void __launch_bounds__(HYG_NT, 1) sg_filter_kernel_sparse(const SgModelDev* mdl, const SgChainDev* chains, SgRunDev run) {
  sg_filter_entry<RT, false>(mdl, chains, run);
}
#endif

}  // namespace hyg
#endif
