// hygeia_b200/csrc/hyg_tg.cuh -- K4/K5: the two-group (case/control) particle filter and backward simulation.
//
// Reference (/root/reference/src/two_group, TensorFlow 2.3 / TFP 0.11 -- not runnable here; DESIGN.md section 2 describes
// the restatement this kernel is checked against and the documented differences):
//   filter step                 hygeia/filter_and_smoother_algorithm.py:176-288
//   first step / padding        hygeia/filter_and_smoother_algorithm.py:141-172,334-365
//   backward simulation         hygeia/filter_and_smoother_algorithm.py:368-446, hygeia/smoothing_functions.py:46-59
//   transition densities        hygeia/case_control_regime_model.py:80-193, hygeia/case_control_distributions.py:138-151,246-291
//   the 2R + R^2 proposals      hygeia/case_control_proposal_mappings.py:11-134,175-207
//   optimal finite-state + systematic resampling   hygeia/resampling_functions.py:7-69
//   emission                    hygeia/case_control_regime_model.py:197-231  (= two K1 tables: control and case)
//
// B200 design: one 512-thread CTA per chain (chromosome x batch x seed); the <= 2400 particles of a site live in shared
// memory; the emission term is the sum of two T x R tables produced by K1 (the reference evaluates 2400 x S beta-binomial
// densities per site); only the FINITE-weight particles (typically 300-600) are compacted and sorted (shared-memory
// bitonic network sized to the next power of two) for the optimal finite-state selection of the <= 50 ancestors.
// The reference keeps the whole filter history (T x 2400 weights and states, 16 GB+ of RAM per task) for the backward
// pass; here each site stores only its <= 50 ancestors (~1.4 KB) in HBM and the backward pass RECOMPUTES the children.
// All arithmetic fp64; draws are Philox(seed, chain, site).
#ifndef HYG_TG_CUH
#define HYG_TG_CUH

#include "hyg_common.cuh"

// Threads per CTA x resident CTAs per SM, measured on B200 (us per site and chain slot -> site-chains/s per GPU):
//   round 1, one CTA per SM: 256 -> 173, 512 -> 133, 768 -> 153, 1024 -> 168 us
//   round 2: 512 x 1: 25.7 us, 5.8e6/s;  256 x 2: 35.9 us per chain but two chains per SM, 8.2e6/s  <- the kernel is latency-bound
//   (IPC 1.3 with one CTA per SM), so a second resident chain fills the idle issue slots
#ifndef HYG_TG_NT
#define HYG_TG_NT 256
#endif
#define HYG_TG_NW (HYG_TG_NT / 32)
#define HYG_TG_NPMAX 2432     // >= M * (2R + R^2) for M = 50, R = 6, multiple of 256... (2400)
#define HYG_TG_MMAX 64        // max resampled ancestors
#define HYG_TG_BMAX 32        // max backward trajectories
#ifndef HYG_TG_SORTMAX
#define HYG_TG_SORTMAX 2048   // records the shared-memory sort arrays hold (two CTAs of 92 KB fit an SM); more -> global scratch
#endif
#ifndef HYG_TG_CTAS
#define HYG_TG_CTAS 2         // resident CTAs per SM the kernel is compiled and launched for
#endif
#define HYG_TG_EMAX (HYG_TG_SORTMAX > HYG_TG_NPMAX ? HYG_TG_SORTMAX : HYG_TG_NPMAX)
#define HYG_TG_BIGMAX 4096    // the global scratch of a CTA holds this many (key, value, index) records: >= the next power of two of NPMAX
#define HYG_TG_BINS 256       // 1-nat bins of the normalised log-weight for the pre-selection of the sort
#define HYG_TG_SELMAX 512     // pre-selection: at most this many of the heaviest particles are sorted on the first attempt
#define HYG_TAG_FILTER 0x46494C54u
#define HYG_TAG_BACKWARD 0x42414B57u
#define HYG_TAG_PHANTOM 0x5048414Eu

namespace hyg {

struct TgModelDev {
  int R, u, M, B;
  uint32_t dmax;                 // hazard tables hold d = 0..dmax per regime (row pitch dmax + 1)
  double logP[HYG_RMAX][HYG_RMAX];   // control regime transitions (log, -inf diagonal)
  double logPm[2][2];            // merged-indicator transitions (log), rows = previous
  // hazards as host-built logs (libm, like the oracle): {log rho(d, r), log(1 - rho(d, r))}, [R][dmax+1]
  const double2* lrho_c;
  const double2* lrho_k;
  double nl_rm1, nl_rm2;         // -log(R - 1), -log(R - 2)
  int presel[2];                 // the sort covers the heaviest particles only: at least this many on the 1st / 2nd attempt
  int big_from;                  // more selected particles than this: the sort runs in the CTA's global scratch (<= HYG_TG_SORTMAX)
};

struct TgState {   // one particle
  int m, dc, rc, dk, rk;
};

struct TgAncRec {  // per-ancestor record kept for the backward pass (HBM)
  int dc, dk;
  unsigned char m, rc, rk, pad;
  double w_prev;     // unnormalised log-weight of the ancestor
  double logW_prev;  // normalised log-weight
};
struct TgStepRec {   // per-site header
  int n_anc;         // M'
  int mode;          // 0: F <= M (plain), 1: optimal, 2: unbiased (log c infinite)
  double log_c;
  double lse;        // logsumexp of the previous unnormalised weights
};

struct TgChainDev {
  unsigned long long T;
  const double* lo_c;   // T x R
  const double* lo_k;   // T x R
  unsigned long long seed;
  uint32_t chain;
  // outputs (device)
  int* traj;            // T x B x 5 : m, dc, rc, dk, rk
  double* log_norm;     // 1
  int* taps;            // optional T x 4 : n_particles, K, n_finite, sort attempts (0: no sort, 1..3)
};

struct TgRunDev {
  unsigned char* ws;               // per-CTA workspace: T_max step headers + T_max * anc_pitch ancestor records
  unsigned long long ws_stride;    // bytes per CTA
  unsigned long long t_max;
  unsigned long long anc_pitch;    // ancestor records per site (>= M and >= the R^2 particles of the first site)
  unsigned long long scratch_off;  // byte offset of the CTA's sort scratch inside its workspace (HYG_TG_BIGMAX x 18 bytes)
  unsigned int* queue;
  int n_chains;
};

__host__ __device__ __forceinline__ double tg_uniform(uint64_t seed, uint32_t chain, uint32_t tag, uint64_t index) {
  uint32_t c[4] = {static_cast<uint32_t>(index), static_cast<uint32_t>(index >> 32), chain, tag};
  philox4x32_10(c, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  const uint64_t hi = c[0] >> 5, lo = c[1] >> 6;
  return (static_cast<double>(hi) * 67108864.0 + static_cast<double>(lo)) * (1.0 / 9007199254740992.0);
}

// log transition density of `n` given `p` (case_control_regime_model.py:97-193; case_control_distributions.py:138-151,246-291)
__device__ __forceinline__ double tg_log_trans(const TgModelDev& md, const TgState& p, const TgState& n, bool step0) {
  const int u = md.u;
  // merged indicator
  double lm;
  if (step0) lm = (n.m == 1) ? 0.0 : -HYG_INF;
  else if ((p.dk < p.dc ? p.dk : p.dc) >= u) lm = md.logPm[p.m][n.m];
  else lm = (n.m == p.m) ? 0.0 : -HYG_INF;
  if (lm == -HYG_INF) return -HYG_INF;
  // control
  const uint32_t pitch = md.dmax + 1;
  const double2 lr_c = step0 ? make_double2(0.0, -HYG_INF) : md.lrho_c[p.rc * pitch + (static_cast<uint32_t>(p.dc) < md.dmax ? p.dc : md.dmax)];
  double lc;
  if (n.dc == 1) lc = lr_c.x + md.logP[p.rc][n.rc];
  else lc = (n.dc == p.dc + 1 && n.rc == p.rc) ? lr_c.y : -HYG_INF;
  if (!(lc > -HYG_INF)) return -HYG_INF;
  // case, first matching rule
  double lk;
  if (n.m == 1) {
    lk = (n.rc == n.rk && n.dc == n.dk) ? 0.0 : -HYG_INF;
  } else if (p.m == 1 && n.dc != 1) {
    lk = (n.rk != n.rc && n.dk == 1) ? md.nl_rm1 : -HYG_INF;
  } else {
    const bool allowed = (n.rk != n.rc) && (n.rk != p.rk);
    const double unif = allowed ? ((n.rc == p.rk) ? md.nl_rm1 : md.nl_rm2) : -HYG_INF;
    if (n.rc == p.rk && p.m == 0) {
      lk = (n.dk == 1) ? unif : -HYG_INF;
    } else {
      const double2 lr_k = step0 ? make_double2(0.0, -HYG_INF) : md.lrho_k[p.rk * pitch + (static_cast<uint32_t>(p.dk) < md.dmax ? p.dk : md.dmax)];
      if (n.dk == 1) lk = lr_k.x + unif;
      else lk = (n.dk == p.dk + 1 && n.rk == p.rk) ? lr_k.y : -HYG_INF;
    }
  }
  const double out = lm + lc + lk;
  return (out == out) ? out : -HYG_INF;
}

// proposal q (0 .. 2R + R^2 - 1) of ancestor a (case_control_proposal_mappings.py:11-134)
// (x * ceil(2^20 / d)) >> 20 == x / d for 0 <= x < 4096, 1 <= d <= 64 (the error x (ceil - exact) / 2^20 < 1/256 < 1/d)
__device__ __forceinline__ int tg_div_magic(int d) { return ((1 << 20) + d - 1) / d; }
__device__ __forceinline__ int tg_div(int x, int magic) { return (x * magic) >> 20; }

__device__ __forceinline__ TgState tg_propose(int R, int magicR, const TgState& a, int q) {
  TgState n;
  if (q == 0) {
    n.m = a.m; n.dc = a.dc + 1; n.rc = a.rc; n.dk = a.dk + 1; n.rk = a.rk;
  } else if (q < R) {                       // control jumps, skipping the case regime
    n.m = 0; n.dc = 1; n.rc = (q <= a.rk) ? q - 1 : q; n.dk = a.dk + 1; n.rk = a.rk;
  } else if (q < 2 * R - 1) {               // case jumps, skipping the control regime
    n.m = 0; n.dc = a.dc + 1; n.rc = a.rc; n.dk = 1; n.rk = (q < R + a.rc) ? q - R : q - R + 1;
  } else if (q == 2 * R - 1) {              // merge (durations 0 = impossible state if already merged)
    const int dm = (a.m == 0) ? a.dc + 1 : 0;
    n.m = 1; n.dc = dm; n.rc = a.rc; n.dk = dm; n.rk = a.rc;
  } else {                                  // both change: control regime i, case regime j
    const int ij = q - 2 * R, i = tg_div(ij, magicR), j = ij - i * R;
    n.m = (i == j) ? 1 : 0; n.dc = 1; n.rc = i; n.dk = 1; n.rk = j;
  }
  return n;
}

struct TgSmem {
  // particles of the current site
  double w[HYG_TG_NPMAX];
  int dc[HYG_TG_NPMAX], dk[HYG_TG_NPMAX];
  unsigned char m[HYG_TG_NPMAX], rc[HYG_TG_NPMAX], rk[HYG_TG_NPMAX];
  // sorting / resampling scratch
  unsigned long long key[HYG_TG_SORTMAX];   // exact order-preserving image of the normalised log-weight
  unsigned short sidx[HYG_TG_SORTMAX];      // particle index travelling with the key
  double e[HYG_TG_EMAX];           // exp(normalised log-weight) by particle, then in sorted order, then cumulative sums
  // ancestors
  TgState anc[HYG_TG_MMAX];
  double anc_w[HYG_TG_MMAX], anc_logW[HYG_TG_MMAX];
  int parents[HYG_TG_MMAX];
  // backward trajectories
  TgState nxt[HYG_TG_BMAX];
  int pick[HYG_TG_BMAX];
  int rep[HYG_TG_BMAX];
  int trj[HYG_TG_BMAX * 5];      // backward pass: the site's sampled states, staged for one coalesced write
  int glist[HYG_TG_BMAX];       // backward pass: the first trajectory of every group of equal next states
  double ub[HYG_TG_BMAX];      // backward pass: the site's uniform of every trajectory (one Philox evaluation each)
  double ctot[HYG_TG_NPMAX / 32 + 2];    // per-chunk totals / carries of the chunked cumulative sums (32 particles per chunk)
  double ccar[HYG_TG_NPMAX / 32 + 2];
  double lcn_tab[HYG_TG_MMAX + 1];       // log(M - a) - log(rcs[a]) for every candidate a of the K loop        // backward pass: first trajectory with the same next state (its predecessor law is reused)
  double red[2][HYG_TG_NW][4];
  int ired[2][HYG_TG_NW];
  double bc[8];
  int ibc[12];
  int hist[HYG_TG_BINS];          // finite particles by floor(-normalised log-weight): where the heaviest few hundred end
  TgModelDev mdl;
};

__device__ __forceinline__ double tg_block_sum(double v, TgSmem& s, int& flip) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) s.red[flip][warp][0] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < HYG_TG_NW; w++) t += s.red[flip][w][0];
  flip ^= 1;
  return t;
}
__device__ __forceinline__ double tg_block_max(double v, TgSmem& s, int& flip) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_max(v);
  if (lane == 0) s.red[flip][warp][0] = v;
  __syncthreads();
  // max is exact whatever the order: one load per lane and a shuffle tree instead of HYG_TG_NW loads per thread
  double t = (lane < HYG_TG_NW) ? s.red[flip][lane][0] : -HYG_INF;
  t = warp_max(t);
  flip ^= 1;
  return t;
}
// block-wide exclusive prefix sum of one int per thread; returns the exclusive prefix, total in `total`
__device__ __forceinline__ int tg_block_excl_scan(int v, int& total, TgSmem& s, int& flip) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(HYG_FULL, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) s.ired[flip][warp] = inc;
  __syncthreads();
  int off = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < HYG_TG_NW; w++) { const int x = s.ired[flip][w]; if (w < warp) off += x; tot += x; }
  flip ^= 1;
  total = tot;
  return off + inc - v;
}

// in-place bitonic sort of (s.key, s.sidx)[0..n), n a power of two, all threads of the CTA: descending in the key, ties by
// ascending particle index (= a stable descending sort of the weights in particle order)
template <class KP, class SP>
__device__ __forceinline__ void tg_sort_desc(KP key, SP sidx, int n) {
  const int half = n >> 1;
  // only the warps that hold pairs take part (n = 256: four of sixteen) and meet at their own named barrier; the others wait
  // at the block barrier below
  const int nthr = (half < HYG_TG_NT) ? ((half + 31) & ~31) : HYG_TG_NT;
  if (static_cast<int>(threadIdx.x) < nthr) {
    for (int k = 2; k <= n; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        // one compare-exchange per (thread, iteration): pair q -> elements i (bit log2(j) cleared) and i | j, so no lane idles
        const int lj = __ffs(j) - 1;
        for (int q = threadIdx.x; q < half; q += HYG_TG_NT) {
          const int i = ((q >> lj) << (lj + 1)) | (q & (j - 1));
          const int l = i | j;
          const unsigned long long a = key[i], b = key[l];
          const unsigned short ia = sidx[i], ib = sidx[l];
          const bool desc = ((i & k) == 0);
          const bool a_after_b = (a < b) || (a == b && ia > ib);   // a belongs after b in the final order
          if (a_after_b == desc) { key[i] = b; key[l] = a; sidx[i] = ib; sidx[l] = ia; }
        }
        // The 32 pairs q of a warp (q = tid + NT m) with a stride <= 32 lie in ONE block of 64 consecutive elements, the same
        // block for every such stride: between two such stages a warp barrier is enough.
        const int next = (j > 1) ? (j >> 1) : k;   // stride of the following stage
        if (j >= 64 || next >= 64) named_barrier(2, nthr);
        else __syncwarp();
      }
    }
  }
  __syncthreads();
}

// One attempt of OptimalFiniteState (resampling_functions.py:7-52) on the particles of the histogram bins <= selbin: compaction,
// sort, suffix sums, the K search, the systematic resampling of the residual.  Returns 1 when a tooth of the comb fell behind
// the sorted prefix (the caller repeats with more particles), else 0 with s.parents / K / log_c / mode set.
// BIG = false: the sort's arrays are the shared-memory ones (<= HYG_TG_SORTMAX particles); BIG = true: more particles than
// that were selected (only possible when > 2048 of a site's 2400 particles are finite AND a tooth reached the lightest of them),
// and the same code runs on a global scratch area of the CTA.
template <bool BIG>
__device__ __noinline__ int tg_ofs_attempt(TgSmem& s, const TgChainDev& ch, unsigned char* scratch, int attempt, int selbin, int n_sel, int F, int M,
                                           unsigned long long t, double lse, double mx, double inv_se, int n_part, int c0, int c1, int& flip,
                                           int& K, double& log_c, int& mode) {
  const TgModelDev& md = s.mdl;
  (void)md;
  const int tid = threadIdx.x, lane = tid & 31;
  unsigned long long* key = BIG ? reinterpret_cast<unsigned long long*>(scratch) : s.key;
  double* ev = BIG ? reinterpret_cast<double*>(scratch + sizeof(unsigned long long) * HYG_TG_BIGMAX) : s.e;
  unsigned short* sidx = BIG ? reinterpret_cast<unsigned short*>(scratch + (sizeof(unsigned long long) + sizeof(double)) * HYG_TG_BIGMAX) : s.sidx;
  if (tid == 0) { s.ibc[4] = 0; s.ibc[5] = 0; }
  __syncthreads();
  // the selected particles' keys (any order: the sort's order is total, ties by particle index) and the others' total weight
  double rest = 0.0;
  for (int c = c0; c < c1; c++) {
    const double v = s.w[c];
    if (v > -HYG_INF) {
      const double dn = lse - v;
      const int bin = (dn < static_cast<double>(HYG_TG_BINS - 1)) ? static_cast<int>(dn < 0.0 ? 0.0 : dn) : HYG_TG_BINS - 1;
      if (bin <= selbin) {
        // order-preserving key of the normalised log-weight; the particle index travels beside it (stable ties)
        unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(v - lse));
        b = (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
        const int pos = atomicAdd(&s.ibc[4], 1);
        key[pos] = b;
        sidx[pos] = static_cast<unsigned short>(c);
      } else {
        rest += (attempt == 0) ? s.e[c] : exp(v - mx);   // exp(v - mx) was kept by the pass above; a repeated attempt finds s.e reused
      }
    }
  }
  int n_sort = 64;
  while (n_sort < n_sel) n_sort <<= 1;
  for (int i = n_sel + tid; i < n_sort; i += HYG_TG_NT) { key[i] = 0ull; sidx[i] = 0xFFFFu; }
  rest = (n_sel < F) ? tg_block_sum(rest, s, flip) * inv_se : 0.0;
  __syncthreads();
#ifndef HYG_TG_SKIP_SORT
  tg_sort_desc(key, sidx, n_sort);
#endif
  // e[p] = exp(sorted normalised log-weight); reverse cumulative sums rcs[p] = sum_{i >= p} e[i]
  double* slw = reinterpret_cast<double*>(key);   // sorted normalised log-weights (the keys are not needed after the sort)
  for (int p = tid; p < n_sort; p += HYG_TG_NT) {
    double v = 0.0, lwn = -HYG_INF;
    if (p < n_sel) { lwn = s.w[sidx[p]] - lse; v = exp(lwn); }
    ev[p] = v;
    slw[p] = lwn;      // overwrites key[p]: the sort ended with a block barrier and nothing reads the keys after it
  }
  __syncthreads();
  // Reverse cumulative sums in chunks of 32, walked from the end: suf[p] = (shuffle-tree suffix inside the chunk) + carry,
  // carry = the unsorted particles' total, then the later chunks' totals added one after the other.  The trees are
  // independent, so every warp takes the chunks c = warp, warp + NW, ...; only the carry additions stay sequential.
  // The sums are only needed at positions <= M (K < M): anc_w doubles as rcs[0..M].
  const int n_chunks = (n_sel + 31) / 32;
  for (int cidx = (tid >> 5); cidx < n_chunks; cidx += HYG_TG_NW) {
    const int p = cidx * 32 + lane;
    double inc = (p < n_sel) ? ev[p] : 0.0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double tt = __shfl_down_sync(HYG_FULL, inc, o);
      if (lane + o < 32) inc += tt;
    }
    if (lane == 0) s.ctot[cidx] = inc;
  }
  __syncthreads();
  if (tid < 32) {
    double carry = rest;
    for (int cidx = n_chunks - 1; cidx >= 0; cidx--) {
      if (cidx * 32 <= M) {   // a chunk that holds positions <= M: its in-chunk suffix sums are needed too
        const int p = cidx * 32 + lane;
        double inc = (p < n_sel) ? ev[p] : 0.0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const double tt = __shfl_down_sync(HYG_FULL, inc, o);
          if (lane + o < 32) inc += tt;
        }
        if (p <= M && p < n_sel) s.anc_w[p] = inc + carry;
      }
      carry += s.ctot[cidx];
    }
    __syncwarp();   // anc_w[0..M] written above is read by every lane below
    // fixed point with the reference's loop structure: (a, b, log_c) <- (k_new, a, log_c(a)) while a != b, a < n, a < M
    // the loop below needs log(M - a) - log(rcs[a]) at a data-dependent sequence of a < M: all M values are taken now, by
    // the lanes in parallel, so that no logarithm sits on the loop's dependency chain
    for (int a0 = lane; a0 < M && a0 < n_sel; a0 += 32) s.lcn_tab[a0] = log(static_cast<double>(M - a0)) - log(s.anc_w[a0]);
    __syncwarp();
    int a = 0, b = -1;
    double lc = -1.0;
    while (a != b && a < F && a < M) {
      const double lcn = s.lcn_tab[a];
      int cnt = 0;
      // counts beyond M do not change the outcome; a + 96 <= M + 95 < n_sel whenever n_sel < F
      // three independent loads, then three ballots (slw is -inf beyond n_sel, up to n_sort >= 64; beyond that: no particle)
      const int p0 = a + lane, p1 = p0 + 32, p2 = p0 + 64;
      const double x0 = (p0 < n_sel) ? slw[p0] : -HYG_INF, x1 = (p1 < n_sel) ? slw[p1] : -HYG_INF, x2 = (p2 < n_sel) ? slw[p2] : -HYG_INF;
      cnt = __popc(__ballot_sync(HYG_FULL, lcn + x0 > 0.0)) + __popc(__ballot_sync(HYG_FULL, lcn + x1 > 0.0)) +
            __popc(__ballot_sync(HYG_FULL, lcn + x2 > 0.0));
      b = a; a = a + cnt; lc = lcn;
    }
    int Kf = b;
    if (!(Kf < F)) { Kf = F; lc = -HYG_INF; }
    if (lane == 0) { s.ibc[0] = Kf; s.bc[0] = lc; }
  }
  __syncthreads();
  K = s.ibc[0];
  log_c = s.bc[0];
  if (log_c - log_c != 0.0) {
    // log c infinite: multinomial ancestors by inverse CDF over the particle-order weights, unbiased weights
    mode = 2;
    // cumulative sums over the particles in index order (s.e reused)
    for (int c = tid; c < n_part; c += HYG_TG_NT) s.e[c] = (s.w[c] > -HYG_INF) ? exp(s.w[c] - lse) : 0.0;
    __syncthreads();
    if (tid == 0) { double acc = 0.0; for (int c = 0; c < n_part; c++) { acc += s.e[c]; s.e[c] = acc; } }
    __syncthreads();
    for (int a = tid; a < M; a += HYG_TG_NT) {
      const double uu = tg_uniform(ch.seed, ch.chain, HYG_TAG_FILTER, (static_cast<uint64_t>(a + 1) << 32) + t) * s.e[n_part - 1];
      int lo = 0, hi = n_part - 1;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (s.e[mid] < uu) lo = mid + 1; else hi = mid; }
      s.parents[a] = lo;
    }
    log_c = 0.0;
    __syncthreads();
    return 0;
  }
  mode = 1;
  const int L = M - K;
  // kept particles
  for (int a = tid; a < K; a += HYG_TG_NT) s.parents[a] = static_cast<int>(sidx[a]);
  // residual: cumulative sums of e[K..n_sel) in sorted order, in chunks of 32 from K: shuffle-tree prefix inside a chunk (all
  // warps, chunk c = warp, warp + NW, ...) + the earlier chunks' totals added one after the other (one thread)
  const int r_chunks = (L > 0) ? (n_sel - K + 31) / 32 : 0;
  for (int cidx = (tid >> 5); cidx < r_chunks; cidx += HYG_TG_NW) {
    const int p = K + cidx * 32 + lane;
    double inc = (p < n_sel) ? ev[p] : 0.0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double tt = __shfl_up_sync(HYG_FULL, inc, o);
      if (lane >= o) inc += tt;
    }
    if (p < n_sel) ev[p] = inc;
    if (lane == 31) s.ctot[cidx] = inc;
  }
  __syncthreads();
  if (tid == 0) {
    double carry = 0.0;
    for (int cidx = 0; cidx < r_chunks; cidx++) { s.ccar[cidx] = carry; carry += s.ctot[cidx]; }
  }
  if (tid == 32) s.ub[0] = tg_uniform(ch.seed, ch.chain, HYG_TAG_FILTER, t);   // the site's resampling uniform, evaluated once
  __syncthreads();
  for (int p = K + tid; p < n_sel && L > 0; p += HYG_TG_NT) ev[p] = ev[p] + s.ccar[(p - K) >> 5];
  __syncthreads();
  if (L > 0) {
    const double tot = ev[n_sel - 1] + rest;   // the whole residual, sorted or not
    const double u = s.ub[0];
    for (int j = tid; j < L; j += HYG_TG_NT) {
      // first residual position i with T_j <= Q_i, T_j = (j + u) / L  (resampling_functions.py:56-69); 0 if none
      const double Tj = (static_cast<double>(j) + u) / static_cast<double>(L) * tot;
      int lo = K, hi = n_sel;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (ev[mid] < Tj) lo = mid + 1; else hi = mid; }
      if (lo >= n_sel && n_sel < F) s.ibc[5] = 1;   // the tooth lies among the unsorted particles: sort them all
      const int ps = (lo < n_sel) ? lo : K;
      s.parents[K + j] = static_cast<int>(sidx[ps]);
    }
  }
  __syncthreads();
  return s.ibc[5] ? 1 : 0;
}

// ---- one chain: filter over all sites, then backward simulation ----
__device__ void tg_chain(const TgChainDev& ch, const TgRunDev& run, TgSmem& s, unsigned char* ws) {
  const TgModelDev& md = s.mdl;
  const int tid = threadIdx.x, lane = tid & 31;
  const int R = md.R, M = md.M, B = md.B, I = 2 * R + R * R;
  const int magicR = tg_div_magic(R);
  const unsigned long long T = ch.T;
  int flip = 0;
  TgStepRec* steps = reinterpret_cast<TgStepRec*>(ws);
  TgAncRec* ancs = reinterpret_cast<TgAncRec*>(ws + sizeof(TgStepRec) * run.t_max);

  // ---- first site (filter_and_smoother_algorithm.py:141-172; case_control_regime_model.py:233-244) ----
  int n_part = R * R;
  {
    const int r_ph = static_cast<int>(tg_uniform(ch.seed, ch.chain, HYG_TAG_PHANTOM, 0) * R);
    TgState ph; ph.m = 1; ph.dc = 0; ph.rc = r_ph; ph.dk = 0; ph.rk = r_ph;
    for (int c = tid; c < n_part; c += HYG_TG_NT) {
      TgState n; n.rc = c / R; n.rk = c % R; n.m = (n.rc == n.rk) ? 1 : 0; n.dc = 1; n.dk = 1;
      const double lt = tg_log_trans(md, ph, n, true);
      s.w[c] = (lt > -HYG_INF) ? lt + ch.lo_c[n.rc] + ch.lo_k[n.rk] : -HYG_INF;
      s.m[c] = n.m; s.dc[c] = 1; s.rc[c] = n.rc; s.dk[c] = 1; s.rk[c] = n.rk;
    }
    if (tid == 0) { steps[0].n_anc = 0; steps[0].mode = 0; steps[0].log_c = 0.0; steps[0].lse = 0.0; }
    __syncthreads();
    if (ch.taps && tid == 0) {
      int nf = 0;
      for (int c = 0; c < n_part; c++) nf += (s.w[c] > -HYG_INF);
      ch.taps[0] = n_part; ch.taps[1] = -1; ch.taps[2] = nf; ch.taps[3] = 0;
    }
  }

  // ---- filter (filter_and_smoother_algorithm.py:176-288) ----
  for (unsigned long long t = 1; t < T; t++) {
    // log-sum-exp of the previous weights and the number of finite ones
    double mx = -HYG_INF;
    int nf = 0;
    for (int c = tid; c < n_part; c += HYG_TG_NT) { const double v = s.w[c]; mx = v > mx ? v : mx; nf += (v > -HYG_INF); }
    mx = tg_block_max(mx, s, flip);
    double se = 0.0;
    for (int c = tid; c < n_part; c += HYG_TG_NT) { const double v = s.w[c]; const double ev = (v > -HYG_INF) ? exp(v - mx) : 0.0; s.e[c] = ev; se += ev; }
    se = tg_block_sum(se, s, flip);
    const double lse = mx + log(se);
    const double inv_se = 1.0 / se;
    // ---- which particles take part in the sort ----
    // Optimal finite-state resampling needs the finite particles in descending order of weight: the K < M heaviest keep their
    // weight, the rest is resampled systematically IN THAT ORDER.  Of the typically 1000-1800 finite particles of a site only
    // the heaviest 60-250 can be reached by one of the <= M teeth of the systematic comb (the lighter ones together weigh less
    // than the gap behind the last tooth), and the K loop never looks beyond position M + 96.  So: histogram of the finite
    // particles over 1-nat bins of the normalised log-weight, the smallest whole-bin prefix holding >= M + 110 particles is
    // sorted (<= HYG_TG_SELMAX, else everything), the others enter only through their total weight `rest`.  If a tooth does
    // fall behind the sorted prefix (measured: 0.2 % of the sites of one-sample data, never with 8 samples) the site is
    // redone with every finite particle sorted -- the results are those of the full sort in every case.
    for (int b = tid; b < HYG_TG_BINS; b += HYG_TG_NT) s.hist[b] = 0;
    __syncthreads();
    const int per = (n_part + HYG_TG_NT - 1) / HYG_TG_NT;   // contiguous chunk per thread keeps index order
    const int c0 = tid * per, c1 = (c0 + per < n_part) ? c0 + per : n_part;
    for (int c = c0; c < c1; c++) {
      const double v = s.w[c];
      if (v > -HYG_INF) {
        const double dn = lse - v;
        atomicAdd(&s.hist[(dn < static_cast<double>(HYG_TG_BINS - 1)) ? static_cast<int>(dn < 0.0 ? 0.0 : dn) : HYG_TG_BINS - 1], 1);
      }
    }
    __syncthreads();
    if (tid < 32) {   // prefix counts over the bins, 8 bins per lane
      int loc[HYG_TG_BINS / 32], run_ = 0;
#pragma unroll
      for (int i = 0; i < HYG_TG_BINS / 32; i++) { run_ += s.hist[lane * (HYG_TG_BINS / 32) + i]; loc[i] = run_; }
      int inc = run_;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int tt = __shfl_up_sync(HYG_FULL, inc, d);
        if (lane >= d) inc += tt;
      }
      const int excl = inc - run_;
      const int total = __shfl_sync(HYG_FULL, inc, 31);
      // first bin whose inclusive prefix reaches the wanted count, for the first and the second attempt
#pragma unroll
      for (int att = 0; att < 2; att++) {
        const int want = md.presel[att];
        int mybin = HYG_TG_BINS, mycnt = 0;
#pragma unroll
        for (int i = HYG_TG_BINS / 32 - 1; i >= 0; i--)
          if (excl + loc[i] >= want) { mybin = lane * (HYG_TG_BINS / 32) + i; mycnt = excl + loc[i]; }
        const unsigned has = __ballot_sync(HYG_FULL, mybin < HYG_TG_BINS);
        int selbin = HYG_TG_BINS - 1, selcnt = total;
        if (has) {
          const int src = __ffs(static_cast<int>(has)) - 1;
          selbin = __shfl_sync(HYG_FULL, mybin, src);
          selcnt = __shfl_sync(HYG_FULL, mycnt, src);
        }
        if (selcnt > HYG_TG_SELMAX) { selbin = HYG_TG_BINS - 1; selcnt = total; }
        if (lane == 0) { s.ibc[2 + 4 * att] = selbin; s.ibc[3 + 4 * att] = selcnt; }
      }
      if (lane == 0) s.ibc[1] = total;
    }
    __syncthreads();
    const int F = s.ibc[1];
    int Mp, mode = 0, K = -1, n_attempts = 0;
    double log_c = 0.0;
    if (F <= M) {
      // all finite particles become ancestors, in particle order
      int nfc = 0;
      for (int c = c0; c < c1; c++) nfc += (s.w[c] > -HYG_INF);
      int Fs;
      int pos = tg_block_excl_scan(nfc, Fs, s, flip);
      for (int c = c0; c < c1; c++)
        if (s.w[c] > -HYG_INF) s.parents[pos++] = c;
      Mp = F;
      __syncthreads();
    } else {
      // ---- OptimalFiniteState (resampling_functions.py:7-52) over the finite particles ----
      Mp = M;
      for (int attempt = 0; attempt < 3; attempt++) {
        const int selbin = (attempt == 2) ? HYG_TG_BINS - 1 : s.ibc[2 + 4 * attempt];
        const int n_sel = (attempt == 2) ? F : s.ibc[3 + 4 * attempt];
        n_attempts = attempt + 1;
        if (attempt == 1 && n_sel == s.ibc[3]) continue;   // the second selection is the first one: go straight to the full sort
        {
          const bool big = n_sel > md.big_from;
          const int redo = big ? tg_ofs_attempt<true>(s, ch, ws + run.scratch_off, attempt, selbin, n_sel, F, M, t, lse, mx, inv_se, n_part, c0, c1, flip, K, log_c, mode)
                               : tg_ofs_attempt<false>(s, ch, ws + run.scratch_off, attempt, selbin, n_sel, F, M, t, lse, mx, inv_se, n_part, c0, c1, flip, K, log_c, mode);
          if (!redo) break;
        }
        __syncthreads();   // everyone has read the flag before the next attempt clears it
      }
    }
    // ---- gather the ancestors, record them for the backward pass ----
    for (int a = tid; a < Mp; a += HYG_TG_NT) {
      const int c = s.parents[a];
      TgState st; st.m = s.m[c]; st.dc = s.dc[c]; st.rc = s.rc[c]; st.dk = s.dk[c]; st.rk = s.rk[c];
      s.anc[a] = st;
      s.anc_w[a] = s.w[c];
      s.anc_logW[a] = s.w[c] - lse;
      TgAncRec rec; rec.dc = st.dc; rec.dk = st.dk; rec.m = static_cast<unsigned char>(st.m); rec.rc = static_cast<unsigned char>(st.rc);
      rec.rk = static_cast<unsigned char>(st.rk); rec.pad = 0; rec.w_prev = s.anc_w[a]; rec.logW_prev = s.anc_logW[a];
      ancs[t * run.anc_pitch + a] = rec;
    }
    if (tid == 0) { steps[t].n_anc = Mp; steps[t].mode = mode; steps[t].log_c = log_c; steps[t].lse = lse; }
    __syncthreads();
    // ---- propose the 48 children of every ancestor and weight them (proposal-major order) ----
    const double* loc = ch.lo_c + t * R;
    const double* lok = ch.lo_k + t * R;
    const int n_new = I * Mp;
    const int magicMp = tg_div_magic(Mp);
    int nfin = 0;
    for (int c = tid; c < n_new; c += HYG_TG_NT) {
      const int q = tg_div(c, magicMp), a = c - q * Mp;
      const TgState pa = s.anc[a];
      const TgState n = tg_propose(R, magicR, pa, q);
      const double lt = tg_log_trans(md, pa, n, false);
      double wn = -HYG_INF;
      if (lt > -HYG_INF) {
        const double lg = lt + loc[n.rc] + lok[n.rk];
        if (mode == 0) wn = s.anc_w[a] + lg;
        else if (mode == 2) wn = -log(static_cast<double>(M)) + lse + lg;
        else { const double adj = log_c + s.anc_logW[a]; wn = s.anc_w[a] + lg - (adj < 0.0 ? adj : 0.0); }
      }
      nfin += (wn > -HYG_INF);
      s.w[c] = wn; s.m[c] = static_cast<unsigned char>(n.m); s.dc[c] = n.dc; s.rc[c] = static_cast<unsigned char>(n.rc);
      s.dk[c] = n.dk; s.rk[c] = static_cast<unsigned char>(n.rk);
    }
    n_part = n_new;
    __syncthreads();
    if (ch.taps) {
      const double tot = tg_block_sum(static_cast<double>(nfin), s, flip);
      if (tid == 0) { ch.taps[t * 4] = n_part; ch.taps[t * 4 + 1] = K; ch.taps[t * 4 + 2] = static_cast<int>(tot + 0.5); ch.taps[t * 4 + 3] = n_attempts; }
    }
  }

  // ---- log normalising constant ----
  {
    double mx = -HYG_INF;
    for (int c = tid; c < n_part; c += HYG_TG_NT) mx = s.w[c] > mx ? s.w[c] : mx;
    mx = tg_block_max(mx, s, flip);
    double se = 0.0;
    for (int c = tid; c < n_part; c += HYG_TG_NT) se += (s.w[c] > -HYG_INF) ? exp(s.w[c] - mx) : 0.0;
    se = tg_block_sum(se, s, flip);
    if (tid == 0 && ch.log_norm) *ch.log_norm = mx + log(se);
  }

  // ---- backward simulation (filter_and_smoother_algorithm.py:368-446) ----
  // the particles of site T-1 are still in shared memory; earlier sites are recomputed from their ancestor records
  for (long long t = static_cast<long long>(T) - 1; t >= 0; t--) {
    if (t < static_cast<long long>(T) - 1) {
      if (t == 0) {
        n_part = R * R;
        const int r_ph = static_cast<int>(tg_uniform(ch.seed, ch.chain, HYG_TAG_PHANTOM, 0) * R);
        TgState ph; ph.m = 1; ph.dc = 0; ph.rc = r_ph; ph.dk = 0; ph.rk = r_ph;
        for (int c = tid; c < n_part; c += HYG_TG_NT) {
          TgState n; n.rc = c / R; n.rk = c % R; n.m = (n.rc == n.rk) ? 1 : 0; n.dc = 1; n.dk = 1;
          const double lt = tg_log_trans(md, ph, n, true);
          s.w[c] = (lt > -HYG_INF) ? lt + ch.lo_c[n.rc] + ch.lo_k[n.rk] : -HYG_INF;
          s.m[c] = n.m; s.dc[c] = 1; s.rc[c] = n.rc; s.dk[c] = 1; s.rk[c] = n.rk;
        }
      } else {
        const TgStepRec hd = steps[t];
        const int Mp = hd.n_anc;
        for (int a = tid; a < Mp; a += HYG_TG_NT) {
          const TgAncRec rec = ancs[t * run.anc_pitch + a];
          TgState st; st.m = rec.m; st.dc = rec.dc; st.rc = rec.rc; st.dk = rec.dk; st.rk = rec.rk;
          s.anc[a] = st; s.anc_w[a] = rec.w_prev; s.anc_logW[a] = rec.logW_prev;
        }
        __syncthreads();
        const double* loc = ch.lo_c + t * R;
        const double* lok = ch.lo_k + t * R;
        n_part = I * Mp;
        const int magicMp = tg_div_magic(Mp);
        for (int c = tid; c < n_part; c += HYG_TG_NT) {
          const int q = tg_div(c, magicMp), a = c - q * Mp;
          const TgState pa = s.anc[a];
          const TgState n = tg_propose(R, magicR, pa, q);
          const double lt = tg_log_trans(md, pa, n, false);
          double wn = -HYG_INF;
          if (lt > -HYG_INF) {
            const double lg = lt + loc[n.rc] + lok[n.rk];
            if (hd.mode == 0) wn = s.anc_w[a] + lg;
            else if (hd.mode == 2) wn = -log(static_cast<double>(M)) + hd.lse + lg;
            else { const double adj = hd.log_c + s.anc_logW[a]; wn = s.anc_w[a] + lg - (adj < 0.0 ? adj : 0.0); }
          }
          s.w[c] = wn; s.m[c] = static_cast<unsigned char>(n.m); s.dc[c] = n.dc; s.rc[c] = static_cast<unsigned char>(n.rc);
          s.dk[c] = n.dk; s.rk[c] = static_cast<unsigned char>(n.rk);
        }
      }
      __syncthreads();
    }
    // one categorical draw per trajectory: logits_i = w_t[i] (+ log f(x_{t+1}^j | x_t^i)), inverse CDF in particle order
    const bool last = (t == static_cast<long long>(T) - 1);
    // The predecessor law depends on the trajectory only through its next state x_{t+1}^j, and backward trajectories coalesce:
    // it is evaluated once per DISTINCT next state (rep[j] = first trajectory with that state) and every trajectory of the
    // group draws from it with its own uniform -- the same numbers, 25 / #distinct times less work.
    if (tid < B) {
      int r = last ? 0 : tid;
      if (!last) {
        const TgState me = s.nxt[tid];
        for (int jj = 0; jj < tid; jj++) {
          const TgState o = s.nxt[jj];
          if (o.m == me.m && o.dc == me.dc && o.rc == me.rc && o.dk == me.dk && o.rk == me.rk) { r = jj; break; }
        }
      }
      s.rep[tid] = r;
      s.ub[tid] = tg_uniform(ch.seed, ch.chain, HYG_TAG_BACKWARD, (static_cast<uint64_t>(tid) << 32) + static_cast<uint64_t>(t));
    }
    if (tid < 32) {   // the distinct next states, in trajectory order (B <= 32: one warp)
      const bool is_rep = (tid < B) && (s.rep[tid] == tid);
      const unsigned gm = __ballot_sync(HYG_FULL, is_rep);
      if (is_rep) s.glist[__popc(gm & ((1u << tid) - 1u))] = tid;
      if (tid == 0) s.ibc[6] = __popc(gm);
    }
    __syncthreads();
    const int n_groups = s.ibc[6];
    double* thi = reinterpret_cast<double*>(s.key);   // per-thread inclusive [0, NT) and exclusive [NT, 2 NT) cumulative probability (the sort's key array is idle here)
    for (int gi = 0; gi < n_groups; gi++) {
      const int j = s.glist[gi];   // the group's first trajectory: its predecessor law serves the whole group
      TgState nx;
      if (!last) nx = s.nxt[j];
      double mx = -HYG_INF;
      for (int c = tid; c < n_part; c += HYG_TG_NT) {
        double lg = s.w[c];
        if (!last && lg > -HYG_INF) {
          TgState pc; pc.m = s.m[c]; pc.dc = s.dc[c]; pc.rc = s.rc[c]; pc.dk = s.dk[c]; pc.rk = s.rk[c];
          const double lt = tg_log_trans(md, pc, nx, false);
          lg = (lt > -HYG_INF) ? lg + lt : -HYG_INF;
        }
        s.e[c] = lg;
        mx = lg > mx ? lg : mx;
      }
      mx = tg_block_max(mx, s, flip);
      // cumulative probabilities in particle order: per-thread contiguous chunks + block scan of chunk totals
      const int per = (n_part + HYG_TG_NT - 1) / HYG_TG_NT;
      const int c0 = tid * per, c1 = (c0 + per < n_part) ? c0 + per : n_part;
      double loc_sum = 0.0;
      for (int c = c0; c < c1; c++) { const double p = (s.e[c] > -HYG_INF) ? exp(s.e[c] - mx) : 0.0; loc_sum += p; s.e[c] = loc_sum; }
      // exclusive scan of loc_sum over threads (warp scan + cross-warp)
      double inc = loc_sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double tt = __shfl_up_sync(HYG_FULL, inc, o);
        if (lane >= o) inc += tt;
      }
      if (lane == 31) s.red[flip][tid >> 5][1] = inc;
      __syncthreads();
      double off = 0.0, tot = 0.0;
#pragma unroll
      for (int w = 0; w < HYG_TG_NW; w++) { const double x = s.red[flip][w][1]; if (w < (tid >> 5)) off += x; tot += x; }
      flip ^= 1;
      const double excl = off + inc - loc_sum;
      thi[HYG_TG_NT + tid] = excl;
      thi[tid] = excl + loc_sum;
      __syncthreads();
      // every trajectory of the group draws with its own uniform: the first particle whose cumulative probability reaches the
      // target lies in the first thread chunk whose inclusive sum reaches it
      if (tid < B && s.rep[tid] == j) {
        const double target = s.ub[tid] * tot;
        int pickc = 0;
        if (tot > 0.0) {
          int lo = 0, hi = HYG_TG_NT - 1;
          while (lo < hi) { const int mid = (lo + hi) >> 1; if (thi[mid] < target) lo = mid + 1; else hi = mid; }
          const int k0 = lo * per, k1 = (k0 + per < n_part) ? k0 + per : n_part;
          const double ex = thi[HYG_TG_NT + lo];
          pickc = (k1 > k0) ? k1 - 1 : n_part - 1;
          for (int c = k0; c < k1; c++) if (ex + s.e[c] >= target) { pickc = c; break; }
        }
        s.pick[tid] = pickc;
      }
      __syncthreads();
    }
    // record the sampled states; they become x_{t+1} of the next (earlier) site
    for (int j = tid; j < B; j += HYG_TG_NT) {
      const int c = s.pick[j];
      TgState st; st.m = s.m[c]; st.dc = s.dc[c]; st.rc = s.rc[c]; st.dk = s.dk[c]; st.rk = s.rk[c];
      s.nxt[j] = st;
      int* o = s.trj + j * 5;
      o[0] = st.m; o[1] = st.dc; o[2] = st.rc; o[3] = st.dk; o[4] = st.rk;
    }
    __syncthreads();
    // the site's B x 5 integers leave as one run of consecutive words (they may go straight to pinned host memory over PCIe)
    for (int i = tid; i < B * 5; i += HYG_TG_NT) ch.traj[static_cast<size_t>(t) * B * 5 + i] = s.trj[i];
  }
}

#ifdef HYG_EMU
static unsigned char hyg_tg_smem_storage[sizeof(TgSmem) + 64];
#define HYG_TG_SMEM hyg_tg_smem_storage
#else
extern __shared__ __align__(16) unsigned char hyg_tg_smem_dyn[];
#define HYG_TG_SMEM hyg_tg_smem_dyn
#endif

__device__ __forceinline__ void tg_entry(const TgModelDev* mdl, const TgChainDev* chains, TgRunDev run) {
  TgSmem& s = *reinterpret_cast<TgSmem*>(HYG_TG_SMEM);
  __shared__ int s_next;
  if (threadIdx.x == 0) s.mdl = *mdl;
  __syncthreads();
  unsigned char* ws = run.ws + static_cast<size_t>(blockIdx.x) * run.ws_stride;
  for (;;) {
    if (threadIdx.x == 0) s_next = static_cast<int>(atomicAdd(run.queue, 1u));
    __syncthreads();
    const int c = s_next;
    __syncthreads();
    if (c >= run.n_chains) break;
    tg_chain(chains[c], run, s, ws);
  }
}

#ifndef HYG_EMU
__global__ void __launch_bounds__(HYG_TG_NT, HYG_TG_CTAS) tg_kernel(const TgModelDev* mdl, const TgChainDev* chains, TgRunDev run) {
  tg_entry(mdl, chains, run);
}
#endif

}  // namespace hyg
#endif
