// hygeia_b200/csrc/hyg_dmp.cuh -- K6: per-site statistics of the aggregated backward trajectories (DMP calling).
//
// Reference (paths relative to /root/reference/src/two_group):
//   aggregate_results.py:125-147,181   trajectories of all seeds concatenated along the particle axis;
//                                      split_probs = mean(merge_states == 0, axis = 1)
//   get_dmps.py:63-76                  null statistic 1 - sum(control_regime != case_regime) / num_particles,
//                                      per regime pair (i, j): 1 - sum((control == i) * (case == j)) / num_particles
//   get_dmps.py:119-126                regime frequencies np.bincount(row, minlength = n_regimes) / row.shape[0]
// Input: three int8 matrices [T][P] (site-major, P = seeds x backward trajectories), exactly the matrices the reference
// writes as merge_states_/control_regimes_/case_regimes_chrom_*.csv.gz.  Output: fp64 per site.  All counts are exact
// integers and every output is ONE fp64 division (and one subtraction), in the reference's order -- bit-identical results.
//
// B200 mapping: a streaming byte kernel, bound by HBM (3 P bytes in, 16 + 16 R bytes out per site).  Persistent CTAs take
// tiles of HYG_DMP_TILE consecutive sites; the tile's rows are contiguous in memory, so the three [tile][P] byte blocks are staged in
// shared memory by TMA bulk copies (cp.async.bulk + mbarrier, one issuing thread); then HYG_DMP_LANES threads per site walk its rows
// with 4-byte shared-memory loads, transpose 32 particles at a time into bit planes and count every regime with one LOP3 + POPC.
#ifndef HYG_DMP_CUH
#define HYG_DMP_CUH

#include <stdint.h>

#ifndef HYG_EMU
#include <cuda_runtime.h>
#endif

#ifndef HYG_DMP_TILE
#define HYG_DMP_TILE 32       // sites per tile (a multiple of 16, so that tiles start on 16-byte boundaries)
#endif
#ifndef HYG_DMP_LANES
#define HYG_DMP_LANES 2       // threads per site (each walks every 2nd word of the site's rows); a power of two <= 8
// (tile, lanes) measured on B200 at P = 200: (128, 4) 5.2 TB/s, (64, 4) 5.8, (32, 4) 6.2, (64, 2) 6.3, (32, 2) 6.5 -- small CTAs
// (64 threads, 19 KB) keep ~10 tiles in flight per SM, which is what hides the TMA latency
#endif
#define HYG_DMP_NT (HYG_DMP_TILE * HYG_DMP_LANES)

namespace hyg {

struct DmpArgs {
  unsigned long long T;
  unsigned int P;           // particles per site
  unsigned int R;           // regimes (<= 8)
  const signed char* merged;   // [T][P], 0 = split, 1 = merged
  const signed char* control;  // [T][P] regime of the control group
  const signed char* cse;      // [T][P] regime of the case group
  double* split_prob;       // [T]      mean(merged == 0)
  double* null_stat;        // [T]      1 - #(control != case) / P
  double* control_freq;     // [T][R]   or null
  double* case_freq;        // [T][R]   or null
  double* pair_stat;        // [T][R][R] 1 - #(control == i and case == j) / P, or null
  unsigned long long n_tiles;
};

// Bit planes of up to 32 particles: bit j of plane k = bit k of the regime of particle j.  A regime count is then ONE
// three-input logic op (LOP3) + POPC per 32 particles.
struct DmpPlanes {
  unsigned c0, c1, c2, k0, k1, k2, m;
  __device__ __forceinline__ void clear() { c0 = c1 = c2 = k0 = k1 = k2 = m = 0u; }
  // bit k of every byte of w, moved to bit j of that byte: one funnel shift + one AND (the OR into the plane fuses with it)
  template <int k, int j> __device__ __forceinline__ static unsigned plane(unsigned w) {
    return ((j >= k) ? (w << (j >= k ? j - k : 0)) : (w >> (k > j ? k - j : 0))) & (0x01010101u << j);
  }
  // add the four particles of one word of each matrix at bit positions j, 8 + j, 16 + j, 24 + j (j is a compile-time constant)
  template <int j> __device__ __forceinline__ void add(unsigned mw, unsigned cw, unsigned kw) {
    m |= plane<0, j>(mw);
    c0 |= plane<0, j>(cw); c1 |= plane<1, j>(cw); c2 |= plane<2, j>(cw);
    k0 |= plane<0, j>(kw); k1 |= plane<1, j>(kw); k2 |= plane<2, j>(kw);
  }
  template <int r> __device__ __forceinline__ static int count(unsigned p0, unsigned p1, unsigned p2) {
    return __popc(((r & 1) ? p0 : ~p0) & ((r & 2) ? p1 : ~p1) & ((r & 4) ? p2 : ~p2));
  }
  // counts of regimes 0..7 (positions that hold no particle count as regime 0: the caller subtracts them)
  __device__ __forceinline__ void flush(int (&cc)[8], int (&kc)[8], int& msum, int& ne) {
    cc[0] += count<0>(c0, c1, c2); cc[1] += count<1>(c0, c1, c2); cc[2] += count<2>(c0, c1, c2); cc[3] += count<3>(c0, c1, c2);
    cc[4] += count<4>(c0, c1, c2); cc[5] += count<5>(c0, c1, c2); cc[6] += count<6>(c0, c1, c2); cc[7] += count<7>(c0, c1, c2);
    kc[0] += count<0>(k0, k1, k2); kc[1] += count<1>(k0, k1, k2); kc[2] += count<2>(k0, k1, k2); kc[3] += count<3>(k0, k1, k2);
    kc[4] += count<4>(k0, k1, k2); kc[5] += count<5>(k0, k1, k2); kc[6] += count<6>(k0, k1, k2); kc[7] += count<7>(k0, k1, k2);
    msum += __popc(m);
    ne += __popc((c0 ^ k0) | (c1 ^ k1) | (c2 ^ k2));
    clear();
  }
};

// Stage `bytes` from global to shared memory (dst, src 16-byte aligned).  Device: one thread issues TMA bulk copies
// (cp.async.bulk) that complete on an mbarrier every thread then waits on; `phase` is the barrier's parity for this use.
__device__ __forceinline__ void dmp_stage3(unsigned char* d0, unsigned char* d1, unsigned char* d2, const signed char* s0, const signed char* s1,
                                           const signed char* s2, size_t bytes, unsigned long long* mbar, unsigned phase) {
  const size_t vec = bytes / 16 * 16;
#ifdef HYG_EMU
  (void)mbar; (void)phase;
  for (size_t i = threadIdx.x; i < vec; i += blockDim.x) { d0[i] = static_cast<unsigned char>(s0[i]); d1[i] = static_cast<unsigned char>(s1[i]); d2[i] = static_cast<unsigned char>(s2[i]); }
#else
  const uint32_t bar = static_cast<uint32_t>(__cvta_generic_to_shared(mbar));
  if (threadIdx.x == 0 && vec > 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(static_cast<uint32_t>(3 * vec)) : "memory");
    const uint32_t CH = 32768;
#pragma unroll
    for (int a = 0; a < 3; a++) {
      unsigned char* d = (a == 0) ? d0 : ((a == 1) ? d1 : d2);
      const signed char* g = (a == 0) ? s0 : ((a == 1) ? s1 : s2);
      for (size_t off = 0; off < vec; off += CH) {
        const uint32_t n = static_cast<uint32_t>((vec - off < CH) ? vec - off : CH);
        const uint32_t da = static_cast<uint32_t>(__cvta_generic_to_shared(d + off));
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(da), "l"(g + off), "r"(n),
                     "r"(bar)
                     : "memory");
      }
    }
  }
#endif
  // the last tile may end inside a 16-byte vector
  for (size_t i = vec + threadIdx.x; i < bytes; i += blockDim.x) {
    d0[i] = static_cast<unsigned char>(s0[i]); d1[i] = static_cast<unsigned char>(s1[i]); d2[i] = static_cast<unsigned char>(s2[i]);
  }
#ifndef HYG_EMU
  if (vec > 0) {
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(bar), "r"(phase)
          : "memory");
    }
  }
#endif
  __syncthreads();
}

__device__ __forceinline__ void dmp_site_stats_entry(const DmpArgs& a, unsigned char* sm) {
  const int tid = threadIdx.x;
  const int site = tid / HYG_DMP_LANES, q = tid % HYG_DMP_LANES;
  const unsigned P = a.P;
  const size_t tile_bytes = static_cast<size_t>(HYG_DMP_TILE) * P;
  const size_t arr_pitch = (tile_bytes + 15) / 16 * 16;
  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(sm);
  unsigned char* sm_m = sm + 16;
  unsigned char* sm_c = sm_m + arr_pitch;
  unsigned char* sm_k = sm_c + arr_pitch;
#ifndef HYG_EMU
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(mbar))));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
#endif
  // quot[i] = i / P for i = 0..P: every output is one of these IEEE quotients (the reference divides counts by P)
  double* quot = reinterpret_cast<double*>(sm_k + arr_pitch);
  for (unsigned i = tid; i <= P; i += HYG_DMP_NT) quot[i] = static_cast<double>(i) / static_cast<double>(P);
  __syncthreads();
  unsigned phase = 0;
  for (unsigned long long tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const unsigned long long t0 = tile * HYG_DMP_TILE;
    const unsigned nt = static_cast<unsigned>((a.T - t0 < HYG_DMP_TILE) ? (a.T - t0) : HYG_DMP_TILE);
    const size_t bytes = static_cast<size_t>(nt) * P;
    const size_t goff = static_cast<size_t>(t0) * P;   // multiple of 16: HYG_DMP_TILE is
    dmp_stage3(sm_m, sm_c, sm_k, a.merged + goff, a.control + goff, a.cse + goff, bytes, mbar, phase);
    if (bytes >= 16) phase ^= 1u;
    int cnt_c[8], cnt_k[8];
#pragma unroll
    for (int r = 0; r < 8; r++) { cnt_c[r] = 0; cnt_k[r] = 0; }
    int msum = 0, ne = 0;
    const bool active = static_cast<unsigned>(site) < nt;
    if (active) {
      const unsigned char* rm = sm_m + static_cast<size_t>(site) * P;
      const unsigned char* rc = sm_c + static_cast<size_t>(site) * P;
      const unsigned char* rk = sm_k + static_cast<size_t>(site) * P;
      if ((P & 3u) == 0u) {   // rows are 4-byte aligned: one shared-memory word = four particles
        const unsigned nw = P / 4;
        DmpPlanes pl;
        pl.clear();
        const unsigned* wm = reinterpret_cast<const unsigned*>(rm);
        const unsigned* wc = reinterpret_cast<const unsigned*>(rc);
        const unsigned* wk = reinterpret_cast<const unsigned*>(rk);
        int mine = 0, groups = 0;
        // this lane's words are q, q + 4, q + 8, ...; eight of them (32 particles) fill the planes
#define HYG_DMP_ADD(J)                                                       \
  {                                                                          \
    const unsigned w = w0 + (J) * HYG_DMP_LANES;                             \
    if (w < nw) { pl.add<J>(wm[w], wc[w], wk[w]); mine += 4; }               \
  }
        for (unsigned w0 = q; w0 < nw; w0 += 8 * HYG_DMP_LANES) {
          HYG_DMP_ADD(0) HYG_DMP_ADD(1) HYG_DMP_ADD(2) HYG_DMP_ADD(3) HYG_DMP_ADD(4) HYG_DMP_ADD(5) HYG_DMP_ADD(6) HYG_DMP_ADD(7)
          pl.flush(cnt_c, cnt_k, msum, ne);
          groups++;
        }
#undef HYG_DMP_ADD
        // bit positions that held no particle were counted as regime 0 (each flush covers 32 positions)
        const int pad = groups * 32 - mine;
        cnt_c[0] -= pad; cnt_k[0] -= pad;
      } else {
        // rows start at any byte offset: every 4-particle word of a row is cut out of two aligned shared-memory words with one
        // funnel shift; the row's last word keeps only its P mod 4 valid bytes (the dropped ones count as padding below)
        const unsigned off = static_cast<unsigned>(site) * P;
        const unsigned sh = 8u * (off & 3u), base = off >> 2;
        const unsigned nw = (P + 3u) / 4u, tail = P - 4u * (nw - 1u);
        const unsigned tmask = (tail == 4u) ? 0xFFFFFFFFu : ((1u << (8u * tail)) - 1u);
        const unsigned* wm = reinterpret_cast<const unsigned*>(sm_m) + base;
        const unsigned* wc = reinterpret_cast<const unsigned*>(sm_c) + base;
        const unsigned* wk = reinterpret_cast<const unsigned*>(sm_k) + base;
        DmpPlanes pl;
        pl.clear();
        int mine = 0, groups = 0;
#define HYG_DMP_ADDU(J)                                                                        \
  {                                                                                            \
    const unsigned w = w0 + (J) * HYG_DMP_LANES;                                               \
    if (w < nw) {                                                                              \
      unsigned mw = __funnelshift_r(wm[w], wm[w + 1], sh), cw = __funnelshift_r(wc[w], wc[w + 1], sh),   \
               kw = __funnelshift_r(wk[w], wk[w + 1], sh);                                    \
      const bool lastw = (w == nw - 1u);                                                       \
      if (lastw) { mw &= tmask; cw &= tmask; kw &= tmask; }                                    \
      pl.add<J>(mw, cw, kw);                                                                   \
      mine += lastw ? static_cast<int>(tail) : 4;                                              \
    }                                                                                          \
  }
        for (unsigned w0 = q; w0 < nw; w0 += 8 * HYG_DMP_LANES) {
          HYG_DMP_ADDU(0) HYG_DMP_ADDU(1) HYG_DMP_ADDU(2) HYG_DMP_ADDU(3) HYG_DMP_ADDU(4) HYG_DMP_ADDU(5) HYG_DMP_ADDU(6) HYG_DMP_ADDU(7)
          pl.flush(cnt_c, cnt_k, msum, ne);
          groups++;
        }
#undef HYG_DMP_ADDU
        const int pad = groups * 32 - mine;
        cnt_c[0] -= pad; cnt_k[0] -= pad;
      }
    }
    // ---- combine the HYG_DMP_LANES partial counts of a site (adjacent lanes of one warp) ----
#pragma unroll
    for (int o = 1; o < HYG_DMP_LANES; o <<= 1) {
#pragma unroll
      for (int r = 0; r < 8; r++) {
        cnt_c[r] += __shfl_xor_sync(0xffffffffu, cnt_c[r], o);
        cnt_k[r] += __shfl_xor_sync(0xffffffffu, cnt_k[r], o);
      }
      msum += __shfl_xor_sync(0xffffffffu, msum, o);
      ne += __shfl_xor_sync(0xffffffffu, ne, o);
    }
    if (active) {
      const unsigned long long t = t0 + site;
      const double dP = static_cast<double>(P);
      if (q == 0) {
        a.split_prob[t] = quot[static_cast<int>(P) - msum];       // np.mean(merged == 0, axis = 1)
        a.null_stat[t] = 1.0 - quot[ne];                           // 1. - np.sum(control != case) / P
      }
      // lane q writes regimes q and q + 4 of both frequency rows: np.bincount(row) / row.shape[0]
#pragma unroll
      for (int r = 0; r < 8; r++) {
        if ((r % HYG_DMP_LANES) == q && static_cast<unsigned>(r) < a.R) {
          if (a.control_freq) a.control_freq[t * a.R + r] = quot[cnt_c[r]];
          if (a.case_freq) a.case_freq[t * a.R + r] = quot[cnt_k[r]];
        }
      }
      if (a.pair_stat && q == 0) {   // --test_regime_combinations: R x R table per site, built in local memory
        const unsigned char* rc = sm_c + static_cast<size_t>(site) * P;
        const unsigned char* rk = sm_k + static_cast<size_t>(site) * P;
        int pair[64];
        for (int i = 0; i < 64; i++) pair[i] = 0;
        for (unsigned p = 0; p < P; p++) pair[(rc[p] & 7u) * 8u + (rk[p] & 7u)]++;
        for (unsigned i = 0; i < a.R; i++)
          for (unsigned jj = 0; jj < a.R; jj++) a.pair_stat[(t * a.R + i) * a.R + jj] = 1.0 - static_cast<double>(pair[i * 8 + jj]) / dP;
      }
    }
    __syncthreads();   // every row has been read before the next tile's copies land
  }
}

#ifndef HYG_EMU
extern __shared__ __align__(16) unsigned char hyg_dmp_smem[];
__global__ void __launch_bounds__(HYG_DMP_NT) dmp_site_stats_kernel(DmpArgs a) { dmp_site_stats_entry(a, hyg_dmp_smem); }

// ---- FDR procedures (multiple_testing.py) : element-wise pieces; the sort, scan and count kernels follow ----
// Qs[i] = 1 / (i + 1) * cumsum[i]   (multiple_testing.py:5-6: 1./np.linspace(1, n, n) * np.cumsum(sorted))
__global__ void dmp_running_mean_kernel(const double* cumsum, double* qs, unsigned long long n) {
  for (unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<unsigned long long>(gridDim.x) * blockDim.x)
    qs[i] = __dmul_rn(__ddiv_rn(1.0, static_cast<double>(i + 1)), cumsum[i]);
}
// ranking and excessive error rate of the weighted procedure (multiple_testing.py:14-17); no FMA contraction
__global__ void dmp_weighted_rank_kernel(const double* t, const double* wfp, const double* wfn, double alpha, double* ranking, double* excess,
                                         unsigned long long* index, unsigned long long n) {
  for (unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<unsigned long long>(gridDim.x) * blockDim.x) {
    const double d = __dsub_rn(t[i], alpha);
    const double num = __dmul_rn(wfp[i], d);
    const double den = __dadd_rn(__dmul_rn(wfn[i], __dsub_rn(1.0, t[i])), __dmul_rn(wfp[i], fabs(d)));
    ranking[i] = __ddiv_rn(num, den);
    excess[i] = num;
    index[i] = i;
  }
}

// ---- device sort / scan / count for the two procedures (no library calls) -----------------------------------------------------
// Stable LSD radix sort of (64-bit key, 64-bit payload) records, 8 bits per pass: per pass a histogram kernel (256 bins per tile),
// one single-block exclusive scan over the bin-major (bin, tile) counts, and a scatter kernel that walks its tile IN ORDER, 256
// records at a time, ranking equal digits with __match_any_sync inside a warp and a per-warp count table across warps -- so equal
// keys keep their input order (np.sort / a stable argsort with ties by index).
#define HYG_RS_NT 256
#define HYG_RS_WARPS (HYG_RS_NT / 32)

__device__ __forceinline__ unsigned long long dmp_key_of_double(double v) {
  const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(v));
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);   // ascending order of the doubles (-0 < +0; NaNs last)
}
__device__ __forceinline__ double dmp_double_of_key(unsigned long long k) {
  const unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double(static_cast<long long>(b));
}
__global__ void dmp_make_keys_kernel(const double* v, unsigned long long* keys, unsigned long long n) {
  for (unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<unsigned long long>(gridDim.x) * blockDim.x)
    keys[i] = dmp_key_of_double(v[i]);
}
__global__ void dmp_keys_to_doubles_kernel(const unsigned long long* keys, double* v, unsigned long long n) {
  for (unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<unsigned long long>(gridDim.x) * blockDim.x)
    v[i] = dmp_double_of_key(keys[i]);
}
// counts[bin * n_tiles + tile]
__global__ void __launch_bounds__(HYG_RS_NT) dmp_radix_hist_kernel(const unsigned long long* keys, unsigned long long n, unsigned long long tile, int shift,
                                                                   unsigned int* counts, unsigned int n_tiles) {
  __shared__ unsigned int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const unsigned long long lo = blockIdx.x * tile, hi = (lo + tile < n) ? lo + tile : n;
  for (unsigned long long i = lo + threadIdx.x; i < hi; i += HYG_RS_NT) atomicAdd(&h[(keys[i] >> shift) & 0xFFull], 1u);
  __syncthreads();
  counts[static_cast<size_t>(threadIdx.x) * n_tiles + blockIdx.x] = h[threadIdx.x];
}
// exclusive scan of m unsigned counts into 64-bit offsets, one block: per-thread contiguous chunks + a block scan of the chunk sums
__global__ void __launch_bounds__(1024) dmp_scan_counts_kernel(const unsigned int* counts, unsigned long long* offsets, unsigned int m) {
  __shared__ unsigned long long part[1024];
  const unsigned int per = (m + 1023u) / 1024u;
  const unsigned int lo = threadIdx.x * per, hi = (lo + per < m) ? lo + per : m;
  unsigned long long sum = 0;
  for (unsigned int i = lo; i < hi; i++) sum += counts[i];
  part[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long acc = 0;
    for (int i = 0; i < 1024; i++) { const unsigned long long v = part[i]; part[i] = acc; acc += v; }
  }
  __syncthreads();
  unsigned long long acc = part[threadIdx.x];
  for (unsigned int i = lo; i < hi; i++) { offsets[i] = acc; acc += counts[i]; }
}
template <bool PAY>
__global__ void __launch_bounds__(HYG_RS_NT) dmp_radix_scatter_kernel(const unsigned long long* keys, const unsigned long long* pay, unsigned long long* keys_out,
                                                                      unsigned long long* pay_out, unsigned long long n, unsigned long long tile, int shift,
                                                                      const unsigned long long* offsets, unsigned int n_tiles) {
  __shared__ unsigned long long base[256];                 // next free output position per digit for this tile
  __shared__ unsigned int wcnt[HYG_RS_WARPS][256];         // records of each digit in each warp of the current chunk
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  base[tid] = offsets[static_cast<size_t>(tid) * n_tiles + blockIdx.x];
#pragma unroll
  for (int w = 0; w < HYG_RS_WARPS; w++) wcnt[w][tid] = 0;
  __syncthreads();
  const unsigned long long lo = blockIdx.x * tile, hi = (lo + tile < n) ? lo + tile : n;
  for (unsigned long long c0 = lo; c0 < hi; c0 += HYG_RS_NT) {
    const unsigned long long i = c0 + tid;
    const bool live = i < hi;
    const unsigned long long k = live ? keys[i] : 0ull;
    const unsigned int d = live ? static_cast<unsigned int>((k >> shift) & 0xFFull) : 0x100u;   // 0x100: no record
    const unsigned int peers = __match_any_sync(0xFFFFFFFFu, d);
    const unsigned int before = __popc(peers & ((1u << lane) - 1u));
    if (live && before == 0) wcnt[warp][d] = __popc(peers);
    __syncthreads();
    if (live) {
      unsigned long long pos = base[d] + before;
#pragma unroll
      for (int w = 0; w < HYG_RS_WARPS; w++) pos += (w < warp) ? wcnt[w][d] : 0u;
      keys_out[pos] = k;
      if (PAY) pay_out[pos] = pay[i];
    }
    __syncthreads();
    {   // thread = digit: advance the base and clear the table for the next chunk
      unsigned int tot = 0;
#pragma unroll
      for (int w = 0; w < HYG_RS_WARPS; w++) { tot += wcnt[w][tid]; wcnt[w][tid] = 0; }
      base[tid] += tot;
    }
    __syncthreads();
  }
}

// inclusive scan of doubles: (1) tile sums, (2) one-block exclusive scan of the tile sums, (3) in-tile scan + offset.  A tile is
// walked in chunks of 256 x 4 elements: thread-sequential over its 4, shuffle scan over the warp, warp totals through shared memory.
#define HYG_SCAN_ITEMS 4
__device__ __forceinline__ double dmp_block_excl(double v, double* sh, double& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const double t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) sh[warp] = inc;
  __syncthreads();
  double off = 0.0, tot = 0.0;
#pragma unroll
  for (int w = 0; w < HYG_RS_WARPS; w++) { const double x = sh[w]; if (w < warp) off += x; tot += x; }
  __syncthreads();
  total = tot;
  return off + inc - v;
}
__global__ void __launch_bounds__(HYG_RS_NT) dmp_tile_sum_kernel(const double* x, unsigned long long n, unsigned long long tile, double* sums) {
  __shared__ double sh[HYG_RS_WARPS];
  const unsigned long long lo = blockIdx.x * tile, hi = (lo + tile < n) ? lo + tile : n;
  double carry = 0.0;
  for (unsigned long long c0 = lo; c0 < hi; c0 += HYG_RS_NT * HYG_SCAN_ITEMS) {
    double v = 0.0;
#pragma unroll
    for (int j = 0; j < HYG_SCAN_ITEMS; j++) { const unsigned long long i = c0 + threadIdx.x * HYG_SCAN_ITEMS + j; if (i < hi) v += x[i]; }
    double tot;
    dmp_block_excl(v, sh, tot);
    carry += tot;
  }
  if (threadIdx.x == 0) sums[blockIdx.x] = carry;
}
__global__ void dmp_scan_tile_sums_kernel(double* sums, unsigned int m) {   // one thread: m <= a few hundred
  if (blockIdx.x == 0 && threadIdx.x == 0) { double acc = 0.0; for (unsigned int i = 0; i < m; i++) { const double v = sums[i]; sums[i] = acc; acc += v; } }
}
__global__ void __launch_bounds__(HYG_RS_NT) dmp_tile_scan_kernel(const double* x, double* out, unsigned long long n, unsigned long long tile, const double* offs) {
  __shared__ double sh[HYG_RS_WARPS];
  const unsigned long long lo = blockIdx.x * tile, hi = (lo + tile < n) ? lo + tile : n;
  double carry = offs[blockIdx.x];
  for (unsigned long long c0 = lo; c0 < hi; c0 += HYG_RS_NT * HYG_SCAN_ITEMS) {
    double e[HYG_SCAN_ITEMS], v = 0.0;
#pragma unroll
    for (int j = 0; j < HYG_SCAN_ITEMS; j++) { const unsigned long long i = c0 + threadIdx.x * HYG_SCAN_ITEMS + j; e[j] = (i < hi) ? x[i] : 0.0; v += e[j]; }
    double tot;
    double run = carry + dmp_block_excl(v, sh, tot);
#pragma unroll
    for (int j = 0; j < HYG_SCAN_ITEMS; j++) { const unsigned long long i = c0 + threadIdx.x * HYG_SCAN_ITEMS + j; run += e[j]; if (i < hi) out[i] = run; }
    carry += tot;
  }
}
// #{i : x[i] <= bound}
__global__ void dmp_count_le_kernel(const double* x, unsigned long long n, double bound, unsigned long long* count) {
  unsigned long long c = 0;
  for (unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<unsigned long long>(gridDim.x) * blockDim.x)
    c += (x[i] <= bound) ? 1ull : 0ull;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}
// gather a double array through the sorted payload (index) array
__global__ void dmp_gather_kernel(const double* src, const unsigned long long* idx, double* dst, unsigned long long n) {
  for (unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<unsigned long long>(gridDim.x) * blockDim.x)
    dst[i] = src[idx[i]];
}
#endif

}  // namespace hyg
#endif
