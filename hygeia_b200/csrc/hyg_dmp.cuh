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
// tiles of 128 consecutive sites; the tile's rows are contiguous in memory, so the three [128][P] byte blocks are staged in
// shared memory with coalesced 16-byte loads; then one thread per site walks its rows with 4-byte shared-memory loads and
// counts with byte-sliced 64-bit accumulators (one 8-bit field per regime, flushed every 252 particles).
#ifndef HYG_DMP_CUH
#define HYG_DMP_CUH

#include <stdint.h>

#ifndef HYG_EMU
#include <cuda_runtime.h>
#endif

#define HYG_DMP_TILE 128   // sites per tile = threads per CTA
#define HYG_DMP_FLUSH 252  // particles between flushes of the 8-bit count fields (multiple of 4, <= 255)

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

__device__ __forceinline__ void dmp_count_particle(unsigned c, unsigned k, unsigned long long& cc, unsigned long long& kc, int& ne) {
  cc += 1ull << ((c & 7u) * 8u);
  kc += 1ull << ((k & 7u) * 8u);
  ne += (c != k) ? 1 : 0;
}

__device__ __forceinline__ void dmp_site_stats_entry(const DmpArgs& a, unsigned char* sm) {
  const int tid = threadIdx.x;
  const unsigned P = a.P;
  const size_t tile_bytes = static_cast<size_t>(HYG_DMP_TILE) * P;
  const size_t arr_pitch = (tile_bytes + 15) / 16 * 16;
  unsigned char* sm_m = sm;
  unsigned char* sm_c = sm + arr_pitch;
  unsigned char* sm_k = sm + 2 * arr_pitch;
  for (unsigned long long tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const unsigned long long t0 = tile * HYG_DMP_TILE;
    const unsigned nt = static_cast<unsigned>((a.T - t0 < HYG_DMP_TILE) ? (a.T - t0) : HYG_DMP_TILE);
    const size_t bytes = static_cast<size_t>(nt) * P;
    const size_t goff = static_cast<size_t>(t0) * P;   // multiple of 16: HYG_DMP_TILE = 8 x 16
    // ---- stage the three byte blocks (coalesced 16-byte loads; the tail of the last tile byte by byte) ----
#pragma unroll
    for (int arr = 0; arr < 3; arr++) {
      const signed char* g = (arr == 0 ? a.merged : (arr == 1 ? a.control : a.cse)) + goff;
      unsigned char* d = (arr == 0 ? sm_m : (arr == 1 ? sm_c : sm_k));
      const size_t nvec = bytes / 16;
      const uint4* gv = reinterpret_cast<const uint4*>(g);
      uint4* dv = reinterpret_cast<uint4*>(d);
      for (size_t i = tid; i < nvec; i += HYG_DMP_TILE) dv[i] = gv[i];
      for (size_t i = nvec * 16 + tid; i < bytes; i += HYG_DMP_TILE) d[i] = static_cast<unsigned char>(g[i]);
    }
    __syncthreads();
    if (static_cast<unsigned>(tid) < nt) {
      const unsigned char* rm = sm_m + static_cast<size_t>(tid) * P;
      const unsigned char* rc = sm_c + static_cast<size_t>(tid) * P;
      const unsigned char* rk = sm_k + static_cast<size_t>(tid) * P;
      int cnt_c[8], cnt_k[8], pair[64];
#pragma unroll
      for (int r = 0; r < 8; r++) { cnt_c[r] = 0; cnt_k[r] = 0; }
      int msum = 0, ne = 0;
      const bool want_pairs = a.pair_stat != nullptr;
      if (want_pairs)
        for (int i = 0; i < 64; i++) pair[i] = 0;
      for (unsigned p0 = 0; p0 < P; p0 += HYG_DMP_FLUSH) {
        const unsigned p1 = (p0 + HYG_DMP_FLUSH < P) ? p0 + HYG_DMP_FLUSH : P;
        unsigned long long cc = 0ull, kc = 0ull;
        unsigned p = p0;
        if ((P & 3u) == 0u) {   // rows are 4-byte aligned: one shared-memory word = four particles
          for (; p + 4 <= p1; p += 4) {
            const unsigned mw = *reinterpret_cast<const unsigned*>(rm + p);
            const unsigned cw = *reinterpret_cast<const unsigned*>(rc + p);
            const unsigned kw = *reinterpret_cast<const unsigned*>(rk + p);
            msum += static_cast<int>((mw & 0xFFu) + ((mw >> 8) & 0xFFu) + ((mw >> 16) & 0xFFu) + (mw >> 24));
            dmp_count_particle(cw & 0xFFu, kw & 0xFFu, cc, kc, ne);
            dmp_count_particle((cw >> 8) & 0xFFu, (kw >> 8) & 0xFFu, cc, kc, ne);
            dmp_count_particle((cw >> 16) & 0xFFu, (kw >> 16) & 0xFFu, cc, kc, ne);
            dmp_count_particle(cw >> 24, kw >> 24, cc, kc, ne);
          }
        }
        for (; p < p1; p++) {
          msum += rm[p];
          dmp_count_particle(rc[p], rk[p], cc, kc, ne);
        }
#pragma unroll
        for (int r = 0; r < 8; r++) {
          cnt_c[r] += static_cast<int>((cc >> (8 * r)) & 0xFFull);
          cnt_k[r] += static_cast<int>((kc >> (8 * r)) & 0xFFull);
        }
        if (want_pairs)
          for (unsigned q = p0; q < p1; q++) pair[(rc[q] & 7u) * 8u + (rk[q] & 7u)]++;
      }
      const unsigned long long t = t0 + tid;
      const double dP = static_cast<double>(P);
      a.split_prob[t] = static_cast<double>(static_cast<int>(P) - msum) / dP;       // np.mean(merged == 0, axis = 1)
      a.null_stat[t] = 1.0 - static_cast<double>(ne) / dP;                            // 1. - np.sum(control != case) / P
      if (a.control_freq) {
#pragma unroll
        for (unsigned r = 0; r < 8; r++)
          if (r < a.R) a.control_freq[t * a.R + r] = static_cast<double>(cnt_c[r]) / dP;   // np.bincount(row) / row.shape[0]
      }
      if (a.case_freq) {
#pragma unroll
        for (unsigned r = 0; r < 8; r++)
          if (r < a.R) a.case_freq[t * a.R + r] = static_cast<double>(cnt_k[r]) / dP;
      }
      if (want_pairs)
        for (unsigned i = 0; i < a.R; i++)
          for (unsigned j = 0; j < a.R; j++)
            a.pair_stat[(t * a.R + i) * a.R + j] = 1.0 - static_cast<double>(pair[i * 8 + j]) / dP;
    }
    __syncthreads();
  }
}

#ifndef HYG_EMU
extern __shared__ __align__(16) unsigned char hyg_dmp_smem[];
__global__ void __launch_bounds__(HYG_DMP_TILE) dmp_site_stats_kernel(DmpArgs a) { dmp_site_stats_entry(a, hyg_dmp_smem); }

// ---- FDR procedures (multiple_testing.py) : element-wise pieces; sorts and scans are thrust calls in hyg_api.cu ----
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
#endif

}  // namespace hyg
#endif
