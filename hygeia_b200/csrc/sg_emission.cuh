// hygeia_b200/csrc/sg_emission.cuh -- K1: fused per-CpG x per-sample emission log-likelihood.
//
// Computes logObs[t][r] = sum_{s=0..S-1} logBetaBinomial(x[s][t]; n[s][t], alpha_r, beta_r), the quantity
// Model::evaluateLogObservationDensity returns (/root/reference/src/single_group/src/cpp/singleGroup.h:610-627 on top
// of evaluateLogBetaBinomialDensity, misc/misc.h:630-640) -- once per (site, regime) instead of the reference's
// 1744 evaluations per site (algorithms/Smc.h:552-573).
//
// B200 design:
//   * counts are uint16, layout [S][pitch] with the SITE index fastest: a warp reads 32 x 4 B consecutive site pairs of
//     one sample (coalesced, vectorised); the sum over samples is a per-thread register accumulation in the order
//     s = 0..S-1, which is exactly the reference's summation order (no cross-lane reduction, no atomics);
//   * every lgamma argument is integer + constant, so the nine-term density is tabulated on the host with libm in the
//     reference's left-to-right order: a triangular (n, x) table whose row holds the R regime values (48 B for R = 6,
//     three 16-B shared-memory loads).  Rows for n <= nmax_smem live in shared memory (up to 227 KB per CTA, staged ONCE
//     per CTA by the TMA bulk-copy engine: cp.async.bulk + mbarrier), the rest of the table (n <= nmax_table) is read
//     through L2, and only counts beyond that evaluate lgamma on the device;
//   * ONE persistent launch covers every chromosome: the grid (= #SMs) strides over tiles of 1024 site pairs of a
//     flattened (data set, site) space, so the table is staged once per SM per sweep, not once per chromosome;
//   * the common case (all eight lookups of an unrolled batch hit the shared-memory rows) is branch-free, so the
//     24 shared-memory loads of a batch are issued back to back; a warp vote diverts rare batches to the general path.
// HBM traffic per site x sample is 4 B in + 48/S B out.  Results are bit-identical to a non-fast-math build of the
// reference for n <= nmax_table (same addends, same order).
#ifndef HYG_SG_EMISSION_CUH
#define HYG_SG_EMISSION_CUH

#include "hyg_common.cuh"

#ifndef HYG_EM_NT
#define HYG_EM_NT 1024
#endif
#define HYG_EM_TILE HYG_EM_NT      // site pairs per tile (= one pair per thread)
#define HYG_EM_SMEM_DOUBLES 29040  // 232320 B of table rows; + 16 B mbarrier slot <= the 227 KB opt-in limit of sm_100

namespace hyg {

struct SgEmissionSet {        // one data set (chromosome)
  unsigned long long T;       // sites
  unsigned long long pitch;   // elements per sample row (even, >= T, the pad is readable)
  const uint16_t* n_total;    // [S][pitch]
  const uint16_t* n_meth;     // [S][pitch]
  double* logobs;             // T x R
  unsigned long long tile0;   // first tile of this set in the flattened tile space
  uint32_t S;
  uint32_t pad_;
};

struct SgEmissionArgs {
  const SgEmissionSet* sets;  // device array
  uint32_t n_sets;
  unsigned long long n_tiles;
  const double* table;        // triangular [(n(n+1)/2 + x)][R] for n <= nmax_table (global)
  int nmax_table;
  int nmax_smem;              // rows with n <= nmax_smem are staged in shared memory
  double alpha[HYG_RMAX], beta[HYG_RMAX];
};

#ifdef HYG_EMU
static double hyg_em_smem_storage[HYG_EM_SMEM_DOUBLES + 2];
#define HYG_EM_SMEM hyg_em_smem_storage
#else
extern __shared__ __align__(128) double hyg_em_smem_dyn[];
#define HYG_EM_SMEM hyg_em_smem_dyn
#endif

// Stage `bytes` (multiple of 16) from global to shared memory.  On the device: TMA bulk copies tracked by an mbarrier
// (one elected thread issues, everybody waits on the phase).  `mbar` is an 8-byte shared slot.
__device__ __forceinline__ void stage_table(double* dst, const double* src, uint32_t bytes, unsigned long long* mbar) {
#ifdef HYG_EMU
  (void)mbar;
  for (uint32_t i = threadIdx.x; i < bytes / 8; i += blockDim.x) dst[i] = src[i];
  __syncthreads();
#else
  const uint32_t bar = static_cast<uint32_t>(__cvta_generic_to_shared(mbar));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    const uint32_t CH = 32768;
    for (uint32_t off = 0; off < bytes; off += CH) {
      const uint32_t n = (bytes - off < CH) ? bytes - off : CH;
      const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(reinterpret_cast<char*>(dst) + off));
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
                   "l"(reinterpret_cast<const char*>(src) + off), "r"(n), "r"(bar)
                   : "memory");
    }
  }
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar)
        : "memory");
  }
#endif
}

// misc.h:630-640 evaluated on the device (only for counts beyond the host table)
template <int R>
__device__ __noinline__ void emission_direct(const double* alpha, const double* beta, uint32_t x, uint32_t n, double* acc) {
  const double c0 = lgamma(static_cast<double>(n + 1)) - lgamma(static_cast<double>(x + 1)) - lgamma(static_cast<double>(n - x + 1));
  for (int r = 0; r < R; r++) {
    const double al = alpha[r], be = beta[r];
    const double v = c0 + lgamma(x + al) + lgamma(n - x + be) - lgamma(n + al + be) + lgamma(al + be) - lgamma(al) - lgamma(be);
    acc[r] += v;
  }
}

// general (branchy) path: any count
template <int R>
__device__ __forceinline__ void emission_add_general(const SgEmissionArgs& a, const double* stab, uint32_t rows_smem, uint32_t x, uint32_t n,
                                                     double (&acc)[R]) {
  if (x > n) {  // impossible count: density zero (misc.h:636-639)
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] += -HYG_INF;
    return;
  }
  const uint32_t row = n * (n + 1) / 2 + x;
  if (row < rows_smem) {
    const double* p = stab + static_cast<size_t>(row) * R;
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] += p[r];
  } else if (n <= static_cast<uint32_t>(a.nmax_table)) {
    const double* p = a.table + static_cast<size_t>(row) * R;
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] += __ldg(p + r);
  } else {
    double tmp[R];
#pragma unroll
    for (int r = 0; r < R; r++) tmp[r] = 0.0;
    emission_direct<R>(a.alpha, a.beta, x, n, tmp);
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] += tmp[r];
  }
}

template <int R>
__device__ __forceinline__ void sg_emission_entry(const SgEmissionArgs a) {
  double* stab = HYG_EM_SMEM;
  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(stab + HYG_EM_SMEM_DOUBLES);
  const uint32_t rows_smem = static_cast<uint32_t>(a.nmax_smem + 1) * (a.nmax_smem + 2) / 2;
  stage_table(stab, a.table, ((rows_smem * R * 8u) + 15u) & ~15u, mbar);
  const uint32_t nmax_s = static_cast<uint32_t>(a.nmax_smem);

  uint32_t set = 0;
  for (unsigned long long tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    while (set + 1 < a.n_sets && a.sets[set + 1].tile0 <= tile) set++;  // tiles are visited in increasing order
    const SgEmissionSet ds = a.sets[set];
    const unsigned long long n_pairs = (ds.T + 1) / 2;
    const unsigned long long i = (tile - ds.tile0) * HYG_EM_TILE + threadIdx.x;  // site pair handled by this thread
    const bool active = i < n_pairs;
    const uint32_t* nt32 = reinterpret_cast<const uint32_t*>(ds.n_total);
    const uint32_t* nm32 = reinterpret_cast<const uint32_t*>(ds.n_meth);
    const unsigned long long pitch32 = ds.pitch / 2;
    const unsigned long long ii = active ? i : 0;

    double acc0[R], acc1[R];
#pragma unroll
    for (int r = 0; r < R; r++) { acc0[r] = 0.0; acc1[r] = 0.0; }
    uint32_t s = 0;
    for (; s + 4 <= ds.S; s += 4) {  // four samples (eight lookups) in flight per thread
      uint32_t nn[4], xx[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        nn[k] = __ldg(nt32 + (s + k) * pitch32 + ii);
        xx[k] = __ldg(nm32 + (s + k) * pitch32 + ii);
      }
      bool fast = true;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint32_t n0 = nn[k] & 0xFFFFu, n1 = nn[k] >> 16, x0 = xx[k] & 0xFFFFu, x1 = xx[k] >> 16;
        fast = fast && (n0 <= nmax_s) && (n1 <= nmax_s) && (x0 <= n0) && (x1 <= n1);
      }
      if (__all_sync(HYG_FULL, fast)) {
        // branch-free: every lookup hits the shared-memory rows
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const uint32_t n0 = nn[k] & 0xFFFFu, n1 = nn[k] >> 16, x0 = xx[k] & 0xFFFFu, x1 = xx[k] >> 16;
          const double* p0 = stab + static_cast<size_t>(n0 * (n0 + 1) / 2 + x0) * R;
          const double* p1 = stab + static_cast<size_t>(n1 * (n1 + 1) / 2 + x1) * R;
          if (R % 2 == 0) {
            const double2* q0 = reinterpret_cast<const double2*>(p0);
            const double2* q1 = reinterpret_cast<const double2*>(p1);
#pragma unroll
            for (int r = 0; r < R / 2; r++) {
              const double2 v0 = q0[r], v1 = q1[r];
              acc0[2 * r] += v0.x; acc0[2 * r + 1] += v0.y;
              acc1[2 * r] += v1.x; acc1[2 * r + 1] += v1.y;
            }
          } else {
#pragma unroll
            for (int r = 0; r < R; r++) { acc0[r] += p0[r]; acc1[r] += p1[r]; }
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; k++) {
          emission_add_general<R>(a, stab, rows_smem, xx[k] & 0xFFFFu, nn[k] & 0xFFFFu, acc0);
          emission_add_general<R>(a, stab, rows_smem, xx[k] >> 16, nn[k] >> 16, acc1);
        }
      }
    }
    for (; s < ds.S; s++) {
      const uint32_t nn = __ldg(nt32 + s * pitch32 + ii), xx = __ldg(nm32 + s * pitch32 + ii);
      emission_add_general<R>(a, stab, rows_smem, xx & 0xFFFFu, nn & 0xFFFFu, acc0);
      emission_add_general<R>(a, stab, rows_smem, xx >> 16, nn >> 16, acc1);
    }
    if (active) {
      double* out = ds.logobs + 2 * i * R;
      const bool second = (2 * i + 1 < ds.T);
      if (R % 2 == 0) {
        double2* o2 = reinterpret_cast<double2*>(out);
#pragma unroll
        for (int r = 0; r < R / 2; r++) o2[r] = make_double2(acc0[2 * r], acc0[2 * r + 1]);
        if (second) {
#pragma unroll
          for (int r = 0; r < R / 2; r++) o2[R / 2 + r] = make_double2(acc1[2 * r], acc1[2 * r + 1]);
        }
      } else {
#pragma unroll
        for (int r = 0; r < R; r++) out[r] = acc0[r];
        if (second) {
#pragma unroll
          for (int r = 0; r < R; r++) out[R + r] = acc1[r];
        }
      }
    }
  }
}

#ifndef HYG_EMU
template <int R>
__global__ void __launch_bounds__(HYG_EM_NT, 1) sg_emission_kernel(const SgEmissionArgs a) {
  sg_emission_entry<R>(a);
}
#endif

}  // namespace hyg
#endif
