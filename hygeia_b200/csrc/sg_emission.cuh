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
//     three 16-B shared-memory loads).  Rows for n <= nmax_smem live in shared memory (up to 227 KB per CTA), the rest
//     of the table (n <= nmax_table) is read through L2, and only counts beyond that evaluate lgamma on the device;
//   * grid = #SMs (persistent, grid-stride over site pairs), 1024 threads per CTA so that one resident CTA per SM
//     still has 32 warps to cover HBM latency; HBM traffic per site x sample is 4 B in + 48/S B out.
// Results are bit-identical to a non-fast-math build of the reference for n <= nmax_table (same addends, same order).
#ifndef HYG_SG_EMISSION_CUH
#define HYG_SG_EMISSION_CUH

#include "hyg_common.cuh"

#define HYG_EM_NT 1024
#define HYG_EM_SMEM_DOUBLES 29056  // 232448 B = the 227 KB opt-in limit of sm_100

namespace hyg {

struct SgEmissionArgs {
  unsigned long long T;      // sites
  uint32_t S;                // samples
  unsigned long long pitch;  // elements per sample row (>= T rounded up to even; the pad is readable)
  const uint16_t* n_total;   // [S][pitch]
  const uint16_t* n_meth;    // [S][pitch]
  double* logobs;            // T x R
  const double* table;       // triangular [(n(n+1)/2 + x)][R] for n <= nmax_table (global)
  int nmax_table;
  int nmax_smem;             // rows with n <= nmax_smem are staged in shared memory
  double alpha[HYG_RMAX], beta[HYG_RMAX];
};

#ifdef HYG_EMU
static double hyg_em_smem_storage[HYG_EM_SMEM_DOUBLES];
#define HYG_EM_SMEM hyg_em_smem_storage
#else
extern __shared__ __align__(16) double hyg_em_smem_dyn[];
#define HYG_EM_SMEM hyg_em_smem_dyn
#endif

// misc.h:630-640 evaluated on the device (only for counts beyond the host table)
template <int R>
__device__ __noinline__ void emission_direct(const SgEmissionArgs& a, uint32_t x, uint32_t n, double (&acc)[R]) {
  const double c0 = lgamma(static_cast<double>(n + 1)) - lgamma(static_cast<double>(x + 1)) - lgamma(static_cast<double>(n - x + 1));
#pragma unroll
  for (int r = 0; r < R; r++) {
    const double al = a.alpha[r], be = a.beta[r];
    const double v = c0 + lgamma(x + al) + lgamma(n - x + be) - lgamma(n + al + be) + lgamma(al + be) - lgamma(al) - lgamma(be);
    acc[r] += v;
  }
}

template <int R>
__device__ __forceinline__ void emission_add(const SgEmissionArgs& a, const double* stab, uint32_t rows_smem, uint32_t x, uint32_t n,
                                             double (&acc)[R]) {
  if (x > n) {  // impossible count: density zero (misc.h:636-639)
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] += -HYG_INF;
    return;
  }
  const uint32_t row = n * (n + 1) / 2 + x;
  if (row < rows_smem) {
    const double* p = stab + static_cast<size_t>(row) * R;
    if (R % 2 == 0) {
      const double2* p2 = reinterpret_cast<const double2*>(p);
#pragma unroll
      for (int r = 0; r < R / 2; r++) {
        const double2 v = p2[r];
        acc[2 * r] += v.x;
        acc[2 * r + 1] += v.y;
      }
    } else {
#pragma unroll
      for (int r = 0; r < R; r++) acc[r] += p[r];
    }
  } else if (n <= static_cast<uint32_t>(a.nmax_table)) {
    const double* p = a.table + static_cast<size_t>(row) * R;
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] += __ldg(p + r);
  } else {
    emission_direct<R>(a, x, n, acc);
  }
}

template <int R>
__device__ __forceinline__ void sg_emission_entry(const SgEmissionArgs a) {
  double* stab = HYG_EM_SMEM;
  const uint32_t rows_smem = static_cast<uint32_t>(a.nmax_smem + 1) * (a.nmax_smem + 2) / 2;
  // stage the hot part of the table in shared memory
  for (uint32_t i = threadIdx.x; i < rows_smem * R; i += blockDim.x) stab[i] = __ldg(a.table + i);
  __syncthreads();

  const unsigned long long n_pairs = (a.T + 1) / 2;
  const uint32_t* nt32 = reinterpret_cast<const uint32_t*>(a.n_total);
  const uint32_t* nm32 = reinterpret_cast<const uint32_t*>(a.n_meth);
  const unsigned long long pitch32 = a.pitch / 2;
  for (unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_pairs;
       i += static_cast<unsigned long long>(gridDim.x) * blockDim.x) {
    double acc0[R], acc1[R];
#pragma unroll
    for (int r = 0; r < R; r++) { acc0[r] = 0.0; acc1[r] = 0.0; }
    uint32_t s = 0;
    for (; s + 4 <= a.S; s += 4) {  // four samples in flight per thread
      uint32_t nn[4], xx[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        nn[k] = __ldg(nt32 + (s + k) * pitch32 + i);
        xx[k] = __ldg(nm32 + (s + k) * pitch32 + i);
      }
#pragma unroll
      for (int k = 0; k < 4; k++) {
        emission_add<R>(a, stab, rows_smem, xx[k] & 0xFFFFu, nn[k] & 0xFFFFu, acc0);
        emission_add<R>(a, stab, rows_smem, xx[k] >> 16, nn[k] >> 16, acc1);
      }
    }
    for (; s < a.S; s++) {
      const uint32_t nn = __ldg(nt32 + s * pitch32 + i), xx = __ldg(nm32 + s * pitch32 + i);
      emission_add<R>(a, stab, rows_smem, xx & 0xFFFFu, nn & 0xFFFFu, acc0);
      emission_add<R>(a, stab, rows_smem, xx >> 16, nn >> 16, acc1);
    }
    double* out = a.logobs + 2 * i * R;
    const bool second = (2 * i + 1 < a.T);
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

#ifndef HYG_EMU
template <int R>
__global__ void __launch_bounds__(HYG_EM_NT, 1) sg_emission_kernel(const SgEmissionArgs a) {
  sg_emission_entry<R>(a);
}
#endif

}  // namespace hyg
#endif
