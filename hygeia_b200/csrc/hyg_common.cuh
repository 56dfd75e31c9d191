// hygeia_b200/csrc/hyg_common.cuh -- device-side building blocks shared by the kernels.
//
// Written for sm_100a (nvcc) but deliberately limited to full-warp collectives and
// __syncthreads so that tests/emu/ can run the same code under a CPU emulation of the
// CUDA execution model on the GPU-less build box (test infrastructure only).
#ifndef HYG_COMMON_CUH
#define HYG_COMMON_CUH

#include <stdint.h>

#ifndef HYG_EMU
#include <cuda_runtime.h>
#include <math_constants.h>
#define HYG_INF CUDART_INF
#else
#include <limits>
#define HYG_INF (std::numeric_limits<double>::infinity())
#endif

#define HYG_FULL 0xffffffffu
#define HYG_WORKER_WARPS 8  // warps that own particles (one slot per thread)
#define HYG_NT 288          // threads per chain CTA: 256 particle slots + one service warp
#define HYG_NW (HYG_NT / 32)
#define HYG_RMAX 8          // max number of regimes
#define HYG_NPMAX 256       // max particles per chain

// Debug builds (-DHYG_DEBUG_CHECKS): index checks in the recursion kernels; a failing one prints its code and traps.
#if defined(HYG_DEBUG_CHECKS) && !defined(HYG_EMU)
#include <cstdio>
#define HYG_CHECK(cond, code, a, b) do { if (!(cond)) { printf("HYG_CHECK %d failed: tid %d block %d a=%lld b=%lld\n", (code), (int)threadIdx.x, (int)blockIdx.x, (long long)(a), (long long)(b)); __trap(); } } while (0)
#else
#define HYG_CHECK(cond, code, a, b) do { } while (0)
#endif

namespace hyg {

__device__ __forceinline__ int hyg_tid() { return static_cast<int>(threadIdx.x); }

// bar.sync id, count: a barrier among `count` threads (whole warps) of the CTA; id 1..15 (0 is __syncthreads)
__device__ __forceinline__ void named_barrier(int id, int count) {
#ifdef HYG_EMU
  hyg_emu_named_barrier(id, static_cast<unsigned>(count));
#else
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
#endif
}

// bar.arrive id, count: signals arrival without waiting (producer side of a producer / consumer barrier; the consumers'
// bar.sync on the same id completes once `count` threads have arrived or are waiting)
__device__ __forceinline__ void named_arrive(int id, int count) {
#ifdef HYG_EMU
  hyg_emu_named_arrive(id, static_cast<unsigned>(count));
#else
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
#endif
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(HYG_FULL, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double t = __shfl_xor_sync(HYG_FULL, v, o);
    v = t > v ? t : v;
  }
  return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(HYG_FULL, v, o);
  return v;
}

// Double-buffered cross-warp scratch: consecutive block-wide reductions need only ONE barrier each.
struct BlockScratch {
  double d[2][HYG_NW][16];
  int flip;
};

// Block-wide sum of K (<= 16) per-thread values; every thread gets every total.  One __syncthreads.
template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], BlockScratch& sc, int& flip) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; k++) {
    double s = warp_sum(v[k]);
    if (lane == 0) sc.d[flip][warp][k] = s;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; k++) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < HYG_NW; w++) s += sc.d[flip][w][k];
    v[k] = s;
  }
  flip ^= 1;
}
__device__ __forceinline__ double block_max(double v, BlockScratch& sc, int& flip) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double s = warp_max(v);
  if (lane == 0) sc.d[flip][warp][0] = s;
  __syncthreads();
  s = sc.d[flip][0][0];
#pragma unroll
  for (int w = 1; w < HYG_NW; w++) {
    double t = sc.d[flip][w][0];
    s = t > s ? t : s;
  }
  flip ^= 1;
  return s;
}

// Monotone map double -> uint64 (ascending, all 64 bits kept): -inf -> 0x000f..f, finite values above it.  0 is below
// every image and marks "no particle".
__device__ __forceinline__ unsigned long long order_key(double x) {
  const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(x));
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
#define HYG_KEY_NEGINF 0x000fffffffffffffull
// Canonical particle order: log-weight descending, exact ties by (regime, sojourn) ascending -- a
// rule that does not depend on where a particle is stored.  pay = (regime << 28 | sojourn) << 8 | slot.
__device__ __forceinline__ unsigned long long order_pay(int r, uint32_t d, int slot) {
  return (((static_cast<unsigned long long>(r) << 28) | static_cast<unsigned long long>(d & 0x0fffffffu)) << 8) | static_cast<unsigned long long>(slot);
}
__device__ __forceinline__ bool order_before(unsigned long long ka, unsigned long long pa, unsigned long long kb, unsigned long long pb) {
  return (ka > kb) || (ka == kb && pa < pb);
}
// order-independent hash of a support point (parity tap; splitmix64 finaliser -- the CPU checker uses the same one)
__host__ __device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  unsigned long long z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter-based: the resampling uniform of
// site t of chain c under seed s is a pure function of (s, c, t).  Host copy in
// hygeia_b200/philox.py; known-answer vectors in tests/test_philox.py.
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; i++) {
    const uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * c[0];
    const uint64_t p1 = static_cast<uint64_t>(0xCD9E8D57u) * c[2];
    const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c[1] ^ k0;
    const uint32_t n1 = static_cast<uint32_t>(p1);
    const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c[3] ^ k1;
    const uint32_t n3 = static_cast<uint32_t>(p0);
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
#define HYG_PHILOX_TAG 0x48594745u  // "HYGE"
__host__ __device__ __forceinline__ double philox_uniform(uint64_t seed, uint32_t chain, uint64_t t) {
  uint32_t c[4] = {static_cast<uint32_t>(t), static_cast<uint32_t>(t >> 32), chain, HYG_PHILOX_TAG};
  philox4x32_10(c, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  const uint64_t hi = c[0] >> 5, lo = c[1] >> 6;  // 27 + 26 = 53 random bits
  return (static_cast<double>(hi) * 67108864.0 + static_cast<double>(lo)) * (1.0 / 9007199254740992.0);
}

}  // namespace hyg
#endif
