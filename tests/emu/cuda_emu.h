// tests/emu/cuda_emu.h -- TEST INFRASTRUCTURE, not product code.
//
// A tiny CUDA execution-model emulator so that the *device code* of this repo
// (hygeia_b200/csrc/*.cuh) can be compiled with plain g++ and exercised on the
// GPU-less build box.  It is NOT a fallback: nothing in the product links it;
// only tests/emu/*.cpp (CPU-side `pytest -m "not gpu"` checks of kernel logic
// against the oracle) include it.
//
// Model: one thread block at a time; every CUDA thread is a ucontext fiber on a
// single OS thread.  A fiber runs until it reaches a collective
// (__syncthreads, __syncwarp, __shfl*_sync, __ballot_sync, ...), where it
// parks until every participating fiber has arrived.  The scheduler runs the
// 32 lanes of a warp back to back so warp collectives complete in one sweep.
// Only full-mask, converged warp collectives are supported (which is all the
// kernels use); a collective reached by a partial warp deadlocks and aborts
// loudly instead of silently mis-computing.
#ifndef HYG_CUDA_EMU_H
#define HYG_CUDA_EMU_H


#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define HYG_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct emu_dim3 {
  unsigned x, y, z;
  emu_dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
typedef emu_dim3 dim3;

namespace emu {

// Minimal x86-64 SysV context switch (glibc's swapcontext makes a sigprocmask syscall per switch, ~100x slower).
extern "C" void hyg_emu_ctx_switch(void** save_sp, void* new_sp);
__asm__(
    ".text\n"
    ".globl hyg_emu_ctx_switch\n"
    ".type hyg_emu_ctx_switch,@function\n"
    "hyg_emu_ctx_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n"
    "  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n"
    "  ret\n"
    ".size hyg_emu_ctx_switch,.-hyg_emu_ctx_switch\n");

struct Fiber {
  void* sp = nullptr;
  std::vector<char> stack;
  bool done = false;
  unsigned tid = 0;
  // parked at: 0 none, 1 block barrier, 2 warp collective, 3 named barrier `bar_id`
  int parked = 0;
  int bar_id = 0;
  unsigned long long gen = 0;
};

struct NamedBarrier {   // bar.sync id, count
  unsigned long long gen = 0;
  unsigned arrived = 0;
};

struct WarpState {
  unsigned long long gen = 0;
  unsigned arrived = 0;
  uint64_t slot[32];
  uint64_t result[32];
  unsigned ballot = 0;
};

struct BlockState {
  std::vector<Fiber> fibers;
  std::vector<WarpState> warps;
  unsigned long long gen = 0;
  unsigned arrived = 0;
  NamedBarrier named[16];
  unsigned live = 0;
  unsigned nthreads = 0;
  void* sched_sp = nullptr;
  int current = -1;
  std::function<void()> body;
};

inline BlockState*& cur_block() { static BlockState* b = nullptr; return b; }

}  // namespace emu

static emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
static const int warpSize = 32;

namespace emu {

inline void yield_to_sched() {
  BlockState* b = cur_block();
  Fiber& f = b->fibers[b->current];
  hyg_emu_ctx_switch(&f.sp, b->sched_sp);
  threadIdx = emu_dim3(f.tid);
}

inline void fiber_entry() {
  BlockState* b = cur_block();
  b->body();
  Fiber& f = b->fibers[b->current];
  f.done = true;
  b->live--;
  hyg_emu_ctx_switch(&f.sp, b->sched_sp);
  std::abort();  // a finished fiber is never resumed
}

inline void block_barrier() {
  BlockState* b = cur_block();
  Fiber& f = b->fibers[b->current];
  unsigned long long my = b->gen;
  b->arrived++;
  if (b->arrived == b->live) {
    b->arrived = 0;
    b->gen++;
    return;
  }
  f.parked = 1;
  while (b->gen == my) yield_to_sched();
  f.parked = 0;
}

// bar.sync id, count: the first `count` threads to arrive are released together
inline void named_barrier(int id, unsigned count) {
  BlockState* b = cur_block();
  Fiber& f = b->fibers[b->current];
  NamedBarrier& nb = b->named[id];
  unsigned long long my = nb.gen;
  nb.arrived++;
  if (nb.arrived == count) {
    nb.arrived = 0;
    nb.gen++;
    return;
  }
  f.parked = 3;
  f.bar_id = id;
  while (nb.gen == my) yield_to_sched();
  f.parked = 0;
}

// bar.arrive id, count: counts as an arrival, does not wait
inline void named_arrive(int id, unsigned count) {
  BlockState* b = cur_block();
  NamedBarrier& nb = b->named[id];
  nb.arrived++;
  if (nb.arrived == count) {
    nb.arrived = 0;
    nb.gen++;
  }
}

inline WarpState& my_warp() {
  BlockState* b = cur_block();
  return b->warps[b->fibers[b->current].tid / 32];
}
inline unsigned warp_width() {
  BlockState* b = cur_block();
  unsigned w = b->fibers[b->current].tid / 32;
  unsigned n = b->nthreads - w * 32;
  return n > 32 ? 32 : n;
}
// Two-phase warp rendezvous: everyone deposits, last arriver runs `combine`, everyone picks up.
template <class Combine> inline void warp_rendezvous(uint64_t deposit, Combine combine) {
  BlockState* b = cur_block();
  Fiber& f = b->fibers[b->current];
  WarpState& w = my_warp();
  unsigned lane = f.tid % 32;
  unsigned long long my = w.gen;
  w.slot[lane] = deposit;
  w.arrived++;
  if (w.arrived == warp_width()) {
    combine(w);
    w.arrived = 0;
    w.gen++;
    return;
  }
  f.parked = 2;
  while (w.gen == my) yield_to_sched();
  f.parked = 0;
}

template <class K> void launch(dim3 grid, dim3 block, K kernel_body) {
  gridDim = grid;
  blockDim = block;
  const unsigned nthreads = block.x;
  const size_t STACK = 256 * 1024;
  for (unsigned bx = 0; bx < grid.x; bx++) {
    BlockState bs;
    bs.nthreads = nthreads;
    bs.live = nthreads;
    bs.fibers.resize(nthreads);
    bs.warps.resize((nthreads + 31) / 32);
    bs.body = kernel_body;
    cur_block() = &bs;
    blockIdx = emu_dim3(bx);
    for (unsigned t = 0; t < nthreads; t++) {
      Fiber& f = bs.fibers[t];
      f.tid = t;
      f.stack.resize(STACK);
      // initial frame: six callee-saved registers (zero) + return address = fiber_entry; after `ret` rsp % 16 == 8
      uintptr_t top = (reinterpret_cast<uintptr_t>(f.stack.data()) + STACK) & ~static_cast<uintptr_t>(15);
      void** sp = reinterpret_cast<void**>(top - 8);   // slot that `ret` leaves above: keeps rsp % 16 == 8 at entry
      *(--sp) = reinterpret_cast<void*>(&fiber_entry);  // return address
      for (int k = 0; k < 6; k++) *(--sp) = nullptr;    // rbp rbx r12 r13 r14 r15
      f.sp = sp;
    }
    // scheduler: sweep warp by warp; within a warp keep sweeping lanes while anything progresses
    unsigned long long idle_sweeps = 0;
    while (bs.live > 0) {
      bool any = false;
      for (unsigned w = 0; w < bs.warps.size(); w++) {
        bool progressed = true;
        while (progressed) {
          progressed = false;
          for (unsigned l = 0; l < 32 && w * 32 + l < nthreads; l++) {
            Fiber& f = bs.fibers[w * 32 + l];
            if (f.done) continue;
            if (f.parked == 1 && f.gen == bs.gen) continue;                 // still waiting on block barrier
            if (f.parked == 2 && f.gen == bs.warps[w].gen) continue;       // still waiting on warp collective
            if (f.parked == 3 && f.gen == bs.named[f.bar_id].gen) continue; // still waiting on a named barrier
            // remember generations so the skip tests above are valid
            bs.current = static_cast<int>(w * 32 + l);
            threadIdx = emu_dim3(f.tid);
            hyg_emu_ctx_switch(&bs.sched_sp, f.sp);
            // after it yields, snapshot what it is waiting for
            if (!f.done) f.gen = (f.parked == 1) ? bs.gen : (f.parked == 2 ? bs.warps[w].gen : (f.parked == 3 ? bs.named[f.bar_id].gen : 0));
            progressed = true;
            any = true;
          }
          // a lane that just parked on a warp collective with generation == current is not runnable; loop ends when
          // no lane in this warp can run
          bool runnable = false;
          for (unsigned l = 0; l < 32 && w * 32 + l < nthreads; l++) {
            Fiber& f = bs.fibers[w * 32 + l];
            if (f.done) continue;
            if (f.parked == 1 && f.gen == bs.gen) continue;
            if (f.parked == 2 && f.gen == bs.warps[w].gen) continue;
            if (f.parked == 3 && f.gen == bs.named[f.bar_id].gen) continue;
            runnable = true;
          }
          progressed = runnable;
        }
      }
      if (!any) {
        if (++idle_sweeps > 4) {
          std::fprintf(stderr, "cuda_emu: deadlock (a collective was reached by a partial warp/block)\n");
          for (unsigned w2 = 0; w2 * 32 < nthreads; w2++) {
            const Fiber& f0 = bs.fibers[w2 * 32];
            std::fprintf(stderr, "  warp %u lane 0: %s parked=%d (1 block barrier, 2 warp collective, 3 named barrier) id=%d; named[%d].arrived=%u block.arrived=%u\n",
                         w2, f0.done ? "done" : "live", f0.parked, f0.bar_id, f0.bar_id, bs.named[f0.bar_id & 15].arrived, bs.arrived);
          }
          { const unsigned w2 = (nthreads - 1) / 32; std::fprintf(stderr, "  last warp lanes parked:");
            for (unsigned l = 0; l < 32 && w2 * 32 + l < nthreads; l++) std::fprintf(stderr, " %d", bs.fibers[w2 * 32 + l].parked);
            std::fprintf(stderr, "\n"); }
          std::abort();
        }
      } else {
        idle_sweeps = 0;
      }
    }
    cur_block() = nullptr;
  }
}

}  // namespace emu

// ---------------------------------------------------------------------------
// CUDA intrinsics used by the kernels
// ---------------------------------------------------------------------------
inline void __syncthreads() { emu::block_barrier(); }
inline void hyg_emu_named_barrier(int id, unsigned count) { emu::named_barrier(id, count); }
inline void hyg_emu_named_arrive(int id, unsigned count) { emu::named_arrive(id, count); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_rendezvous(0, [](emu::WarpState&) {}); }
inline void __threadfence() {}
inline void __threadfence_block() {}

template <class T> inline uint64_t emu_pack(T v) { uint64_t u = 0; std::memcpy(&u, &v, sizeof(T)); return u; }
template <class T> inline T emu_unpack(uint64_t u) { T v; std::memcpy(&v, &u, sizeof(T)); return v; }

template <class T> inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
  unsigned lane = threadIdx.x % 32;
  emu::warp_rendezvous(emu_pack(v), [](emu::WarpState& w) { for (int i = 0; i < 32; i++) w.result[i] = w.slot[i]; });
  int base = (lane / width) * width;
  return emu_unpack<T>(emu::my_warp().result[base + (src % width)]);
}
template <class T> inline T __shfl_xor_sync(unsigned m, T v, int mask, int width = 32) {
  unsigned lane = threadIdx.x % 32;
  emu::warp_rendezvous(emu_pack(v), [](emu::WarpState& w) { for (int i = 0; i < 32; i++) w.result[i] = w.slot[i]; });
  (void)m; (void)width;
  return emu_unpack<T>(emu::my_warp().result[lane ^ mask]);
}
template <class T> inline T __shfl_up_sync(unsigned, T v, unsigned delta, int = 32) {
  unsigned lane = threadIdx.x % 32;
  emu::warp_rendezvous(emu_pack(v), [](emu::WarpState& w) { for (int i = 0; i < 32; i++) w.result[i] = w.slot[i]; });
  return lane >= delta ? emu_unpack<T>(emu::my_warp().result[lane - delta]) : v;
}
template <class T> inline T __shfl_down_sync(unsigned, T v, unsigned delta, int = 32) {
  unsigned lane = threadIdx.x % 32;
  emu::warp_rendezvous(emu_pack(v), [](emu::WarpState& w) { for (int i = 0; i < 32; i++) w.result[i] = w.slot[i]; });
  return lane + delta < 32 ? emu_unpack<T>(emu::my_warp().result[lane + delta]) : v;
}
inline unsigned __ballot_sync(unsigned, int pred) {
  emu::warp_rendezvous(pred ? 1 : 0, [](emu::WarpState& w) {
    unsigned b = 0;
    unsigned n = emu::warp_width();
    for (unsigned i = 0; i < n; i++) if (w.slot[i]) b |= (1u << i);
    w.ballot = b;
  });
  return emu::my_warp().ballot;
}
inline unsigned __reduce_or_sync(unsigned, unsigned v) {
  emu::warp_rendezvous(v, [](emu::WarpState& w) {
    unsigned b = 0;
    unsigned n = emu::warp_width();
    for (unsigned i = 0; i < n; i++) b |= static_cast<unsigned>(w.slot[i]);
    w.ballot = b;
  });
  return emu::my_warp().ballot;
}
inline int __reduce_max_sync(unsigned, int v) {
  emu::warp_rendezvous(static_cast<uint64_t>(static_cast<int64_t>(v)), [](emu::WarpState& w) {
    int64_t b = INT64_MIN;
    unsigned n = emu::warp_width();
    for (unsigned i = 0; i < n; i++) b = static_cast<int64_t>(w.slot[i]) > b ? static_cast<int64_t>(w.slot[i]) : b;
    w.result[0] = static_cast<uint64_t>(b);
  });
  return static_cast<int>(static_cast<int64_t>(emu::my_warp().result[0]));
}
inline int __reduce_min_sync(unsigned, int v) {
  emu::warp_rendezvous(static_cast<uint64_t>(static_cast<int64_t>(v)), [](emu::WarpState& w) {
    int64_t b = INT64_MAX;
    unsigned n = emu::warp_width();
    for (unsigned i = 0; i < n; i++) b = static_cast<int64_t>(w.slot[i]) < b ? static_cast<int64_t>(w.slot[i]) : b;
    w.result[0] = static_cast<uint64_t>(b);
  });
  return static_cast<int>(static_cast<int64_t>(emu::my_warp().result[0]));
}
inline int __reduce_add_sync(unsigned, int v) {
  emu::warp_rendezvous(static_cast<uint64_t>(static_cast<int64_t>(v)), [](emu::WarpState& w) {
    int64_t b = 0;
    unsigned n = emu::warp_width();
    for (unsigned i = 0; i < n; i++) b += static_cast<int64_t>(w.slot[i]);
    w.result[0] = static_cast<uint64_t>(b);
  });
  return static_cast<int>(static_cast<int64_t>(emu::my_warp().result[0]));
}
// mask of the lanes holding the same value as the caller (the last lane to arrive works out all the masks)
inline unsigned __match_any_sync(unsigned, int v) {
  unsigned lane = threadIdx.x % 32;
  emu::warp_rendezvous(static_cast<uint64_t>(static_cast<int64_t>(v)), [](emu::WarpState& w) {
    unsigned n = emu::warp_width();
    for (unsigned i = 0; i < n; i++) {
      unsigned b = 0;
      for (unsigned j = 0; j < n; j++) if (w.slot[j] == w.slot[i]) b |= (1u << j);
      w.result[i] = b;
    }
  });
  return static_cast<unsigned>(emu::my_warp().result[lane]);
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __all_sync(unsigned m, int pred) {
  unsigned n = emu::warp_width();
  unsigned full = (n == 32) ? 0xffffffffu : ((1u << n) - 1);
  return __ballot_sync(m, pred) == full;
}

inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __clz(int x) { return x ? __builtin_clz(static_cast<unsigned>(x)) : 32; }
inline long long __double_as_longlong(double d) { long long l; std::memcpy(&l, &d, 8); return l; }
inline double __longlong_as_double(long long l) { double d; std::memcpy(&d, &l, 8); return d; }
inline unsigned __umulhi(unsigned a, unsigned b) { return static_cast<unsigned>((static_cast<uint64_t>(a) * b) >> 32); }
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T atomicAdd(T* p, T v) { T old = *p; *p = old + v; return old; }
template <class T> inline T atomicMax(T* p, T v) { T old = *p; *p = old > v ? old : v; return old; }
template <class T> inline T atomicMin(T* p, T v) { T old = *p; *p = old < v ? old : v; return old; }
template <class T> inline T atomicOr(T* p, T v) { T old = *p; *p = old | v; return old; }
inline double __dadd_rn(double a, double b) { return a + b; }

struct double2 { double x, y; };
inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
struct ushort2 { unsigned short x, y; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };

#endif
