// cuda_emul.h -- a CPU execution model of the CUDA subset the *_kernels.cuh files of csrc/ are written in.
//
// TEST INFRASTRUCTURE ONLY (like oracle/): it lets `pytest -m "not gpu"` run the very same kernel source and the very
// same host pipelines (allocation sizes, launch geometry, scans) on the CPU, so that index arithmetic, barriers and
// warp-collective logic are checked before any GPU time is spent.  Nothing in the product links or includes it.
//
// Model: one OS thread.  Every CUDA thread of a block is a fiber (ucontext); blocks run one after another.  A fiber
// runs until it reaches a block barrier or a warp collective, where it yields; a barrier releases when every live
// thread of the block (warp) has arrived, exactly the CUDA rule, and a round in which no fiber can make progress is
// reported as a deadlock (divergent __syncthreads / missing lane in a *_sync collective).  Scheduling is round-robin
// and deterministic, in forward, reverse or (seeded) random thread order (JP_EMUL_SCHEDULE): order-dependent results
// expose missing synchronisation, but this is not a memory-model checker.  Asynchronous copies (cp.async, TMA bulk copies)
// complete at issue; mbarrier waits therefore pass.
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define JP_EMUL 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static
// dynamic shared memory: CUDA `extern __shared__ __align__(128) unsigned char name[];`
#define JP_DYN_SMEM(name) unsigned char* name = ::emul::state().dyn_smem

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(16) double2 { double x, y; };
inline int2 make_int2(int x, int y) { return int2{x, y}; }
inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
inline double2 make_double2(double x, double y) { return double2{x, y}; }

namespace emul {

struct Fiber {
  ucontext_t ctx;
  char* stack = nullptr;
  bool done = true;
};
struct Rendezvous {
  int alive = 0, arrived = 0;
  unsigned gen = 0;
};
struct Warp {
  Rendezvous rv;
  uint64_t vals[32];
};
struct State {
  uint3 threadIdx{0, 0, 0}, blockIdx{0, 0, 0};
  dim3 blockDim, gridDim;
  unsigned char* dyn_smem = nullptr;
  std::vector<Fiber> fibers;
  std::vector<Warp> warps;
  Rendezvous block_rv;
  uint64_t block_vals = 0;  // __syncthreads_or / _count accumulator
  ucontext_t sched;
  int cur = -1;
  uint64_t progress = 0;
  const std::function<void()>* body = nullptr;
  uint64_t launches = 0;
};
inline State& state() {
  static State s;
  return s;
}

constexpr size_t kStackBytes = 256 * 1024;

inline void yield() {
  State& s = state();
  swapcontext(&s.fibers[s.cur].ctx, &s.sched);
}
inline void rendezvous(Rendezvous& r) {
  State& s = state();
  const unsigned gen = r.gen;
  if (++r.arrived >= r.alive) {
    r.arrived = 0;
    ++r.gen;
    ++s.progress;
  } else {
    while (r.gen == gen) yield();
  }
}
inline void on_fiber_exit(int tid) {
  State& s = state();
  ++s.progress;
  Rendezvous* rs[2] = {&s.block_rv, &s.warps[tid >> 5].rv};
  for (Rendezvous* r : rs) {
    --r->alive;
    if (r->alive > 0 && r->arrived >= r->alive) {  // the exit completes a pending barrier (CUDA counts exited threads)
      r->arrived = 0;
      ++r->gen;
    }
  }
}
inline void trampoline() {
  State& s = state();
  const int tid = s.cur;
  (*s.body)();
  s.fibers[tid].done = true;
  on_fiber_exit(tid);
  swapcontext(&s.fibers[tid].ctx, &s.sched);
}

inline int linear_tid() {
  State& s = state();
  return (int)(s.threadIdx.x + s.blockDim.x * (s.threadIdx.y + s.blockDim.y * s.threadIdx.z));
}
inline Warp& cur_warp() { return state().warps[linear_tid() >> 5]; }
inline int lane_id() { return linear_tid() & 31; }
inline bool lane_alive(int lane) {
  State& s = state();
  const int t = (linear_tid() & ~31) + lane;
  return t < (int)s.fibers.size() && !s.fibers[t].done;
}

template <class T>
T shfl_from(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
  Warp& w = cur_warp();
  const int lane = lane_id();
  uint64_t raw = 0;
  std::memcpy(&raw, &v, sizeof(T));
  w.vals[lane] = raw;
  rendezvous(w.rv);
  T out = v;
  if (src_lane >= 0 && src_lane < 32 && lane_alive(src_lane)) std::memcpy(&out, &w.vals[src_lane], sizeof(T));
  rendezvous(w.rv);
  return out;
}
inline unsigned ballot(bool pred) {
  Warp& w = cur_warp();
  w.vals[lane_id()] = pred ? 1u : 0u;
  rendezvous(w.rv);
  unsigned bits = 0;
  for (int l = 0; l < 32; ++l)
    if (lane_alive(l) && w.vals[l]) bits |= 1u << l;
  rendezvous(w.rv);
  return bits;
}

// The order in which the runnable fibers of a block get their turn in one scheduling round.  JP_EMUL_SCHEDULE =
// "forward" (default), "reverse", or "random[:seed]" (a fresh permutation every round): a kernel whose result depends
// on the order -- a missing __syncthreads between a shared-memory write and its read, a buffer reused before every
// reader is done -- gives different answers under different schedules, which the tests run under all three.
inline int& schedule_mode() {
  static int mode = -1;  // -1: read JP_EMUL_SCHEDULE at first use
  return mode;
}
inline uint64_t& schedule_rng() {
  static uint64_t rng = 0x9E3779B97F4A7C15ull;
  return rng;
}
inline void schedule_order(std::vector<int>& order, int nthreads) {
  int& mode = schedule_mode();
  uint64_t& rng = schedule_rng();
  if (mode < 0) {
    const char* v = std::getenv("JP_EMUL_SCHEDULE");
    mode = 0;
    if (v && std::strncmp(v, "reverse", 7) == 0) mode = 1;
    if (v && std::strncmp(v, "random", 6) == 0) {
      mode = 2;
      if (v[6] == ':') rng ^= std::strtoull(v + 7, nullptr, 10) * 0xBF58476D1CE4E5B9ull;
    }
  }
  if ((int)order.size() != nthreads) {
    order.resize(nthreads);
    for (int i = 0; i < nthreads; ++i) order[i] = mode == 1 ? nthreads - 1 - i : i;
  }
  if (mode == 2) {
    for (int i = nthreads - 1; i > 0; --i) {  // Fisher-Yates with splitmix64
      rng += 0x9E3779B97F4A7C15ull;
      uint64_t z = rng;
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      z ^= z >> 31;
      std::swap(order[i], order[(int)(z % (uint64_t)(i + 1))]);
    }
  }
}

// run one grid: body() is the kernel call with its arguments bound
inline void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body) {
  State& s = state();
  std::vector<int> order;
  if (s.cur != -1) {
    std::fprintf(stderr, "cuda_emul: nested launch\n");
    std::abort();
  }
  ++s.launches;
  const int nthreads = (int)(block.x * block.y * block.z);
  if (nthreads <= 0 || nthreads > 1024) {
    std::fprintf(stderr, "cuda_emul: bad block size %d\n", nthreads);
    std::abort();
  }
  if (smem_bytes > 227 * 1024) {
    std::fprintf(stderr, "cuda_emul: %zu bytes of dynamic shared memory exceed the 227 KB limit\n", smem_bytes);
    std::abort();
  }
  if ((int)s.fibers.size() < nthreads) {
    const size_t old = s.fibers.size();
    s.fibers.resize(nthreads);
    for (size_t i = old; i < s.fibers.size(); ++i) s.fibers[i].stack = (char*)std::malloc(kStackBytes);
  }
  std::vector<unsigned char> smem(smem_bytes + 256);
  s.dyn_smem = (unsigned char*)(((uintptr_t)smem.data() + 127) & ~(uintptr_t)127);
  s.blockDim = block;
  s.gridDim = grid;
  s.body = &body;
  const int nwarps = (nthreads + 31) / 32;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        s.blockIdx = uint3{bx, by, bz};
        std::memset(s.dyn_smem, 0xCD, smem_bytes);  // shared memory is NOT zero-initialised
        s.warps.assign(nwarps, Warp());
        s.block_rv = Rendezvous();
        s.block_rv.alive = nthreads;
        for (int w = 0; w < nwarps; ++w) s.warps[w].rv.alive = (w + 1) * 32 <= nthreads ? 32 : nthreads - w * 32;
        // fibers beyond nthreads stay `done`; note lane_alive() looks at fibers.size(), so mark the tail done
        for (size_t t = 0; t < s.fibers.size(); ++t) s.fibers[t].done = (int)t >= nthreads;
        for (int t = 0; t < nthreads; ++t) {
          Fiber& f = s.fibers[t];
          getcontext(&f.ctx);
          f.ctx.uc_stack.ss_sp = f.stack;
          f.ctx.uc_stack.ss_size = kStackBytes;
          f.ctx.uc_link = &s.sched;
          makecontext(&f.ctx, (void (*)())trampoline, 0);
        }
        int remaining = nthreads;
        while (remaining > 0) {
          const uint64_t before = s.progress;
          remaining = 0;
          schedule_order(order, nthreads);
          for (int oi = 0; oi < nthreads; ++oi) {
            const int t = order[oi];
            if (s.fibers[t].done) continue;
            s.cur = t;
            s.threadIdx = uint3{(unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y)};
            swapcontext(&s.sched, &s.fibers[t].ctx);
            if (!s.fibers[t].done) ++remaining;
          }
          if (remaining > 0 && s.progress == before) {
            std::fprintf(stderr, "cuda_emul: deadlock in block (%u,%u,%u): %d threads wait at a barrier or warp collective "
                                 "the others never reach\n", bx, by, bz, remaining);
            std::abort();
          }
        }
      }
  s.cur = -1;
  s.body = nullptr;
  s.dyn_smem = nullptr;
}

}  // namespace emul

#define threadIdx (::emul::state().threadIdx)
#define blockIdx (::emul::state().blockIdx)
#define blockDim (::emul::state().blockDim)
#define gridDim (::emul::state().gridDim)

// ---- barriers and warp collectives ----
inline void __syncthreads() { emul::rendezvous(emul::state().block_rv); }
inline int __syncthreads_or(int pred) {
  emul::State& s = emul::state();
  s.block_vals = 0;  // every thread clears, nobody sets before the first barrier
  emul::rendezvous(s.block_rv);
  if (pred) s.block_vals = 1;
  emul::rendezvous(s.block_rv);
  const int r = (int)s.block_vals;
  emul::rendezvous(s.block_rv);  // nobody clears for the next use before everyone has read
  return r;
}
inline void __syncwarp(unsigned = 0xffffffffu) { emul::rendezvous(emul::cur_warp().rv); }
template <class T>
T __shfl_sync(unsigned, T v, int src, int = 32) { return emul::shfl_from(v, src & 31); }
template <class T>
T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return emul::shfl_from(v, emul::lane_id() ^ m); }
template <class T>
T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) { return emul::shfl_from(v, emul::lane_id() + (int)d); }
template <class T>
T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) { return emul::shfl_from(v, emul::lane_id() - (int)d); }
inline unsigned __ballot_sync(unsigned, int pred) { return emul::ballot(pred != 0); }
inline int __any_sync(unsigned, int pred) { return emul::ballot(pred != 0) != 0; }
inline int __all_sync(unsigned m, int pred) { return emul::ballot(pred == 0) == 0; }
inline unsigned __activemask() {
  unsigned bits = 0;
  for (int l = 0; l < 32; ++l)
    if (emul::lane_alive(l)) bits |= 1u << l;
  return bits;
}

// ---- memory: one OS thread, so atomics are plain read-modify-writes ----
inline void __threadfence() {}
inline void __threadfence_block() {}
inline void __threadfence_system() {}
template <class T>
T __ldg(const T* p) { return *p; }
template <class T>
T __ldcs(const T* p) { return *p; }
template <class T, class U>
T atomicAdd(T* p, U v) { T old = *p; *p = (T)(old + (T)v); return old; }
template <class T, class U>
T atomicOr(T* p, U v) { T old = *p; *p = (T)(old | (T)v); return old; }
template <class T, class U>
T atomicAnd(T* p, U v) { T old = *p; *p = (T)(old & (T)v); return old; }
template <class T, class U>
T atomicMax(T* p, U v) { T old = *p; if ((T)v > old) *p = (T)v; return old; }
template <class T, class U>
T atomicMin(T* p, U v) { T old = *p; if ((T)v < old) *p = (T)v; return old; }
template <class T, class U>
T atomicExch(T* p, U v) { T old = *p; *p = (T)v; return old; }
template <class T, class U, class V>
T atomicCAS(T* p, U cmp, V v) { T old = *p; if (old == (T)cmp) *p = (T)v; return old; }

// ---- arithmetic intrinsics (compile with -ffp-contract=off: a*b+c is never fused unless fma() is written) ----
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
using std::fma;
using std::pow;
using std::sqrt;
