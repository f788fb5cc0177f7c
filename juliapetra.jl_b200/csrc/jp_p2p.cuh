// jp_p2p.cuh -- peer-to-peer halo protocol shared by jp_plan.cu (setup, two-stream path) and jp_csr.cu (fused kernel)
#pragma once
#include <cstdint>

namespace jp {

struct P2PSend {
  int32_t start, count;        // slice of the export list (elements)
  double* dst;                 // where it lands in the receiver's halo buffer (peer or local pointer)
  uint64_t* ready_flag;        // receiver's ready[me]
  const uint64_t* ack_flag;    // my ack[receiver]
};
struct P2PSource {
  const uint64_t* ready_flag;  // my ready[source]
  uint64_t* ack_flag;          // source's ack[me]
};

constexpr int kArenaFlagBytes = 4096;        // ready flags at [0, 2048), ack flags at [2048, 4096): 256 ranks max
constexpr unsigned long long kSpinTimeoutNs = 20ull * 1000 * 1000 * 1000;  // a peer that never shows up must not hang the GPU

__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
  return t;
}
// spin until *flag >= want; after kSpinTimeoutNs give up and raise the host-visible timeout flag (mapped pinned
// memory), which the host turns into an error at the next synchronising call -- never a silent wrong answer
__device__ __forceinline__ void wait_flag_ge(const uint64_t* flag, uint64_t want, unsigned long long* timeout_flag) {
  if (ld_acquire_sys(flag) >= want) return;
  const unsigned long long t0 = global_timer_ns();
  while (ld_acquire_sys(flag) < want) {
    if (global_timer_ns() - t0 > kSpinTimeoutNs) {
      atomicAdd_system(timeout_flag, 1ull);
      return;
    }
  }
}

// everything the fused apply kernel needs besides the SpMV arguments
struct FusedHalo {
  const int32_t* lids;      // export LIDs in send order
  int32_t n_export;
  const P2PSend* sends;
  int32_t n_sends;
  const P2PSource* sources;
  int32_t n_sources;
  uint64_t epoch;           // this apply's epoch (host counter of the plan: applies are collective and in order)
  unsigned int* tickets;    // [0] pack blocks done, [1] boundary blocks done (each re-armed by its last block)
  unsigned long long* timeout_flag;  // host-mapped; incremented when a flag wait gives up
};

}  // namespace jp
