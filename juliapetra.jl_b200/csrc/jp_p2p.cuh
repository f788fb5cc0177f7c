// jp_p2p.cuh -- peer-to-peer halo protocol shared by jp_plan.cu (setup, two-stream path) and jp_csr.cu (fused kernel)
//
// Every rank exposes one ARENA through CUDA IPC:
//   [ ready flags: uint64 per source rank | ack flags: uint64 per destination rank | halo buffer 0 | halo buffer 1 ]
// Apply number e (the plan's EPOCH, identical on every rank because applies are collective and ordered) uses halo
// buffer e & 1 on every rank:
//   sender    waits  ack[dst] >= e - 2   (the receiver has consumed what this buffer held two applies ago),
//             stores its exported rows into the receiver's buffer e & 1 over NVLink, then ready[dst][me] := e (release.sys)
//   receiver  waits  ready[me][src] >= e (acquire.sys), reads buffer e & 1, then ack[src][me] := e
// With two buffers a rank may run one whole apply ahead of a neighbour before it stalls, so neighbours are not
// lock-stepped by the back-pressure (tests/test_p2p_protocol_model.py explores the protocol under random schedules).
// Offsets inside a buffer are kept in ELEMENTS of one vector and scaled by k in the kernels, so one arena (sized for
// the largest k seen) serves every k.
#pragma once
#include <cstdint>

namespace jp {

struct P2PSend {
  int32_t start, count;        // slice of the export list (elements)
  double* dst_base;            // the receiver's halo buffer 0 (peer or local pointer)
  int64_t dst_buf_elems;       // doubles between the receiver's buffer 0 and buffer 1
  int64_t dst_off;             // where this rank's rows start in the receiver's buffer, in rows (scaled by k)
  uint64_t* ready_flag;        // receiver's ready[me]
  const uint64_t* ack_flag;    // my ack[receiver]
};
struct P2PSource {
  const uint64_t* ready_flag;  // my ready[source]
  uint64_t* ack_flag;          // source's ack[me]
};

constexpr int kArenaFlagBytes = 4096;        // ready flags at [0, 2048), ack flags at [2048, 4096): 256 ranks max
constexpr int kP2PMaxRanks = 256, kP2PMaxSends = 256, kP2PMaxSources = 32;
constexpr unsigned long long kSpinTimeoutNs = 20ull * 1000 * 1000 * 1000;  // a peer that never shows up must not hang the GPU

#ifdef __CUDACC__
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
// memory), which the host turns into an error at the next synchronising call -- never a silent wrong answer.  Once the
// flag is up every later wait gives up at once, so a rank whose neighbour has left (a non-collective call sequence: a
// usage error) drains its queued applies in seconds instead of 20 s per apply.
__device__ __forceinline__ void wait_flag_ge(const uint64_t* flag, uint64_t want, unsigned long long* timeout_flag) {
  if (ld_acquire_sys(flag) >= want) return;
  const unsigned long long t0 = global_timer_ns();
  unsigned polls = 0;
  while (ld_acquire_sys(flag) < want) {
    if ((++polls & 0x3ff) == 0) {
      if (*reinterpret_cast<volatile unsigned long long*>(timeout_flag) != 0) return;
      if (global_timer_ns() - t0 > kSpinTimeoutNs) {
        atomicAdd_system(timeout_flag, 1ull);
        return;
      }
    }
  }
}
#endif

// everything the fused apply kernel needs besides the SpMV arguments
struct FusedHalo {
  const int32_t* lids;      // export LIDs in send order
  int32_t n_export;
  const P2PSend* sends;
  int32_t n_sends;
  const P2PSource* sources;
  int32_t n_sources;
  uint64_t epoch;           // this apply's epoch (host counter of the plan: applies are collective and in order)
  unsigned int* tickets;    // [0] pack items done, [1] boundary blocks done, [2] pack-item claims (each re-armed by its last user)
  unsigned long long* timeout_flag;  // host-mapped; incremented when a flag wait gives up
};

}  // namespace jp
