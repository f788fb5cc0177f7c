// jp_ptx.cuh -- the few PTX instructions the kernels use directly: cp.async, mbarrier, TMA bulk copies
// (cp.async.bulk, SASS UBLKCP) with an L2 evict-first policy.  Under JP_EMUL (tests/cuda_emul) the same functions are
// plain C++: copies complete at issue and the mbarrier is a small state machine with the PTX phase/parity semantics.
#pragma once
#include <cstdint>

namespace jp {

#ifndef JP_EMUL
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_addr(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n" ::: "memory");
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

// ---- mbarrier + TMA bulk copy (sm_90+ PTX; SASS: SYNCS / UBLKCP) ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_addr(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n" ::"r"(
                   smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(smem_addr(bar)), "l"(pol)
               : "memory");
}
// gather of one multivector entry by an irregular row: read-only path, no L1 allocation (no reuse inside the block) and an
// L2 evict-last hint -- x is what should stay in the L2, the matrix stream passes through with evict-first.  (Measured on
// the power-law matrix, 20 M rows: this form 6.71 ms; plain ld.global.nc 7.15; without the L2 hint 7.76; ld.global 7.16:
// profiles/r2_powerlaw20M_gather_load_ab.json)
__device__ __forceinline__ double ldg_gather(const double* p, uint64_t pol) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;\n" : "=d"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_addr(bar)) : "memory");
}
#else
// ---- CPU execution model ----
struct EmulMbar {  // lives in the 8 bytes of the barrier word
  uint16_t expected, pending;
  uint16_t phase;
  uint16_t pad;
  int32_t tx() const { return 0; }
};
static_assert(sizeof(EmulMbar) == 8, "mbarrier emulation must fit the barrier word");
// transaction bytes outstanding per barrier address (a side table: the 8-byte word has no room for a 32-bit count)
inline int64_t& emul_mbar_tx(uint64_t* bar) {
  static std::vector<std::pair<uint64_t*, int64_t>> table;
  for (auto& e : table)
    if (e.first == bar) return e.second;
  table.emplace_back(bar, 0);
  return table.back().second;
}
inline void emul_mbar_check(uint64_t* bar) {
  EmulMbar* m = reinterpret_cast<EmulMbar*>(bar);
  if (m->pending == 0 && emul_mbar_tx(bar) == 0) {
    m->phase ^= 1;
    m->pending = m->expected;
    ++::emul::state().progress;
  }
}
inline void cp_async16(void* smem, const void* gmem) { std::memcpy(smem, gmem, 16); }
inline void cp_async_wait_all() {}
inline void mbar_init(uint64_t* bar, unsigned count) {
  EmulMbar* m = reinterpret_cast<EmulMbar*>(bar);
  m->expected = m->pending = (uint16_t)count;
  m->phase = 0;
  m->pad = 0;
  emul_mbar_tx(bar) = 0;
}
inline void mbar_expect_tx(uint64_t* bar, unsigned bytes) {  // mbarrier.arrive.expect_tx
  EmulMbar* m = reinterpret_cast<EmulMbar*>(bar);
  emul_mbar_tx(bar) += bytes;
  --m->pending;
  emul_mbar_check(bar);
}
inline void mbar_arrive(uint64_t* bar) {
  EmulMbar* m = reinterpret_cast<EmulMbar*>(bar);
  --m->pending;
  emul_mbar_check(bar);
}
inline void mbar_wait(uint64_t* bar, unsigned parity) {  // returns once the phase with this parity has completed
  const EmulMbar* m = reinterpret_cast<const EmulMbar*>(bar);
  while (m->phase == (parity & 1)) ::emul::yield();
}
inline uint64_t policy_evict_first() { return 0; }
inline uint64_t policy_evict_last() { return 0; }
inline double ldg_gather(const double* p, uint64_t) { return *p; }
inline void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar, uint64_t) {
  if ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src) | bytes) & 15) {
    std::fprintf(stderr, "cuda_emul: cp.async.bulk needs 16-byte aligned addresses and size (dst %p src %p bytes %u)\n", dst, src, bytes);
    std::abort();
  }
  std::memcpy(dst, src, bytes);
  emul_mbar_tx(bar) -= bytes;
  emul_mbar_check(bar);
}
#endif

}  // namespace jp
