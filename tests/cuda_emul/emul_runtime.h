// emul_runtime.h -- the slice of csrc/jp_common.h and of the CUDA runtime API that the host pipelines
// (csrc/*_pipeline.h) use, backed by mmap/memcpy, for the CPU execution model of cuda_emul.h.
// TEST INFRASTRUCTURE ONLY.  "Device" buffers end at an inaccessible page and carry a canary in front; fresh
// allocations are filled with garbage like cudaMalloc'ed memory.
#pragma once
#include <sys/mman.h>
#include <unistd.h>

#include <algorithm>
#include <string>
#include <vector>

#include "cuda_emul.h"

typedef void* cudaStream_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
inline int cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) {
  std::memcpy(d, s, n);
  return 0;
}
inline int cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) {
  std::memcpy(d, s, n);
  return 0;
}
inline int cudaMemset(void* d, int v, size_t n) {
  std::memset(d, v, n);
  return 0;
}
inline int cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) {
  std::memset(d, v, n);
  return 0;
}
inline int cudaStreamSynchronize(cudaStream_t) { return 0; }

enum { JP_OK = 0, JP_INVALID_ARGUMENT = 1, JP_INVALID_STATE = 2, JP_CUDA_ERROR = 3, JP_NCCL_ERROR = 4 };

namespace jp {

struct Error {
  int32_t code;
  std::string msg;
};
[[noreturn]] inline void fail(int32_t code, const std::string& m) { throw Error{code, m}; }

#define JP_CUDA(call) (void)(call)
#define JP_REQUIRE(cond, msg)                            \
  do {                                                   \
    if (!(cond)) ::jp::fail(JP_INVALID_ARGUMENT, (msg)); \
  } while (0)
#define JP_REQUIRE_STATE(cond, msg)                   \
  do {                                                \
    if (!(cond)) ::jp::fail(JP_INVALID_STATE, (msg)); \
  } while (0)
#define JP_LAUNCH_CHECK() \
  do {                    \
  } while (0)
#define JP_KLAUNCH(kernel, grid, block, smem, stream, ...) ::emul::launch(dim3(grid), dim3(block), (smem), [&] { kernel(__VA_ARGS__); })

struct Context {
  int sm_count = 2;  // small, so grid-stride loops actually stride in the tests
  cudaStream_t compute = nullptr;
};
inline Context& ctx() {
  static Context c;
  return c;
}

constexpr size_t kGuard = 256;
constexpr unsigned char kGuardByte = 0xA7;

// "Device" memory: every buffer ends (up to 16-byte alignment) at a PROT_NONE page, so an access past the end of an
// allocation faults on the spot -- reads included -- and a canary in front of it catches stray writes before the start.
struct DeviceBuffer {
  void* p = nullptr;
  size_t bytes = 0;
  DeviceBuffer() {}
  DeviceBuffer(const DeviceBuffer&) = delete;
  DeviceBuffer& operator=(const DeviceBuffer&) = delete;
  ~DeviceBuffer() { release(); }
  static size_t page() { return (size_t)sysconf(_SC_PAGESIZE); }
  void alloc(size_t n) {
    release();
    if (n == 0) n = 16;
    const size_t pg = page();
    const size_t padded = (n + 15) & ~(size_t)15;                       // keeps the start 16-byte aligned
    const size_t body = (kGuard + padded + pg - 1) / pg * pg;           // pages holding [canary | buffer]
    unsigned char* raw = (unsigned char*)mmap(nullptr, body + pg, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (raw == (unsigned char*)MAP_FAILED) {
      std::fprintf(stderr, "cuda_emul: mmap of %zu bytes failed\n", body + pg);
      std::abort();
    }
    mprotect(raw + body, pg, PROT_NONE);                                 // the page right behind the buffer
    unsigned char* start = raw + body - padded;
    std::memset(start - kGuard, kGuardByte, kGuard);
    std::memset(start, 0xCD, padded);                                    // cudaMalloc'ed memory is not zero
    p = start;
    bytes = n;
    map_base = raw;
    map_bytes = body + pg;
  }
  void ensure(size_t n) {
    if (n > bytes) alloc(n);
  }
  void check_guards() const {
    if (!p) return;
    const unsigned char* g = (const unsigned char*)p - kGuard;
    for (size_t i = 0; i < kGuard; ++i)
      if (g[i] != kGuardByte) {
        std::fprintf(stderr, "cuda_emul: out-of-bounds write in front of a %zu-byte device buffer\n", bytes);
        std::abort();
      }
  }
  void release() {
    if (p) {
      check_guards();
      munmap(map_base, map_bytes);
    }
    p = nullptr;
    bytes = 0;
    map_base = nullptr;
    map_bytes = 0;
  }
  void swap(DeviceBuffer& o) {
    std::swap(p, o.p);
    std::swap(bytes, o.bytes);
    std::swap(map_base, o.map_base);
    std::swap(map_bytes, o.map_bytes);
  }
  template <class T>
  T* as() const { return static_cast<T*>(p); }

 private:
  void* map_base = nullptr;
  size_t map_bytes = 0;
};

}  // namespace jp
