// emul_runtime.h -- the slice of csrc/jp_common.h and of the CUDA runtime API that the host pipelines
// (csrc/*_pipeline.h) use, backed by malloc/memcpy, for the CPU execution model of cuda_emul.h.
// TEST INFRASTRUCTURE ONLY.  "Device" buffers carry guard zones that are checked on release, and fresh allocations
// are filled with garbage like cudaMalloc'ed memory.
#pragma once
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

struct DeviceBuffer {
  void* p = nullptr;
  size_t bytes = 0;
  DeviceBuffer() {}
  DeviceBuffer(const DeviceBuffer&) = delete;
  DeviceBuffer& operator=(const DeviceBuffer&) = delete;
  ~DeviceBuffer() { release(); }
  void alloc(size_t n) {
    release();
    if (n == 0) n = 16;
    unsigned char* raw = (unsigned char*)std::malloc(n + 2 * kGuard);
    std::memset(raw, kGuardByte, kGuard);
    std::memset(raw + kGuard, 0xCD, n);
    std::memset(raw + kGuard + n, kGuardByte, kGuard);
    p = raw + kGuard;
    bytes = n;
  }
  void ensure(size_t n) {
    if (n > bytes) alloc(n);
  }
  void check_guards() const {
    if (!p) return;
    const unsigned char* raw = (const unsigned char*)p - kGuard;
    for (size_t i = 0; i < kGuard; ++i)
      if (raw[i] != kGuardByte || raw[kGuard + bytes + i] != kGuardByte) {
        std::fprintf(stderr, "cuda_emul: out-of-bounds write next to a %zu-byte device buffer\n", bytes);
        std::abort();
      }
  }
  void release() {
    if (p) {
      check_guards();
      std::free((unsigned char*)p - kGuard);
    }
    p = nullptr;
    bytes = 0;
  }
  void swap(DeviceBuffer& o) {
    std::swap(p, o.p);
    std::swap(bytes, o.bytes);
  }
  template <class T>
  T* as() const { return static_cast<T*>(p); }
};

}  // namespace jp
