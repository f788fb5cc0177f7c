// jp_plan_kernels.cuh -- the data-movement kernels of doTransfer (src/DistObject.jl:200-258 -> src/MultiVector.jl:308-371)
// and of the ADD transfers behind the transposed / exporter branches of apply! (src/RowMatrix.jl:298-352).  Device code
// only, in the CUDA subset tests/cuda_emul can execute on the CPU; the orchestration is in jp_plan.cu.
#pragma once
#include <cstdint>

namespace jp {
namespace plank {

constexpr int kThreads = 256;

// sendbuf[i*k + j] = src[lid[i] + j*ld]      (src/MultiVector.jl:339-347)
__global__ void __launch_bounds__(kThreads) pack_kernel(const double* __restrict__ src, int64_t ld, const int32_t* __restrict__ lids,
                                                          int32_t n, int32_t k, double* __restrict__ buf) {
  const int64_t total = (int64_t)n * k;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += stride) {
    const int32_t j = (int32_t)(t / n), i = (int32_t)(t - (int64_t)j * n);  // consecutive threads: consecutive LIDs
    buf[(size_t)i * k + j] = __ldg(src + __ldg(lids + i) + (size_t)j * ld);
  }
}

// tgt[lid[i] + j*ld] = buf[i*k + j]          (src/MultiVector.jl:363-369; combine mode ignored, as there)
__global__ void __launch_bounds__(kThreads) unpack_kernel(double* __restrict__ tgt, int64_t ld, const int32_t* __restrict__ lids,
                                                            int32_t n, int32_t k, const double* __restrict__ buf) {
  const int64_t total = (int64_t)n * k;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += stride) {
    const int32_t j = (int32_t)(t / n), i = (int32_t)(t - (int64_t)j * n);
    tgt[__ldg(lids + i) + (size_t)j * ld] = __ldg(buf + (size_t)i * k + j);
  }
}

// tgt[to[i] + j*ldt] = src[from[i] + j*lds]  (src/MultiVector.jl:320-323)
__global__ void __launch_bounds__(kThreads) permute_kernel(double* __restrict__ tgt, int64_t ldt, const double* __restrict__ src,
                                                             int64_t lds, const int32_t* __restrict__ to, const int32_t* __restrict__ from,
                                                             int32_t n, int32_t k) {
  const int64_t total = (int64_t)n * k;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += stride) {
    const int32_t j = (int32_t)(t / n), i = (int32_t)(t - (int64_t)j * n);
    tgt[__ldg(to + i) + (size_t)j * ldt] = __ldg(src + __ldg(from + i) + (size_t)j * lds);
  }
}

// ---- ADD transfers (Tpetra semantics, see apply_trans_distributed / apply_with_exporter in jp_plan.cu) ----
__global__ void __launch_bounds__(kThreads) add_rows_kernel(double* __restrict__ y, int64_t ldy, const double* __restrict__ yc, int64_t ldc,
                                                              int32_t n, int32_t k) {
  const int64_t total = (int64_t)n * k, stride = (int64_t)gridDim.x * kThreads;
  for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += stride) {
    const int32_t j = (int32_t)(t / n), i = (int32_t)(t - (int64_t)j * n);
    y[i + (size_t)j * ldy] += yc[i + (size_t)j * ldc];
  }
}
__global__ void __launch_bounds__(kThreads) add_permute_kernel(double* __restrict__ y, int64_t ldy, const double* __restrict__ yc, int64_t ldc,
                                                                 const int32_t* __restrict__ dom_lids, const int32_t* __restrict__ col_lids,
                                                                 int32_t n, int32_t k) {
  const int64_t total = (int64_t)n * k, stride = (int64_t)gridDim.x * kThreads;
  for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += stride) {
    const int32_t j = (int32_t)(t / n), i = (int32_t)(t - (int64_t)j * n);
    y[dom_lids[i] + (size_t)j * ldy] += yc[col_lids[i] + (size_t)j * ldc];
  }
}
// y[lid[i] + j*ld] += buf[i*k + j]; a row exported to several ranks appears several times in lids -> atomics
__global__ void __launch_bounds__(kThreads) scatter_add_kernel(double* __restrict__ y, int64_t ld, const int32_t* __restrict__ lids, int32_t n,
                                                                 int32_t k, const double* __restrict__ buf) {
  const int64_t total = (int64_t)n * k, stride = (int64_t)gridDim.x * kThreads;
  for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += stride) {
    const int32_t j = (int32_t)(t / n), i = (int32_t)(t - (int64_t)j * n);
    atomicAdd(y + lids[i] + (size_t)j * ld, buf[(size_t)i * k + j]);
  }
}


}  // namespace plank
}  // namespace jp
