// jp_mv_kernels.cuh -- fused MultiVector kernels of the CG building blocks, device code only, in
// the CUDA subset tests/cuda_emul can execute on the CPU.  The plain fill!/scale!/axpby/dot/norm kernels are in jp_mv.cu.
#pragma once
#include <cstdint>

namespace jp {
namespace mvk {

constexpr int kThreads = 256;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// y = (a*x) + (b*y) with the unfused arithmetic of the broadcast `@. y = a*x + b*y` (src/MultiVector.jl:497-502), and in
// the same pass out[vec] = sum_i y_new[i]^2 (the local part of norm(y, 2), src/MultiVector.jl:122-131): the residual
// update of a CG iteration and its norm cost 24 bytes per element instead of 24 + 8, and one launch instead of two.
// Grid (blocks, k); one partial per block and vector, added in block order by the last block of each vector (ticket
// counter, re-armed for the next call) -> deterministic for a given launch geometry.
__global__ void __launch_bounds__(kThreads) axpby_norm2_kernel(double* __restrict__ y, const double* __restrict__ x, int64_t n, double a,
                                                                 double b, double* __restrict__ partials, unsigned int* __restrict__ tickets,
                                                                 double* __restrict__ out) {
  const int vec = blockIdx.y;
  double* yv = y + (size_t)vec * n;
  const double* xv = x + (size_t)vec * n;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {  // eight independent loads in flight per thread
    const double x0 = xv[i], x1 = xv[i + stride], x2 = xv[i + 2 * stride], x3 = xv[i + 3 * stride];
    const double y0 = yv[i], y1 = yv[i + stride], y2 = yv[i + 2 * stride], y3 = yv[i + 3 * stride];
    const double r0 = __dadd_rn(__dmul_rn(a, x0), __dmul_rn(b, y0));
    const double r1 = __dadd_rn(__dmul_rn(a, x1), __dmul_rn(b, y1));
    const double r2 = __dadd_rn(__dmul_rn(a, x2), __dmul_rn(b, y2));
    const double r3 = __dadd_rn(__dmul_rn(a, x3), __dmul_rn(b, y3));
    yv[i] = r0;
    yv[i + stride] = r1;
    yv[i + 2 * stride] = r2;
    yv[i + 3 * stride] = r3;
    a0 += r0 * r0;
    a1 += r1 * r1;
    a2 += r2 * r2;
    a3 += r3 * r3;
  }
  for (; i < n; i += stride) {
    const double r0 = __dadd_rn(__dmul_rn(a, xv[i]), __dmul_rn(b, yv[i]));
    yv[i] = r0;
    a0 += r0 * r0;
  }
  const double s = warp_sum((a0 + a1) + (a2 + a3));
  __shared__ double warp_part[kThreads / 32];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_part[warp] = s;
  __syncthreads();
  if (warp == 0) {
    double t = lane < kThreads / 32 ? warp_part[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) {
      partials[(size_t)vec * gridDim.x + blockIdx.x] = t;
      __threadfence();
      is_last = atomicAdd(&tickets[vec], 1u) == gridDim.x - 1;
    }
  }
  __syncthreads();
  if (is_last && warp == 0) {
    __threadfence();
    double t = 0.0;
    const volatile double* pv = partials + (size_t)vec * gridDim.x;
    for (unsigned int blk = lane; blk < gridDim.x; blk += 32) t += pv[blk];
    t = warp_sum(t);
    if (lane == 0) {
      out[vec] = t;
      tickets[vec] = 0;
    }
  }
}

}  // namespace mvk
}  // namespace jp
