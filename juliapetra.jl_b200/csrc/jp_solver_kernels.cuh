// jp_solver_kernels.cuh -- update kernels whose coefficients live on the device (jp_solver.cu).  Device code only, in the
// CUDA subset tests/cuda_emul can execute on the CPU.  Included by exactly one .cu file (the kernels are not templates).
#pragma once
#include <cstdint>

namespace jp {
namespace solk {

constexpr int kThreads = 256;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- device-resident scalars (jp_solver.cu): a coefficient is  scale * s[num] / s[den]  evaluated on the device, so the
//      alpha = rr / p'Ap of a CG iteration never travels to the host.  num / den < 0 stand for 1. ----
struct Coeff {
  double scale;
  int32_t num, den;
};
__device__ __forceinline__ double coeff_value(const Coeff& c, const double* __restrict__ s) {
  double v = c.scale;
  if (c.num >= 0) v = __dmul_rn(v, s[c.num]);
  if (c.den >= 0) v = v / s[c.den];
  return v;
}

// y = (a*x) + (b*y) with a, b device-resident coefficients; the same unfused arithmetic as axpby_kernel
__global__ void __launch_bounds__(kThreads) axpby_dev_kernel(double* __restrict__ y, const double* __restrict__ x, int64_t n, Coeff ca, Coeff cb,
                                                               const double* __restrict__ s) {
  const double a = coeff_value(ca, s), b = coeff_value(cb, s);
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride)
    y[i] = __dadd_rn(__dmul_rn(a, x[i]), __dmul_rn(b, y[i]));
}

// axpby_norm2_kernel (one vector) with device-resident coefficients; the sum of squares lands in out[0]
__global__ void __launch_bounds__(kThreads) axpby_norm2_dev_kernel(double* __restrict__ y, const double* __restrict__ x, int64_t n, Coeff ca,
                                                                     Coeff cb, const double* __restrict__ s, double* __restrict__ partials,
                                                                     unsigned int* __restrict__ tickets, double* __restrict__ out) {
  const double a = coeff_value(ca, s), b = coeff_value(cb, s);
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    const double x0 = x[i], x1 = x[i + stride], x2 = x[i + 2 * stride], x3 = x[i + 3 * stride];
    const double y0 = y[i], y1 = y[i + stride], y2 = y[i + 2 * stride], y3 = y[i + 3 * stride];
    const double r0 = __dadd_rn(__dmul_rn(a, x0), __dmul_rn(b, y0));
    const double r1 = __dadd_rn(__dmul_rn(a, x1), __dmul_rn(b, y1));
    const double r2 = __dadd_rn(__dmul_rn(a, x2), __dmul_rn(b, y2));
    const double r3 = __dadd_rn(__dmul_rn(a, x3), __dmul_rn(b, y3));
    y[i] = r0;
    y[i + stride] = r1;
    y[i + 2 * stride] = r2;
    y[i + 3 * stride] = r3;
    a0 += r0 * r0;
    a1 += r1 * r1;
    a2 += r2 * r2;
    a3 += r3 * r3;
  }
  for (; i < n; i += stride) {
    const double r0 = __dadd_rn(__dmul_rn(a, x[i]), __dmul_rn(b, y[i]));
    y[i] = r0;
    a0 += r0 * r0;
  }
  const double sum = warp_sum((a0 + a1) + (a2 + a3));
  __shared__ double warp_part[kThreads / 32];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_part[warp] = sum;
  __syncthreads();
  if (warp == 0) {
    double t = lane < kThreads / 32 ? warp_part[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) {
      partials[blockIdx.x] = t;
      __threadfence();
      is_last = atomicAdd(&tickets[0], 1u) == gridDim.x - 1;
    }
  }
  __syncthreads();
  if (is_last) {  // block-uniform: every thread of the last block takes part in the fixed-order final sum
    __threadfence();
    const volatile double* pv = partials;
    double acc = 0.0;
    for (unsigned int base = 0; base < gridDim.x; base += 4 * kThreads) {
      const unsigned int i0 = base + threadIdx.x, i1 = i0 + kThreads, i2 = i1 + kThreads, i3 = i2 + kThreads;
      const double v0 = i0 < gridDim.x ? pv[i0] : 0.0, v1 = i1 < gridDim.x ? pv[i1] : 0.0;
      const double v2 = i2 < gridDim.x ? pv[i2] : 0.0, v3 = i3 < gridDim.x ? pv[i3] : 0.0;
      acc += (v0 + v1) + (v2 + v3);
    }
    const double ws = warp_sum(acc);
    __syncthreads();
    if (lane == 0) warp_part[warp] = ws;
    __syncthreads();
    if (warp == 0) {
      double t = lane < kThreads / 32 ? warp_part[lane] : 0.0;
      t = warp_sum(t);
      if (lane == 0) {
        out[0] = t;
        tickets[0] = 0;
      }
    }
  }
}

}  // namespace solk
}  // namespace jp
