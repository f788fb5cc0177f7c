// jp_mv_kernels.cuh -- the MultiVector kernels (src/MultiVector.jl:52-161, 497-502): fill!, scale!, axpby, dot, norm and
// the fused axpby+norm2 of the CG building blocks.  Device code only, in the CUDA subset
// tests/cuda_emul can execute on the CPU; launches and the reduction workspace are in jp_mv.cu (see its header).
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

// Sum of pv[0..n) by ALL threads of the block in a fixed order (the tail of the reduction kernels, run by the last block to
// finish): four independent loads per thread at a time, then the usual warp / block tree.  The result is valid in thread 0.
// (One warp walking the <= 1024 partials one dependent L2 load after the other cost ~15 us per dot / norm call.)
__device__ __forceinline__ double block_sum_fixed_order(const volatile double* pv, unsigned n, double* warp_part) {
  const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double a = 0.0;
  for (unsigned base = 0; base < n; base += 4 * kThreads) {
    const unsigned i0 = base + tid, i1 = i0 + kThreads, i2 = i1 + kThreads, i3 = i2 + kThreads;
    const double v0 = i0 < n ? pv[i0] : 0.0, v1 = i1 < n ? pv[i1] : 0.0, v2 = i2 < n ? pv[i2] : 0.0, v3 = i3 < n ? pv[i3] : 0.0;
    a += (v0 + v1) + (v2 + v3);
  }
  const double t = warp_sum(a);
  __syncthreads();  // warp_part may still be read by warp 0 of the caller
  if (lane == 0) warp_part[warp] = t;
  __syncthreads();
  double u = 0.0;
  if (warp == 0) {
    u = lane < kThreads / 32 ? warp_part[lane] : 0.0;
    u = warp_sum(u);
  }
  return u;
}

// How the k results of a reduction reach the host: one warp stores them straight into host-mapped pinned memory and
// then a sequence number the host spins on.  This replaces the cudaMemcpyAsync + cudaStreamSynchronize pair (~20 us of
// fixed cost per dot / norm call); with several ranks it runs behind the NCCL allreduce on the same stream.
__global__ void publish_results_kernel(const double* __restrict__ dev, int k, double* host_out, unsigned long long* host_flag,
                                       unsigned long long seq) {
  for (int v = threadIdx.x; v < k; v += 32) *reinterpret_cast<volatile double*>(host_out + v) = dev[v];
  __threadfence_system();
  __syncwarp();
  if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long*>(host_flag) = seq;
}

template <int MODE>
__device__ __forceinline__ double reduce_term(double x, double y, double p) {
  if (MODE == 0) return x * y;
  if (MODE == 1) return x * x;
  return pow(x, p);
}

// MODE 0: dot  sum x*y      (src/MultiVector.jl:99-105)
// MODE 1: norm2 sum x*x     (src/MultiVector.jl:122-131)
// MODE 2: normp sum x^p     (src/MultiVector.jl:135-142)
// Streaming, HBM-bound: 16-byte loads, eight of them in flight per thread (four per stream for dot).  A column that
// starts on an odd double (odd leading dimension) is peeled by one element so that the vector loads stay aligned.
template <int MODE>
__global__ void __launch_bounds__(kThreads) reduce_kernel(const double* __restrict__ x, const double* __restrict__ y, int64_t n,
                                                            double p, double* __restrict__ partials,
                                                            unsigned int* __restrict__ tickets, double* __restrict__ out) {
  const int vec = blockIdx.y;
  const double* xv = x + (size_t)vec * n;
  const double* yv = MODE == 0 ? y + (size_t)vec * n : xv;
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  const int64_t gtid = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const int64_t head = n > 0 ? (int64_t)((reinterpret_cast<uintptr_t>(xv) >> 3) & 1) : 0;
  const bool vec16 = ((reinterpret_cast<uintptr_t>(yv) >> 3) & 1) == (uintptr_t)head;  // both streams 16-byte aligned after the peel
  if (vec16) {
    const double2* x2 = reinterpret_cast<const double2*>(xv + head);
    const double2* y2 = reinterpret_cast<const double2*>(yv + head);
    const int64_t n2 = (n - head) >> 1;
    int64_t i = gtid;
    if (MODE == 0) {
      for (; i + 3 * stride < n2; i += 4 * stride) {
        const double2 p0 = x2[i], p1 = x2[i + stride], p2 = x2[i + 2 * stride], p3 = x2[i + 3 * stride];
        const double2 q0 = y2[i], q1 = y2[i + stride], q2 = y2[i + 2 * stride], q3 = y2[i + 3 * stride];
        a0 += p0.x * q0.x; a1 += p0.y * q0.y; a2 += p1.x * q1.x; a3 += p1.y * q1.y;
        a0 += p2.x * q2.x; a1 += p2.y * q2.y; a2 += p3.x * q3.x; a3 += p3.y * q3.y;
      }
    } else {
      for (; i + 7 * stride < n2; i += 8 * stride) {
        const double2 p0 = x2[i], p1 = x2[i + stride], p2 = x2[i + 2 * stride], p3 = x2[i + 3 * stride];
        const double2 p4 = x2[i + 4 * stride], p5 = x2[i + 5 * stride], p6 = x2[i + 6 * stride], p7 = x2[i + 7 * stride];
        a0 += reduce_term<MODE>(p0.x, 0, p); a1 += reduce_term<MODE>(p0.y, 0, p); a2 += reduce_term<MODE>(p1.x, 0, p); a3 += reduce_term<MODE>(p1.y, 0, p);
        a0 += reduce_term<MODE>(p2.x, 0, p); a1 += reduce_term<MODE>(p2.y, 0, p); a2 += reduce_term<MODE>(p3.x, 0, p); a3 += reduce_term<MODE>(p3.y, 0, p);
        a0 += reduce_term<MODE>(p4.x, 0, p); a1 += reduce_term<MODE>(p4.y, 0, p); a2 += reduce_term<MODE>(p5.x, 0, p); a3 += reduce_term<MODE>(p5.y, 0, p);
        a0 += reduce_term<MODE>(p6.x, 0, p); a1 += reduce_term<MODE>(p6.y, 0, p); a2 += reduce_term<MODE>(p7.x, 0, p); a3 += reduce_term<MODE>(p7.y, 0, p);
      }
    }
    for (; i < n2; i += stride) {
      const double2 p0 = x2[i];
      const double2 q0 = MODE == 0 ? y2[i] : p0;
      a0 += reduce_term<MODE>(p0.x, q0.x, p);
      a1 += reduce_term<MODE>(p0.y, q0.y, p);
    }
    if (gtid == 0) {  // the peeled head and the odd tail
      if (head) a2 += reduce_term<MODE>(xv[0], yv[0], p);
      if ((n - head) & 1) a3 += reduce_term<MODE>(xv[n - 1], yv[n - 1], p);
    }
  } else {  // the two streams are misaligned against each other: 8-byte loads
    int64_t i = gtid;
    for (; i + 3 * stride < n; i += 4 * stride) {
      const double x0 = xv[i], x1 = xv[i + stride], x2s = xv[i + 2 * stride], x3 = xv[i + 3 * stride];
      const double y0 = yv[i], y1 = yv[i + stride], y2s = yv[i + 2 * stride], y3 = yv[i + 3 * stride];
      a0 += reduce_term<MODE>(x0, y0, p); a1 += reduce_term<MODE>(x1, y1, p); a2 += reduce_term<MODE>(x2s, y2s, p); a3 += reduce_term<MODE>(x3, y3, p);
    }
    for (; i < n; i += stride) a0 += reduce_term<MODE>(xv[i], yv[i], p);
  }
  double s = warp_sum((a0 + a1) + (a2 + a3));
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
      unsigned int ticket = atomicAdd(&tickets[vec], 1u);
      is_last = (ticket == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (is_last) {  // block-uniform
    __threadfence();
    // fixed-order final sum over the blocks of this vector
    const double t = block_sum_fixed_order(partials + (size_t)vec * gridDim.x, gridDim.x, warp_part);
    if (threadIdx.x == 0) {
      out[vec] = t;
      tickets[vec] = 0;  // re-arm for the next call
    }
  }
}

__global__ void __launch_bounds__(kThreads) fill_kernel(double* __restrict__ y, int64_t n, double v) {
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) y[i] = v;
}

// y *= alpha  (rmul!, src/MultiVector.jl:63)
__global__ void __launch_bounds__(kThreads) scale_kernel(double* __restrict__ y, int64_t n, double alpha) {
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {  // four independent loads in flight per thread
    const double a0 = y[i], a1 = y[i + stride], a2 = y[i + 2 * stride], a3 = y[i + 3 * stride];
    y[i] = __dmul_rn(a0, alpha);
    y[i + stride] = __dmul_rn(a1, alpha);
    y[i + 2 * stride] = __dmul_rn(a2, alpha);
    y[i + 3 * stride] = __dmul_rn(a3, alpha);
  }
  for (; i < n; i += stride) y[i] = __dmul_rn(y[i], alpha);
}

// column v scaled by alpha[v]  (src/MultiVector.jl:72-77)
__global__ void __launch_bounds__(kThreads) scale_vec_kernel(double* __restrict__ y, int64_t n, const double* __restrict__ alpha) {
  const double a = alpha[blockIdx.y];
  double* col = y + (size_t)blockIdx.y * n;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    const double a0 = col[i], a1 = col[i + stride], a2 = col[i + 2 * stride], a3 = col[i + 3 * stride];
    col[i] = __dmul_rn(a0, a);
    col[i + stride] = __dmul_rn(a1, a);
    col[i + 2 * stride] = __dmul_rn(a2, a);
    col[i + 3 * stride] = __dmul_rn(a3, a);
  }
  for (; i < n; i += stride) col[i] = __dmul_rn(col[i], a);
}

// y = (a*x) + (b*y), unfused  (broadcast copyto!, src/MultiVector.jl:497-502)
__global__ void __launch_bounds__(kThreads) axpby_kernel(double* __restrict__ y, const double* __restrict__ x, int64_t n, double a,
                                                           double b) {
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride)
    y[i] = __dadd_rn(__dmul_rn(a, x[i]), __dmul_rn(b, y[i]));
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
  if (is_last) {  // block-uniform
    __threadfence();
    const double t = block_sum_fixed_order(partials + (size_t)vec * gridDim.x, gridDim.x, warp_part);
    if (threadIdx.x == 0) {
      out[vec] = t;
      tickets[vec] = 0;
    }
  }
}

}  // namespace mvk
}  // namespace jp
