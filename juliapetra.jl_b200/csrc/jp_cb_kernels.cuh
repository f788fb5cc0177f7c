// jp_cb_kernels.cuh -- SpMV (k = 1) on a large irregular matrix through COLUMN SLABS (jp_colblock.cu).
//
// Why (profiles/r2_ncu_spmv_powerlaw50M_summary.txt): on the 50 M-row power-law matrix of BASELINE configs[3] the
// multivector (400 MB) fits neither the L2 nor the TLBs' reach, every one of the 762 M gathers of the row-block kernel is
// its own memory transaction and costs ~100 bytes of DRAM traffic: 81.7 GB per apply against 10.6 GB algorithmic, and the
// kernel runs at the DRAM's random-access throughput.  The classic fix is cache blocking: cut the columns into slabs whose
// piece of x stays resident in the L2 (32 MB by default), and sweep the matrix slab by slab.
//
// Layout (built once, at the first k = 1 apply; 16 bytes per nonzero on top of the CSR): the nonzeros in SLAB-MAJOR order
// -- a stable sort of the CSR positions by slab id, so that inside a slab they are ordered by row, then by the CSR order of
// the row -- as three arrays  row int32 | col int32 | val f64.
// Kernel (one launch per slab, stream-ordered): a warp takes 32 consecutive nonzeros per pass (coalesced 4 + 4 + 8 byte
// loads), gathers x[col] (an L2 hit), and adds up runs of equal rows with a segmented warp scan (rows ascend inside a slab,
// so a run is a contiguous lane range).  A run that lies strictly inside the pass belongs to this warp alone:
// y[row] += alpha * sum, a plain read-modify-write.  The first and the last run of a pass may continue in the
// neighbouring passes, so they go to a carry array (two entries per pass) and a second small kernel adds runs of equal
// rows among the carries in their order and applies them -- no atomics, a fixed summation order, deterministic.
// y is scaled by beta once, before the first slab.
#pragma once
#include <cstdint>

#include "jp_ptx.cuh"

namespace jp {
namespace cbk {

constexpr int kCbThreads = 256;
constexpr int kCbPasses = 4;                                        // passes of 32 nonzeros per warp
constexpr int kCbPerBlock = kCbThreads * kCbPasses;                 // nonzeros per CTA

struct CbCarry {
  double sum;
  int32_t row;   // -1: no entry
  int32_t pad;
};

// nonzeros [e0, e1) of one slab; carries of pass q (global pass index within the slab) at carry[2q], carry[2q + 1]
__global__ void __launch_bounds__(kCbThreads) spmv_cb_kernel(const int32_t* __restrict__ row, const int32_t* __restrict__ col,
                                                               const double* __restrict__ val, int64_t e0, int64_t e1,
                                                               const double* __restrict__ x, double* __restrict__ y, double alpha,
                                                               CbCarry* __restrict__ carry) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t pol_x = policy_evict_last();
  const int64_t warp_base = e0 + ((int64_t)blockIdx.x * (kCbThreads / 32) + warp) * (32 * kCbPasses);
  int32_t r[kCbPasses], c[kCbPasses];
  double v[kCbPasses], p[kCbPasses];
#pragma unroll
  for (int q = 0; q < kCbPasses; ++q) {  // all loads of the warp's four passes are issued before the first is needed
    const int64_t i = warp_base + q * 32 + lane;
    const bool in = i < e1;
    r[q] = in ? __ldcs(row + i) : -1;
    c[q] = in ? __ldcs(col + i) : 0;
    v[q] = in ? __ldcs(val + i) : 0.0;
  }
#pragma unroll
  for (int q = 0; q < kCbPasses; ++q) p[q] = r[q] >= 0 ? v[q] * ldg_gather(x + c[q], pol_x) : 0.0;
#pragma unroll
  for (int q = 0; q < kCbPasses; ++q) {
    const int64_t pass_first = warp_base + q * 32;
    if (pass_first >= e1) break;  // warp-uniform
    const int64_t pass_index = (pass_first - e0) >> 5;
    const int32_t rq = r[q];
    double s = p[q];
    // inclusive segmented scan: after it, the last lane of a run holds the run's sum (left to right in lane order)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, s, o);
      const int32_t r2 = __shfl_up_sync(0xffffffffu, rq, o);
      if (lane >= o && r2 == rq) s += t;
    }
    const int32_t r_next = __shfl_down_sync(0xffffffffu, rq, 1);
    const int32_t r_first = __shfl_sync(0xffffffffu, rq, 0);
    // the last valid lane of the pass (padding lanes carry row -1 and sit at the end)
    const unsigned valid = __ballot_sync(0xffffffffu, rq >= 0);
    const int last_lane = 31 - __clz((int)valid);
    const int32_t r_last = __shfl_sync(0xffffffffu, rq, last_lane);
    const bool tail = rq >= 0 && (lane == last_lane || r_next != rq);
    if (tail) {
      if (rq == r_first) {
        carry[2 * pass_index].sum = s;
        carry[2 * pass_index].row = rq;
        if (r_first == r_last) {  // the whole pass is one run: the second carry slot repeats the row with nothing to add
          carry[2 * pass_index + 1].sum = 0.0;
          carry[2 * pass_index + 1].row = rq;
        }
      } else if (rq == r_last) {
        carry[2 * pass_index + 1].sum = s;
        carry[2 * pass_index + 1].row = rq;
      } else {
        y[rq] += alpha * s;  // a run strictly inside the pass: no other pass, warp or CTA touches this row in this slab
      }
    }
  }
}

// carries [0, n) of one slab in pass order: every run of equal rows is added up left to right by the thread that owns its
// first entry and applied to y
__global__ void __launch_bounds__(kCbThreads) spmv_cb_carry_kernel(const CbCarry* __restrict__ carry, int64_t n, double* __restrict__ y,
                                                                     double alpha) {
  const int64_t stride = (int64_t)gridDim.x * kCbThreads;
  for (int64_t i = (int64_t)blockIdx.x * kCbThreads + threadIdx.x; i < n; i += stride) {
    const int32_t ri = carry[i].row;
    if (ri < 0 || (i > 0 && carry[i - 1].row == ri)) continue;  // not the head of a run
    double s = carry[i].sum;
    for (int64_t j = i + 1; j < n && carry[j].row == ri; ++j) s += carry[j].sum;
    y[ri] += alpha * s;
  }
}

// ---- builders ----
// key[i] = slab of nonzero i, order[i] = i
__global__ void __launch_bounds__(kCbThreads) cb_keys_kernel(const int32_t* __restrict__ colind, int64_t nnz, int32_t slab_cols,
                                                               uint32_t* __restrict__ key, uint32_t* __restrict__ order) {
  const int64_t stride = (int64_t)gridDim.x * kCbThreads;
  for (int64_t i = (int64_t)blockIdx.x * kCbThreads + threadIdx.x; i < nnz; i += stride) {
    key[i] = (uint32_t)(colind[i] / slab_cols);
    order[i] = (uint32_t)i;
  }
}
// row_of[i] = row of CSR nonzero i (one thread per row)
__global__ void __launch_bounds__(kCbThreads) cb_expand_rows_kernel(const int32_t* __restrict__ rowptr, int32_t n_rows, int32_t* __restrict__ row_of) {
  const int64_t stride = (int64_t)gridDim.x * kCbThreads;
  for (int64_t r = (int64_t)blockIdx.x * kCbThreads + threadIdx.x; r < n_rows; r += stride)
    for (int32_t i = rowptr[r]; i < rowptr[r + 1]; ++i) row_of[i] = (int32_t)r;
}
// slab-major arrays from the sorted order
__global__ void __launch_bounds__(kCbThreads) cb_gather_kernel(const uint32_t* __restrict__ order, const int32_t* __restrict__ row_of,
                                                                 const int32_t* __restrict__ colind, const double* __restrict__ vals, int64_t nnz,
                                                                 int32_t* __restrict__ row, int32_t* __restrict__ col, double* __restrict__ val) {
  const int64_t stride = (int64_t)gridDim.x * kCbThreads;
  for (int64_t i = (int64_t)blockIdx.x * kCbThreads + threadIdx.x; i < nnz; i += stride) {
    const uint32_t src = order[i];
    row[i] = row_of[src];
    col[i] = colind[src];
    val[i] = vals[src];
  }
}
// slab_ptr[s] = first position of slab s in the sorted key array; slab_ptr[n_slabs] = nnz
__global__ void __launch_bounds__(kCbThreads) cb_slab_ptr_kernel(const uint32_t* __restrict__ key, int64_t nnz, int32_t n_slabs,
                                                                   long long* __restrict__ slab_ptr) {
  const int64_t stride = (int64_t)gridDim.x * kCbThreads;
  for (int64_t i = (int64_t)blockIdx.x * kCbThreads + threadIdx.x; i <= nnz; i += stride) {
    const int64_t lo = i == 0 ? -1 : (int64_t)key[i - 1];
    const int64_t hi = i == nnz ? n_slabs : (int64_t)key[i];
    for (int64_t s = lo + 1; s <= hi; ++s) slab_ptr[s] = i;
  }
}

}  // namespace cbk
}  // namespace jp
