// jp_csr_kernels.cuh -- the localApply kernels (src/CSRMatrix.jl:1123-1162); see the header of jp_csr.cu for the
// data layout and the design.  Device code only, in the CUDA subset tests/cuda_emul can execute on the CPU
// (TMA / mbarrier / cp.async go through jp_ptx.cuh).
#pragma once
#include <cstdint>

#include "jp_blockdesc.h"
#include "jp_ptx.cuh"

namespace jp {
namespace csrk {

constexpr int kSpmvThreads = 256;
constexpr int kMaxRowsPerBlock = 512;   // rows staged per block (row pointers in shared memory)
constexpr int kLongRowLen = 64;         // blocks holding a row longer than this reduce warp-per-row

// KIND_SUBWARP carries log2(lanes per row) in bits 20..23 of BlockDesc::nrows_kind
enum BlockKind : int { KIND_DIRECT = 0, KIND_TWOPHASE = 1, KIND_WARPROW = 2, KIND_LONGROW = 3, KIND_SUBWARP = 4 };
constexpr int kSubwarpLongFactor = 16;  // SUBWARP blocks: rows longer than this many passes of their lane group go warp-per-row

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// y_new = beta==0 ? alpha*sum : (y*beta) + (sum*alpha)   -- src/CSRMatrix.jl:1142-1144 with the beta==0
// overwrite contract of src/Operator.jl:42
__device__ __forceinline__ double combine(double yold, double sum, double alpha, double beta) {
  double s = __dmul_rn(sum, alpha);
  return beta == 0.0 ? s : __dadd_rn(__dmul_rn(yold, beta), s);
}

// left-to-right sum of s_val[s..e) (the reference's order); four shared loads are issued ahead of the adds
__device__ __forceinline__ double row_sum_sequential(const double* s_val, int s, int e) {
  double sum = 0.0;
  int i = s;
  for (; i + 4 <= e; i += 4) {
    const double a0 = s_val[i], a1 = s_val[i + 1], a2 = s_val[i + 2], a3 = s_val[i + 3];
    sum = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(sum, a0), a1), a2), a3);
  }
  for (; i < e; ++i) sum = __dadd_rn(sum, s_val[i]);
  return sum;
}

// shared-memory stage: vals [cap+8] | colind [cap+8] | row pointers [kMaxRowsPerBlock+8]
struct StageView {
  double* s_val;
  int32_t* s_col;
  int32_t* s_rp;
};

__device__ __forceinline__ StageView stage_view(unsigned char* smem, int cap) {
  StageView v;
  v.s_val = reinterpret_cast<double*>(smem);
  v.s_col = reinterpret_cast<int32_t*>(v.s_val + cap + 8);
  v.s_rp = v.s_col + cap + 8;
  return v;
}

struct BlockRange {
  int row0, nrows, kind, p0, p1;
  int lg;   // KIND_SUBWARP: log2 of the lanes per row
  int a0;   // 4-aligned first staged nonzero: shared index of nonzero i is i - a0
  int rpo;  // shared index of row r's pointer is rpo + r
};

__device__ __forceinline__ BlockRange decode_block(const BlockDesc* __restrict__ blocks, int idx) {
  const int4 blk = __ldg(reinterpret_cast<const int4*>(blocks + idx));
  BlockRange b;
  b.row0 = blk.x;
  b.nrows = blk.y & 0xffff;
  b.kind = (blk.y >> 16) & 0xf;
  b.lg = (blk.y >> 20) & 0xf;
  b.p0 = blk.z;
  b.p1 = blk.w;
  b.a0 = b.p0 & ~3;
  b.rpo = b.row0 & 3;
  return b;
}

// Copy the block's nonzeros [a0, p1) and row pointers [row0, row0+nrows] into shared memory.
// TMA != 0: three bulk copies by one thread, mbarrier completion.  TMA == 0: 16-byte cp.async by all threads.
template <int TMA>
__device__ __forceinline__ void stage_block(const StageView& sv, uint64_t* bar, const int32_t* __restrict__ rowptr,
                                            const int32_t* __restrict__ colind, const double* __restrict__ vals, const BlockRange& b) {
  const int tid = threadIdx.x;
  const int ngrp = (b.p1 - b.a0 + 3) >> 2;                // groups of 4 nonzeros
  const int ra0 = b.row0 & ~3;                            // 16-byte aligned first row pointer
  const int nrp = (b.row0 + b.nrows + 1 - ra0 + 3) & ~3;  // row pointers staged (multiple of 4)
  if (TMA) {
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
      const uint64_t pol = policy_evict_first();
      mbar_expect_tx(bar, (unsigned)(ngrp * 48 + nrp * 4));
      if (ngrp > 0) {
        bulk_g2s(sv.s_val, vals + b.a0, (unsigned)ngrp * 32, bar, pol);
        bulk_g2s(sv.s_col, colind + b.a0, (unsigned)ngrp * 16, bar, pol);
      }
      bulk_g2s(sv.s_rp, rowptr + ra0, (unsigned)nrp * 4, bar, pol);
    }
    mbar_wait(bar, 0);
  } else {
    for (int g = tid; g < ngrp; g += kSpmvThreads) {
      cp_async16(sv.s_col + 4 * g, colind + b.a0 + 4 * g);
      cp_async16(sv.s_val + 4 * g, vals + b.a0 + 4 * g);
      cp_async16(sv.s_val + 4 * g + 2, vals + b.a0 + 4 * g + 2);
    }
    for (int g = tid; g < (nrp >> 2); g += kSpmvThreads) cp_async16(sv.s_rp + 4 * g, rowptr + ra0 + 4 * g);
    cp_async_wait_all();
    __syncthreads();
  }
}

// a single row longer than the stage: stream it straight from HBM with the whole block
template <bool HALO>
__device__ void long_row(const int32_t* __restrict__ colind, const double* __restrict__ vals, const double* __restrict__ x, int64_t ldx,
                         const double* __restrict__ halo, int32_t halo_k, int32_t nloc, double* __restrict__ y, int64_t ldy, int32_t k,
                         int row, int p0, int p1, double alpha, double beta, double* s_red) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int v = 0; v < k; ++v) {
    double acc = 0.0;
    for (int i = p0 + tid; i < p1; i += kSpmvThreads) {
      const int c = __ldg(colind + i);
      const double xv = (HALO && c >= nloc) ? halo[(size_t)(c - nloc) * halo_k + v] : __ldg(x + c + (size_t)v * ldx);
      acc += __ldg(vals + i) * xv;
    }
    acc = warp_sum(acc);
    __syncthreads();
    if (lane == 0) s_red[warp] = acc;
    __syncthreads();
    if (warp == 0) {
      double t = lane < kSpmvThreads / 32 ? s_red[lane] : 0.0;
      t = warp_sum(t);
      if (lane == 0) {
        double* yp = y + row + (size_t)v * ldy;
        *yp = combine(beta == 0.0 ? 0.0 : *yp, t, alpha, beta);
      }
    }
  }
}

// one row block of y = beta*y + alpha*A*x (k = 1); called by every thread of the CTA.  Returns this thread's share of
// sum_r w[r] * y_new[r] over the block's rows when DOT (0 otherwise).
template <bool HALO, int TMA, bool DOT = false>
__device__ __forceinline__ double spmv_block(const BlockRange& b, const StageView& sv, uint64_t* s_bar_p, double* s_red,
                                             const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                                             const double* __restrict__ vals, const double* __restrict__ x,
                                             const double* __restrict__ halo, int32_t nloc, double* __restrict__ y, double alpha, double beta,
                                             const double* __restrict__ w = nullptr) {
  const int tid = threadIdx.x;
  double dacc = 0.0;
  if (b.kind == KIND_LONGROW) {
    long_row<HALO>(colind, vals, x, 0, halo, 1, nloc, y, 0, 1, b.row0, b.p0, b.p1, alpha, beta, s_red);
    if (DOT && tid == 0) dacc = __ldg(w + b.row0) * y[b.row0];  // thread 0 wrote y[row0] itself
    return dacc;
  }
  // DOT: this thread's first weight is requested before the stage is waited for, so its latency is not serialised behind
  // the row sum (profiles/r2_ncu_spmv_dot_summary.txt: the fused kernel was 33 % slower than the plain SpMV without this)
  double w_first = 0.0;
  if (DOT && tid < b.nrows) w_first = __ldg(w + b.row0 + tid);
  stage_block<TMA>(sv, s_bar_p, rowptr, colind, vals, b);
  const int32_t* rp = sv.s_rp + b.rpo;  // rp[r] is a global nonzero index; shared index = rp[r] - a0
  if (b.kind == KIND_DIRECT) {
    // thread-per-row straight from the stage; lanes = consecutive rows
    for (int r = tid; r < b.nrows; r += kSpmvThreads) {
      const int s = rp[r] - b.a0, e = rp[r + 1] - b.a0;
      double sum = 0.0;
#pragma unroll 4
      for (int i = s; i < e; ++i) {
        const int c = sv.s_col[i];
        const double xv = (HALO && c >= nloc) ? halo[c - nloc] : __ldg(x + c);
        sum = __dadd_rn(sum, __dmul_rn(sv.s_val[i], xv));
      }
      double* yp = y + b.row0 + r;
      const double ynew = combine(beta == 0.0 ? 0.0 : *yp, sum, alpha, beta);
      *yp = ynew;
      if (DOT) dacc += (r == tid ? w_first : __ldg(w + b.row0 + r)) * ynew;
    }
    return dacc;
  }
  if (b.kind == KIND_SUBWARP) {
    // T = 2^lg lanes per row straight out of the stage: no product round trip through shared memory, no second block
    // barrier.  Lanes of a group read consecutive staged entries (conflict-free), gather x, and the group adds up with
    // log2(T) shuffles; every group works on TWO rows at a time so that two gathers per thread are in flight.  Rows much
    // longer than the group (heavy tail) are left to whole warps in a second pass.
    const int T = 1 << b.lg, lane_t = tid & (T - 1), grp = tid >> b.lg, ngrp = kSpmvThreads >> b.lg;
    const int longlen = kSubwarpLongFactor * T;
    const uint64_t pol_x = policy_evict_last();
    bool any_long = false;
    for (int rbase = 0; rbase < b.nrows; rbase += 2 * ngrp) {  // warp-uniform trip count (shuffles below)
      const int ra = rbase + grp, rb = ra + ngrp;
      int ia = 0, ea = 0, ib = 0, eb = 0;
      bool wa = ra < b.nrows, wb = rb < b.nrows;  // this group writes y[ra] / y[rb] in this pass
      if (wa) {
        ia = rp[ra] - b.a0;
        ea = rp[ra + 1] - b.a0;
        if (ea - ia > longlen) { ea = ia; wa = false; any_long = true; }
      }
      if (wb) {
        ib = rp[rb] - b.a0;
        eb = rp[rb + 1] - b.a0;
        if (eb - ib > longlen) { eb = ib; wb = false; any_long = true; }
      }
      double acca = 0.0, accb = 0.0;
      ia += lane_t;
      ib += lane_t;
      while (ia < ea || ib < eb) {
        if (ia < ea) {
          const int c = sv.s_col[ia];
          const double xv = (HALO && c >= nloc) ? halo[c - nloc] : ldg_gather(x + c, pol_x);
          acca = fma(sv.s_val[ia], xv, acca);
          ia += T;
        }
        if (ib < eb) {
          const int c = sv.s_col[ib];
          const double xv = (HALO && c >= nloc) ? halo[c - nloc] : ldg_gather(x + c, pol_x);
          accb = fma(sv.s_val[ib], xv, accb);
          ib += T;
        }
      }
      for (int o = T >> 1; o > 0; o >>= 1) {
        acca += __shfl_xor_sync(0xffffffffu, acca, o);
        accb += __shfl_xor_sync(0xffffffffu, accb, o);
      }
      if (lane_t == 0) {
        if (wa) {
          double* yp = y + b.row0 + ra;
          const double ynew = combine(beta == 0.0 ? 0.0 : *yp, acca, alpha, beta);
          *yp = ynew;
          if (DOT) dacc += __ldg(w + b.row0 + ra) * ynew;
        }
        if (wb) {
          double* yp = y + b.row0 + rb;
          const double ynew = combine(beta == 0.0 ? 0.0 : *yp, accb, alpha, beta);
          *yp = ynew;
          if (DOT) dacc += __ldg(w + b.row0 + rb) * ynew;
        }
      }
    }
    if (__syncthreads_or(any_long)) {
      const int lane = tid & 31, warp = tid >> 5;
      for (int r = warp; r < b.nrows; r += kSpmvThreads / 32) {
        const int s = rp[r] - b.a0, e = rp[r + 1] - b.a0;
        if (e - s <= longlen) continue;  // warp-uniform
        double a0 = 0.0, a1 = 0.0;
        int i = s + lane;
        for (; i + 32 < e; i += 64) {
          const int c0 = sv.s_col[i], c1 = sv.s_col[i + 32];
          const double x0 = (HALO && c0 >= nloc) ? halo[c0 - nloc] : ldg_gather(x + c0, pol_x);
          const double x1 = (HALO && c1 >= nloc) ? halo[c1 - nloc] : ldg_gather(x + c1, pol_x);
          a0 = fma(sv.s_val[i], x0, a0);
          a1 = fma(sv.s_val[i + 32], x1, a1);
        }
        if (i < e) {
          const int c0 = sv.s_col[i];
          a0 = fma(sv.s_val[i], (HALO && c0 >= nloc) ? halo[c0 - nloc] : ldg_gather(x + c0, pol_x), a0);
        }
        const double sum = warp_sum(a0 + a1);
        if (lane == 0) {
          double* yp = y + b.row0 + r;
          const double ynew = combine(beta == 0.0 ? 0.0 : *yp, sum, alpha, beta);
          *yp = ynew;
          if (DOT) dacc += __ldg(w + b.row0 + r) * ynew;
        }
      }
    }
    return dacc;
  }
  // phase A: products, thread-per-nonzero (every gather of the block is independent of row structure);
  // 8 nonzeros per thread per pass so 8 gathers are in flight before the first product is needed
  const int s0 = b.p0 - b.a0, s1 = b.p1 - b.a0;
  for (int base = s0 + tid; base < s1; base += 8 * kSpmvThreads) {
    int c[8];
    double xv[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int i = base + q * kSpmvThreads;
      c[q] = i < s1 ? sv.s_col[i] : -1;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q)
      xv[q] = c[q] < 0 ? 0.0 : ((HALO && c[q] >= nloc) ? halo[c[q] - nloc] : __ldg(x + c[q]));
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int i = base + q * kSpmvThreads;
      if (i < s1) sv.s_val[i] = __dmul_rn(sv.s_val[i], xv[q]);
    }
  }
  __syncthreads();
  // phase B: row sums
  if (b.kind == KIND_TWOPHASE) {
    for (int r = tid; r < b.nrows; r += kSpmvThreads) {
      const int s = rp[r] - b.a0, e = rp[r + 1] - b.a0;
      double* yp = y + b.row0 + r;
      const double ynew = combine(beta == 0.0 ? 0.0 : *yp, row_sum_sequential(sv.s_val, s, e), alpha, beta);
      *yp = ynew;
      if (DOT) dacc += (r == tid ? w_first : __ldg(w + b.row0 + r)) * ynew;
    }
  } else {
    // WARPROW: the block holds at least one long row.  Short rows still go thread-per-row (sequential sum, the
    // reference's order); only rows longer than kLongRowLen are summed by a whole warp.
    for (int r = tid; r < b.nrows; r += kSpmvThreads) {
      const int s = rp[r] - b.a0, e = rp[r + 1] - b.a0;
      if (e - s <= kLongRowLen) {
        double* yp = y + b.row0 + r;
        *yp = combine(beta == 0.0 ? 0.0 : *yp, row_sum_sequential(sv.s_val, s, e), alpha, beta);
        if (DOT) dacc += __ldg(w + b.row0 + r) * *yp;
      }
    }
    const int lane = tid & 31, warp = tid >> 5;
    for (int r = warp; r < b.nrows; r += kSpmvThreads / 32) {
      const int s = rp[r] - b.a0, e = rp[r + 1] - b.a0;
      if (e - s <= kLongRowLen) continue;  // warp-uniform
      double sum = 0.0;
      for (int i = s + lane; i < e; i += 32) sum += sv.s_val[i];
      sum = warp_sum(sum);
      if (lane == 0) {
        double* yp = y + b.row0 + r;
        *yp = combine(beta == 0.0 ? 0.0 : *yp, sum, alpha, beta);
        if (DOT) dacc += __ldg(w + b.row0 + r) * *yp;
      }
    }
  }
  return dacc;
}

// One partial per WARP: partials[bid * (kSpmvThreads / 32) + warp] = sum of `v` over the warp's lanes.  No block-level
// step: a warp stores its sum and is done, so the CTA gives its slot back as soon as its last warp has finished.
// (History: finishing the grid-wide sum inside the kernel with __threadfence + a ticket made the fused kernel 11 us slower
// than the plain SpMV on 7-pt 128^3; one partial per CTA still cost two block barriers at the end of every CTA, 15 % of
// the CTA's lifetime of a kernel that runs at the HBM roofline -- profiles/r2_next_rows_256.json.  sum_partials_kernel
// adds the partials up in a fixed order behind it.)
constexpr int kDotPartialsPerBlock = kSpmvThreads / 32;
__device__ __forceinline__ void warp_partial_store(double v, double* __restrict__ partials, unsigned bid) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double s = warp_sum(v);
  if (lane == 0) partials[(size_t)bid * kDotPartialsPerBlock + warp] = s;
}

// out[0] = sum of partials[0..n) in a fixed order.  Two levels in one launch: CTA c adds partials [c*kSumChunk,
// (c+1)*kSumChunk) -- 16 independent loads per thread -- into level2[c]; the last CTA to finish (ticket, re-armed) adds
// the level-2 values in CTA order.  (A single CTA over the 65536 partials of a 256^3 apply took ~100 us: a third of the
// apply itself, profiles/r2_next_rows_256.json.)
constexpr unsigned kSumChunk = kSpmvThreads * 16;
__host__ __device__ inline unsigned sum_partials_grid(unsigned n) { return n == 0 ? 1u : (n + kSumChunk - 1) / kSumChunk; }

__global__ void __launch_bounds__(kSpmvThreads) sum_partials_kernel(const double* __restrict__ partials, unsigned n, double* __restrict__ out,
                                                                      double* __restrict__ level2, unsigned int* __restrict__ ticket) {
  __shared__ double s_red[kSpmvThreads / 32];
  __shared__ bool s_last;
  const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned base = blockIdx.x * kSumChunk;
  double v[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const unsigned i = base + q * kSpmvThreads + tid;
    v[q] = i < n ? partials[i] : 0.0;
  }
  double a = 0.0;
#pragma unroll
  for (int q = 0; q < 16; ++q) a += v[q];
  const double t = warp_sum(a);
  if (lane == 0) s_red[warp] = t;
  __syncthreads();
  if (warp == 0) {
    double u = lane < kSpmvThreads / 32 ? s_red[lane] : 0.0;
    u = warp_sum(u);
    if (lane == 0) {
      level2[blockIdx.x] = u;
      __threadfence();
      s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
  }
  __syncthreads();
  if (s_last && warp == 0) {
    __threadfence();
    const volatile double* l2 = level2;
    double u = 0.0;
    for (unsigned c = lane; c < gridDim.x; c += 32) u += l2[c];
    u = warp_sum(u);
    if (lane == 0) {
      out[0] = u;
      *ticket = 0;
    }
  }
}

// apply + dot in one pass: y = beta*y + alpha*A*x and partials[block] = sum_r w[r] * y[r] over the
// block's rows, so a CG iteration's p'Ap costs no second sweep over p and Ap (16 bytes per row less HBM traffic)
template <bool HALO, int TMA>
__global__ void __launch_bounds__(kSpmvThreads)
spmv_dot_kernel(const BlockDesc* __restrict__ blocks, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                const double* __restrict__ vals, const double* __restrict__ x, const double* __restrict__ halo, int32_t nloc,
                double* __restrict__ y, int32_t cap, double alpha, double beta, const double* __restrict__ w,
                double* __restrict__ partials) {
  JP_DYN_SMEM(smem_raw);
  __shared__ double s_red[kSpmvThreads / 32];
  __shared__ __align__(8) uint64_t s_bar;
  const StageView sv = stage_view(smem_raw, cap);
  const BlockRange b = decode_block(blocks, blockIdx.x);
  const double dacc = spmv_block<HALO, TMA, true>(b, sv, &s_bar, s_red, rowptr, colind, vals, x, halo, nloc, y, alpha, beta, w);
  warp_partial_store(dacc, partials, blockIdx.x);
}

template <bool HALO, int TMA>
__global__ void __launch_bounds__(kSpmvThreads)
spmv_kernel(const BlockDesc* __restrict__ blocks, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
            const double* __restrict__ vals, const double* __restrict__ x, const double* __restrict__ halo, int32_t nloc,
            double* __restrict__ y, int32_t cap, double alpha, double beta) {
  JP_DYN_SMEM(smem_raw);
  __shared__ double s_red[kSpmvThreads / 32];
  __shared__ __align__(8) uint64_t s_bar;
  const StageView sv = stage_view(smem_raw, cap);
  const BlockRange b = decode_block(blocks, blockIdx.x);
  spmv_block<HALO, TMA>(b, sv, &s_bar, s_red, rowptr, colind, vals, x, halo, nloc, y, alpha, beta);
}

template <int KT, bool HALO, int TMA>
__global__ void __launch_bounds__(kSpmvThreads)
spmm_kernel(const BlockDesc* __restrict__ blocks, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
            const double* __restrict__ vals, const double* __restrict__ x, int64_t ldx, const double* __restrict__ halo,
            int32_t halo_k, int32_t nloc, double* __restrict__ y, int64_t ldy, int32_t k, int32_t cap, int32_t T, double alpha,
            double beta) {
  JP_DYN_SMEM(smem_raw);
  __shared__ double s_red[kSpmvThreads / 32];
  __shared__ __align__(8) uint64_t s_bar;
  const StageView sv = stage_view(smem_raw, cap);
  const BlockRange b = decode_block(blocks, blockIdx.x);
  const int tid = threadIdx.x;
  if (b.kind == KIND_LONGROW) {
    long_row<HALO>(colind, vals, x, ldx, halo, halo_k, nloc, y, ldy, k, b.row0, b.p0, b.p1, alpha, beta, s_red);
    return;
  }
  stage_block<TMA>(sv, &s_bar, rowptr, colind, vals, b);
  const int32_t* rp = sv.s_rp + b.rpo;
  // blocks holding long rows use a full warp per row; otherwise T lanes per row from the matrix statistics
  const int Tb = b.kind == KIND_WARPROW ? 32 : (b.kind == KIND_SUBWARP && (1 << b.lg) > T) ? (1 << b.lg) : T;
  const int lane_t = tid & (Tb - 1), group = tid / Tb, ngroups = kSpmvThreads / Tb;
  for (int v0 = 0; v0 < k; v0 += KT) {
    const int kt = k - v0 < KT ? k - v0 : KT;  // warp-uniform
    const double* xt = x + (size_t)v0 * ldx;
    for (int rbase = 0; rbase < b.nrows; rbase += ngroups) {  // warp-uniform trip count (shuffles below)
      const int r = rbase + group;
      const bool valid = r < b.nrows;
      const int s = valid ? rp[r] - b.a0 : 0, e = valid ? rp[r + 1] - b.a0 : 0;
      double acc[KT];
#pragma unroll
      for (int j = 0; j < KT; ++j) acc[j] = 0.0;
      if (kt == KT) {
        for (int i = s + lane_t; i < e; i += Tb) {
          const int c = sv.s_col[i];
          const double v = sv.s_val[i];
          if (HALO && c >= nloc) {
            const double* h = halo + (size_t)(c - nloc) * halo_k + v0;
#pragma unroll
            for (int j = 0; j < KT; ++j) acc[j] = fma(v, h[j], acc[j]);
          } else {
            const double* xc = xt + c;
#pragma unroll
            for (int j = 0; j < KT; ++j) acc[j] = fma(v, __ldg(xc + (size_t)j * ldx), acc[j]);
          }
        }
      } else {
        for (int i = s + lane_t; i < e; i += Tb) {
          const int c = sv.s_col[i];
          const double v = sv.s_val[i];
          if (HALO && c >= nloc) {
            const double* h = halo + (size_t)(c - nloc) * halo_k + v0;
#pragma unroll
            for (int j = 0; j < KT; ++j)
              if (j < kt) acc[j] = fma(v, h[j], acc[j]);
          } else {
            const double* xc = xt + c;
#pragma unroll
            for (int j = 0; j < KT; ++j)
              if (j < kt) acc[j] = fma(v, __ldg(xc + (size_t)j * ldx), acc[j]);
          }
        }
      }
      for (int o = Tb >> 1; o > 0; o >>= 1) {
#pragma unroll
        for (int j = 0; j < KT; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
      }
      if (valid && lane_t == 0) {
#pragma unroll
        for (int j = 0; j < KT; ++j)
          if (j < kt) {
            double* yp = y + b.row0 + r + (size_t)(v0 + j) * ldy;
            *yp = combine(beta == 0.0 ? 0.0 : *yp, acc[j], alpha, beta);
          }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// SpMM with the multivector tile staged in shared memory (banded / stencil-like matrices).
// A k-vector apply gathers nnz*k doubles of X; with X only in L1/L2 that traffic (12 GB for 27-pt 192^3, k = 8) is
// several times the matrix stream and the L1 that is left beside the staged matrix (~20 KB) cannot hold a block's
// working set.  At build time every row block gets the sorted list of the UNIQUE columns it references (ucols) and
// every nonzero the uint16 position of its column in that list (lidx).  The kernel stages vals, lidx, row pointers
// and ucols with TMA bulk copies, gathers the block's X tile [KT vectors] x [unique columns] ONCE into shared
// memory (coalesced along the column runs of a banded matrix; a column referenced by 3 neighbouring rows is
// fetched once, not 3 times), then runs thread-per-row out of shared memory.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kTileThreads = 256;

struct TileStage {
  double* s_val;    // [cap + 8]
  double* xs;       // [KT][upad]
  int32_t* s_rp;    // [kMaxRowsPerBlock + 8]
  int32_t* s_ucol;  // [upad]
  uint16_t* s_lidx; // [cap + 16]
};
__host__ __device__ inline size_t tile_smem_bytes(int cap, int upad, int KT) {
  return (size_t)(cap + 8) * 8 + (size_t)KT * upad * 8 + (size_t)(kMaxRowsPerBlock + 8) * 4 + (size_t)upad * 4 + (size_t)(cap + 16) * 2;
}

template <int KT, bool HALO>
__global__ void __launch_bounds__(kTileThreads)
spmm_tile_kernel(const BlockDesc* __restrict__ blocks, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                 const double* __restrict__ vals, const uint16_t* __restrict__ lidx, const int32_t* __restrict__ ucols,
                 const double* __restrict__ x, int64_t ldx, const double* __restrict__ halo, int32_t halo_k, int32_t nloc,
                 double* __restrict__ y, int64_t ldy, int32_t k, int32_t cap, int32_t upad, double alpha, double beta) {
  JP_DYN_SMEM(smem_raw);
  __shared__ double s_red[kTileThreads / 32];
  __shared__ __align__(8) uint64_t s_bar;
  TileStage t;
  t.s_val = reinterpret_cast<double*>(smem_raw);
  t.xs = t.s_val + cap + 8;
  t.s_rp = reinterpret_cast<int32_t*>(t.xs + (size_t)KT * upad);
  t.s_ucol = t.s_rp + kMaxRowsPerBlock + 8;
  t.s_lidx = reinterpret_cast<uint16_t*>(t.s_ucol + upad);
  const BlockRange b = decode_block(blocks, blockIdx.x);
  const int4 aux = __ldg(reinterpret_cast<const int4*>(blocks + blockIdx.x) + 1);
  const int ucol0 = aux.x, U = aux.y;
  const int tid = threadIdx.x;
  if (b.kind == KIND_LONGROW) {
    long_row<HALO>(colind, vals, x, ldx, halo, halo_k, nloc, y, ldy, k, b.row0, b.p0, b.p1, alpha, beta, s_red);
    return;
  }
  // stage: vals, lidx, row pointers, unique columns (four TMA bulk copies)
  const int ngrp = (b.p1 - b.a0 + 3) >> 2;
  const int al0 = b.p0 & ~7;                       // 16-byte aligned first staged lidx
  const int nl8 = (b.p1 - al0 + 7) & ~7;           // lidx staged (multiple of 8)
  const int ra0 = b.row0 & ~3;
  const int nrp = (b.row0 + b.nrows + 1 - ra0 + 3) & ~3;
  const int nu4 = (U + 3) & ~3;
  if (tid == 0) mbar_init(&s_bar, 1);
  __syncthreads();
  if (tid == 0) {
    const uint64_t pol = policy_evict_first();
    mbar_expect_tx(&s_bar, (unsigned)(ngrp * 32 + nl8 * 2 + nrp * 4 + nu4 * 4));
    if (ngrp > 0) bulk_g2s(t.s_val, vals + b.a0, (unsigned)ngrp * 32, &s_bar, pol);
    if (nl8 > 0) bulk_g2s(t.s_lidx, lidx + al0, (unsigned)nl8 * 2, &s_bar, pol);
    bulk_g2s(t.s_rp, rowptr + ra0, (unsigned)nrp * 4, &s_bar, pol);
    if (nu4 > 0) bulk_g2s(t.s_ucol, ucols + ucol0, (unsigned)nu4 * 4, &s_bar, pol);
  }
  mbar_wait(&s_bar, 0);
  const int32_t* rp = t.s_rp + b.rpo;
  // lanes per row: as many as keep every thread busy, at most 4 (warp-uniform, power of two)
  int T = 1;
  while (T < 4 && b.nrows * (T << 1) <= kTileThreads) T <<= 1;
  const int lane_t = tid & (T - 1), group = tid / T, ngroups = kTileThreads / T;
  for (int v0 = 0; v0 < k; v0 += KT) {
    // X tile: xs[j][u] = X[ucol[u], v0 + j]
    for (int u = tid; u < U; u += kTileThreads) {
      const int c = t.s_ucol[u];
      if (HALO && c >= nloc) {
        const double* h = halo + (size_t)(c - nloc) * halo_k + v0;
#pragma unroll
        for (int j = 0; j < KT; ++j)
          if (v0 + j < k) t.xs[(size_t)j * upad + u] = h[j];
      } else {
        const double* xc = x + c + (size_t)v0 * ldx;
#pragma unroll
        for (int j = 0; j < KT; ++j)
          if (v0 + j < k) t.xs[(size_t)j * upad + u] = __ldg(xc + (size_t)j * ldx);
      }
    }
    __syncthreads();
    for (int rbase = 0; rbase < b.nrows; rbase += ngroups) {  // warp-uniform trip count (shuffles below)
      const int r = rbase + group;
      const bool valid = r < b.nrows;
      const int s = valid ? rp[r] : 0, e = valid ? rp[r + 1] : 0;
      double acc[KT];
#pragma unroll
      for (int j = 0; j < KT; ++j) acc[j] = 0.0;
      for (int i = s + lane_t; i < e; i += T) {
        const int li = t.s_lidx[i - al0];
        const double v = t.s_val[i - b.a0];
#pragma unroll
        for (int j = 0; j < KT; ++j) acc[j] = fma(v, t.xs[(size_t)j * upad + li], acc[j]);
      }
      for (int o = T >> 1; o > 0; o >>= 1) {
#pragma unroll
        for (int j = 0; j < KT; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
      }
      if (valid && lane_t == 0) {
#pragma unroll
        for (int j = 0; j < KT; ++j)
          if (v0 + j < k) {
            double* yp = y + b.row0 + r + (size_t)(v0 + j) * ldy;
            *yp = combine(beta == 0.0 ? 0.0 : *yp, acc[j], alpha, beta);
          }
      }
    }
    __syncthreads();  // the next vector tile overwrites xs
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Pipelined SpMM: persistent CTAs (one per SM), two shared-memory stages, warp specialisation.
//   producer warps (8): wait(empty[s]) -> TMA bulk copies of the block's vals / lidx / row pointers / unique columns
//                       -> gather the block's X tile [KT vectors] x [unique columns] into shared memory -> arrive(full[s])
//   consumer warps (8): wait(full[s]) -> T lanes per row out of shared memory, KT accumulators in registers -> y
//                       -> arrive(empty[s])
// so the TMA latency and the X-tile gather of item i+1 overlap the arithmetic of item i.  An item is one
// (row block, vector tile) pair; the tiles of a block run back to back on the same CTA.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kPipeConsumers = 256, kPipeProducers = 256, kPipeThreads = kPipeConsumers + kPipeProducers;


__host__ __device__ inline size_t pipe_stage_bytes(int cap, int upad, int KT) { return (tile_smem_bytes(cap, upad, KT) + 127) & ~(size_t)127; }

template <int KT, bool HALO>
__global__ void __launch_bounds__(kPipeThreads, 1)
spmm_pipe_kernel(const BlockDesc* __restrict__ blocks, int32_t n_blocks, const int32_t* __restrict__ rowptr,
                 const double* __restrict__ vals, const uint16_t* __restrict__ lidx, const int32_t* __restrict__ ucols,
                 const double* __restrict__ x, int64_t ldx, const double* __restrict__ halo, int32_t halo_k, int32_t nloc,
                 double* __restrict__ y, int64_t ldy, int32_t k, int32_t cap, int32_t upad, double alpha, double beta) {
  JP_DYN_SMEM(smem_raw);
  __shared__ __align__(8) uint64_t tma_bar[2], full_bar[2], empty_bar[2];
  const int tid = threadIdx.x, lane = tid & 31;
  const bool producer = tid >= kPipeConsumers;
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tma_bar[s], 1);
      mbar_init(&full_bar[s], kPipeProducers / 32);
      mbar_init(&empty_bar[s], kPipeConsumers / 32);
    }
  }
  __syncthreads();
  const size_t stage_bytes = pipe_stage_bytes(cap, upad, KT);
  const int ntiles = (k + KT - 1) / KT;
  const int my_blocks = blockIdx.x < n_blocks ? (n_blocks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int total_items = my_blocks * ntiles;  // item = (row block, vector tile); the tiles of a block run back to back

  // stage carve-up
  auto s_val_of = [&](int s) { return reinterpret_cast<double*>(smem_raw + (size_t)s * stage_bytes); };

  if (producer) {
    const int ptid = tid - kPipeConsumers;
    // the TMA of item it+1 is issued before the X tile of item it is gathered, so its latency is hidden
    auto issue_tma = [&](int it) {
      const int bi = blockIdx.x + (it / ntiles) * gridDim.x;
      const BlockRange b = decode_block(blocks, bi);
      const int4 aux = __ldg(reinterpret_cast<const int4*>(blocks + bi) + 1);
      const int s = it & 1;
      const unsigned ph = (it >> 1) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1);  // the consumers are done with this stage (passes at once the first time)
      double* s_val = s_val_of(s);
      double* xs = s_val + cap + 8;
      int32_t* s_rp = reinterpret_cast<int32_t*>(xs + (size_t)KT * upad);
      int32_t* s_ucol = s_rp + kMaxRowsPerBlock + 8;
      uint16_t* s_lidx = reinterpret_cast<uint16_t*>(s_ucol + upad);
      const int al0 = b.p0 & ~7;
      const int ngrp = (b.p1 - b.a0 + 3) >> 2;
      const int nl8 = (b.p1 - al0 + 7) & ~7;
      const int ra0 = b.row0 & ~3;
      const int nrp = (b.row0 + b.nrows + 1 - ra0 + 3) & ~3;
      const int nu4 = (aux.y + 3) & ~3;
      const uint64_t pol = policy_evict_first();
      mbar_expect_tx(&tma_bar[s], (unsigned)(ngrp * 32 + nl8 * 2 + nrp * 4 + nu4 * 4));
      if (ngrp > 0) bulk_g2s(s_val, vals + b.a0, (unsigned)ngrp * 32, &tma_bar[s], pol);
      if (nl8 > 0) bulk_g2s(s_lidx, lidx + al0, (unsigned)nl8 * 2, &tma_bar[s], pol);
      bulk_g2s(s_rp, rowptr + ra0, (unsigned)nrp * 4, &tma_bar[s], pol);
      if (nu4 > 0) bulk_g2s(s_ucol, ucols + aux.x, (unsigned)nu4 * 4, &tma_bar[s], pol);
    };
    if (ptid == 0 && total_items > 0) issue_tma(0);
    for (int it = 0; it < total_items; ++it) {
      if (ptid == 0 && it + 1 < total_items) issue_tma(it + 1);
      const int bi = blockIdx.x + (it / ntiles) * gridDim.x;
      const int4 aux = __ldg(reinterpret_cast<const int4*>(blocks + bi) + 1);
      const int U = aux.y;
      const int s = it & 1;
      const unsigned ph = (it >> 1) & 1;
      const int v0 = (it % ntiles) * KT;
      const int kt = k - v0 < KT ? k - v0 : KT;
      double* xs = s_val_of(s) + cap + 8;
      const int32_t* s_ucol = reinterpret_cast<const int32_t*>(xs + (size_t)KT * upad) + kMaxRowsPerBlock + 8;
      mbar_wait(&empty_bar[s], ph ^ 1);  // before overwriting xs of this stage
      mbar_wait(&tma_bar[s], ph);        // the unique-column list has landed
      // X tile: xs[j][u] = X[ucol[u], v0 + j]; four columns per thread per pass -> 4*KT loads in flight
      for (int ub = ptid; ub < U; ub += 4 * kPipeProducers) {
        int c[4];
        double a[4][KT];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int u = ub + q * kPipeProducers;
          c[q] = u < U ? s_ucol[u] : -1;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (c[q] < 0) continue;
          if (HALO && c[q] >= nloc) {
            const double* h = halo + (size_t)(c[q] - nloc) * halo_k + v0;
#pragma unroll
            for (int j = 0; j < KT; ++j) a[q][j] = j < kt ? h[j] : 0.0;
          } else {
            const double* xc = x + c[q] + (size_t)v0 * ldx;
#pragma unroll
            for (int j = 0; j < KT; ++j) a[q][j] = j < kt ? __ldg(xc + (size_t)j * ldx) : 0.0;
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (c[q] < 0) continue;
          const int u = ub + q * kPipeProducers;
#pragma unroll
          for (int j = 0; j < KT; ++j) xs[(size_t)j * upad + u] = a[q][j];
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);
    }
  } else {
    for (int it = 0; it < total_items; ++it) {
      const int bi = blockIdx.x + (it / ntiles) * gridDim.x;
      const BlockRange b = decode_block(blocks, bi);
      const int s = it & 1;
      const unsigned ph = (it >> 1) & 1;
      const int v0 = (it % ntiles) * KT;
      const int kt = k - v0 < KT ? k - v0 : KT;
      const int al0 = b.p0 & ~7;
      const double* s_val = s_val_of(s);
      const double* xs = s_val + cap + 8;
      const int32_t* s_rp = reinterpret_cast<const int32_t*>(xs + (size_t)KT * upad);
      const uint16_t* s_lidx = reinterpret_cast<const uint16_t*>(s_rp + kMaxRowsPerBlock + 8 + upad);
      mbar_wait(&full_bar[s], ph);
      mbar_wait(&tma_bar[s], ph);  // completed long ago; orders this thread after the bulk copies themselves
      const int32_t* rp = s_rp + b.rpo;
      int T = 1;
      while (T < 4 && b.nrows * (T << 1) <= kPipeConsumers) T <<= 1;
      const int lane_t = tid & (T - 1), group = tid / T, ngroups = kPipeConsumers / T;
      for (int rbase = 0; rbase < b.nrows; rbase += ngroups) {  // warp-uniform trip count (shuffles below)
        const int r = rbase + group;
        const bool valid = r < b.nrows;
        const int rs = valid ? rp[r] : 0, re = valid ? rp[r + 1] : 0;
        double acc[KT];
#pragma unroll
        for (int j = 0; j < KT; ++j) acc[j] = 0.0;
#pragma unroll 2
        for (int i = rs + lane_t; i < re; i += T) {
          const int li = s_lidx[i - al0];
          const double v = s_val[i - b.a0];
#pragma unroll
          for (int j = 0; j < KT; ++j) acc[j] = fma(v, xs[(size_t)j * upad + li], acc[j]);
        }
        for (int o = T >> 1; o > 0; o >>= 1) {
#pragma unroll
          for (int j = 0; j < KT; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
        }
        if (valid && lane_t == 0) {
#pragma unroll
          for (int j = 0; j < KT; ++j)
            if (j < kt) {
              double* yp = y + b.row0 + r + (size_t)(v0 + j) * ldy;
              *yp = combine(beta == 0.0 ? 0.0 : *yp, acc[j], alpha, beta);
            }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// SpMM with the block's X tile staged through ALIGNED COLUMN RUNS (banded / stencil-like matrices), many small CTAs.
// Lessons of the two kernels above (profiles/r1_ncu_spmm_powerlaw_summary.txt): gathering X through L1 costs ~3x the
// wavefronts of a shared-memory read, and a tile that is gathered through registers AFTER the matrix stage has landed
// serialises two global-memory latencies per block at 3 CTAs/SM.  Here
//   * a block is at most kRunsMaxRows rows; the columns it references are merged at build time into 16-byte aligned
//     runs (a 27-point block of 64 rows = 9 runs of 66 columns), and every nonzero carries the uint16 position of its
//     column in the concatenated runs;
//   * ONE thread issues the TMA bulk copies of vals / positions / row pointers, while ALL threads issue 16-byte
//     cp.async copies of the runs straight from X into the tile (no registers, no dependency on the matrix stage: the
//     run list is read from global memory), so both latencies overlap;
//   * then T lanes per row accumulate KT vectors out of shared memory: consecutive lanes are consecutive rows, their
//     tile positions are consecutive -> conflict-free LDS.64;
//   * 128 or 256 threads and ~40 KB (KT = 4) of shared memory per CTA -> 5 CTAs/SM keep the LDS pipe busy while other
//     CTAs wait for their copies.
// X must be 16-byte aligned with an even leading dimension (the launcher checks; otherwise the gather kernel runs).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kRunsMaxRows = 128;  // upper bound of the rows-per-block setting (JP_SPMM_RUNS_ROWS, default 128)

// Tile columns a run of `len` columns occupies: padded so that consecutive runs start 10 (mod 16) doubles apart.  With
// two lanes per row, the iteration in which a lane pair straddles two runs reads [base_q + r + 2] and [base_q+1 + r]
// for 8 consecutive rows r per half-warp; 10 apart puts the two 8-bank windows side by side instead of on top of each
// other (profiles/r1_ncu_spmm_runs_summary.txt: 38 % of the shared-memory wavefronts were conflict replays without it).
__host__ __device__ inline int run_stride(int len) { return len + ((10 - len) & 15); }

__host__ __device__ inline size_t runs_smem_bytes(int cap, int wpad, int KT) {
  return (size_t)(cap + 8) * 8 + (size_t)KT * wpad * 8 + (size_t)(kRunsMaxRows + 8) * 4 + (size_t)(cap + 16) * 2;
}

// acc[j] += sum over the row's entries i = first, first + T, ... < e of val[i] * xs[j][lidx[i]].  The kernel is bound by the
// latency of its shared-memory loads (profiles/r2_ncu_spmm_runs_summary.txt: 12 warps per SM, short-scoreboard stalls), so
// four entries are in flight at once: their positions and values are read first, then the 4 * KT tile reads, then the FMAs.
template <int KT>
__device__ __forceinline__ void runs_row_accumulate(double (&acc)[KT], const uint16_t* __restrict__ lidx0, const double* __restrict__ val0,
                                                    const double* __restrict__ xs, int wpad, int first, int e, int T) {
  int i = first;
  for (; i + 3 * T < e; i += 4 * T) {
    const int l0 = lidx0[i], l1 = lidx0[i + T], l2 = lidx0[i + 2 * T], l3 = lidx0[i + 3 * T];
    const double v0 = val0[i], v1 = val0[i + T], v2 = val0[i + 2 * T], v3 = val0[i + 3 * T];
    double x0[KT], x1[KT], x2[KT], x3[KT];
#pragma unroll
    for (int j = 0; j < KT; ++j) {
      x0[j] = xs[j * wpad + l0];
      x1[j] = xs[j * wpad + l1];
      x2[j] = xs[j * wpad + l2];
      x3[j] = xs[j * wpad + l3];
    }
#pragma unroll
    for (int j = 0; j < KT; ++j) acc[j] = fma(v3, x3[j], fma(v2, x2[j], fma(v1, x1[j], fma(v0, x0[j], acc[j]))));
  }
  for (; i < e; i += T) {
    const int li = lidx0[i];
    const double v = val0[i];
#pragma unroll
    for (int j = 0; j < KT; ++j) acc[j] = fma(v, xs[j * wpad + li], acc[j]);
  }
}

template <int KT, int THREADS>
__global__ void __launch_bounds__(THREADS)
spmm_runs_kernel(const BlockDesc* __restrict__ blocks, const int32_t* __restrict__ rowptr, const double* __restrict__ vals,
                 const uint16_t* __restrict__ lidx, const int2* __restrict__ runs, const double* __restrict__ x, int64_t ldx,
                 double* __restrict__ y, int64_t ldy, int32_t k, int32_t cap, int32_t wpad, double alpha, double beta) {
  JP_DYN_SMEM(smem_raw);
  __shared__ __align__(8) uint64_t s_bar;
  double* s_val = reinterpret_cast<double*>(smem_raw);                  // [cap + 8]
  double* xs = s_val + cap + 8;                                          // [KT][wpad]
  int32_t* s_rp = reinterpret_cast<int32_t*>(xs + KT * wpad);            // [kRunsMaxRows + 8]
  uint16_t* s_lidx = reinterpret_cast<uint16_t*>(s_rp + kRunsMaxRows + 8);  // [cap + 16]
  const BlockRange b = decode_block(blocks, blockIdx.x);
  const int4 aux = __ldg(reinterpret_cast<const int4*>(blocks + blockIdx.x) + 1);
  const int run0 = aux.x, nruns = aux.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // matrix stage: vals, tile positions, row pointers (three TMA bulk copies by one thread)
  const int ngrp = (b.p1 - b.a0 + 3) >> 2;
  const int al0 = b.p0 & ~7;               // 16-byte aligned first staged position
  const int nl8 = (b.p1 - al0 + 7) & ~7;   // positions staged (multiple of 8)
  const int ra0 = b.row0 & ~3;
  const int nrp = (b.row0 + b.nrows + 1 - ra0 + 3) & ~3;
  if (tid == 0) mbar_init(&s_bar, 1);
  __syncthreads();
  const int kt0 = k < KT ? k : KT;
  if (tid == 0) {  // the first phase of the barrier covers this stage AND the first X tile
    const uint64_t pol = policy_evict_first();
    mbar_expect_tx(&s_bar, (unsigned)(ngrp * 32 + nl8 * 2 + nrp * 4 + aux.w * 8 * kt0));
    if (ngrp > 0) bulk_g2s(s_val, vals + b.a0, (unsigned)ngrp * 32, &s_bar, pol);
    if (nl8 > 0) bulk_g2s(s_lidx, lidx + al0, (unsigned)nl8 * 2, &s_bar, pol);
    bulk_g2s(s_rp, rowptr + ra0, (unsigned)nrp * 4, &s_bar, pol);
  }
  // lanes per row: as many as keep every thread busy, at most 4 (block-uniform, power of two)
  int T = 1;
  while (T < 4 && b.nrows * (T << 1) <= THREADS) T <<= 1;
  const int lane_t = tid & (T - 1), group = tid / T, ngroups = THREADS / T;
  // the shared-memory views are indexed relative to the staged ranges
  const double* val0 = s_val - b.a0;
  const uint16_t* lidx0 = s_lidx - al0;
  unsigned phase = 0;
  for (int v0 = 0; v0 < k; v0 += KT) {
    const int kt = k - v0 < KT ? k - v0 : KT;
    // X tile by TMA as well (no LSU issue slots, no registers): warp 0, one lane per run, one bulk copy per (run, vector):
    // xs[j][base + c - start] = X[c, v0 + j]; aux.w = columns copied per vector.
    if (warp == 0) {
      if (v0 > 0 && lane == 0) mbar_expect_tx(&s_bar, (unsigned)(aux.w * 8 * kt));
      __syncwarp();  // the barrier is armed before any lane's copy can complete on it
      const uint64_t pol_x = policy_evict_last();
      int carry = 0;
      for (int q0 = 0; q0 < nruns; q0 += 32) {  // warp-uniform
        const int q = q0 + lane;
        int2 run = make_int2(0, 0);
        if (q < nruns) run = __ldg(runs + run0 + q);
        const int stride = q < nruns ? run_stride(run.y) : 0;
        int incl = stride;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        const int base = carry + incl - stride;
        carry += __shfl_sync(0xffffffffu, incl, 31);
        if (run.y > 0) {
          const double* src = x + (size_t)v0 * ldx + run.x;
          double* dst = xs + base;
          for (int j = 0; j < kt; ++j) {
            bulk_g2s(dst, src, (unsigned)run.y * 8, &s_bar, pol_x);
            src += ldx;
            dst += wpad;
          }
        }
      }
    }
    mbar_wait(&s_bar, phase);
    phase ^= 1;
    const int32_t* rp = s_rp + b.rpo;
    for (int rbase = 0; rbase < b.nrows; rbase += ngroups) {  // warp-uniform trip count (shuffles below)
      const int r = rbase + group;
      const bool valid = r < b.nrows;
      const int s = valid ? rp[r] : 0, e = valid ? rp[r + 1] : 0;
      double acc[KT];
#pragma unroll
      for (int j = 0; j < KT; ++j) acc[j] = 0.0;
      runs_row_accumulate<KT>(acc, lidx0, val0, xs, wpad, s + lane_t, e, T);
      for (int o = T >> 1; o > 0; o >>= 1) {
#pragma unroll
        for (int j = 0; j < KT; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
      }
      if (valid && lane_t == 0) {
#pragma unroll
        for (int j = 0; j < KT; ++j)
          if (j < kt) {
            double* yp = y + b.row0 + r + (size_t)(v0 + j) * ldy;
            *yp = combine(beta == 0.0 ? 0.0 : *yp, acc[j], alpha, beta);
          }
      }
    }
    __syncthreads();  // the next vector tile overwrites xs
  }
}

// ---------------------------------------------------------------------------------------------------------------
// spmm_runs_kernel with the X tile DOUBLE-BUFFERED across the vector tiles of a block (round-2 candidate, opt-in with
// JP_SPMM_RUNS_DB=1; written from the phase analysis in DESIGN.md 3.2b, not yet measured).  In spmm_runs_kernel every
// (block, vector tile) phase pays the whole chain  run list from global -> TMA issue -> arrival -> barrier  before its
// ~0.5 us of arithmetic.  Here
//   * the run list and the tile bases are read ONCE per block and stay in registers (one run per lane of warp 0);
//   * while the CTA computes vector tile t out of tile buffer t & 1, warp 0 has already issued the bulk copies of tile
//     t + 1 into the other buffer (its own mbarrier), so from the second tile on a phase costs its arithmetic only;
//   * the extra tile buffer takes the CTA from 73 KB to 110 KB: two CTAs per SM instead of three.
// Blocks with more than 32 runs keep re-reading their run list (rare: a banded block has 3..18 runs).
// ---------------------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t runs_db_smem_bytes(int cap, int wpad, int KT) {
  return (size_t)(cap + 8) * 8 + 2 * (size_t)KT * wpad * 8 + (size_t)(kRunsMaxRows + 8) * 4 + (size_t)(cap + 16) * 2;
}

template <int KT, int THREADS>
__global__ void __launch_bounds__(THREADS)
spmm_runs_db_kernel(const BlockDesc* __restrict__ blocks, const int32_t* __restrict__ rowptr, const double* __restrict__ vals,
                    const uint16_t* __restrict__ lidx, const int2* __restrict__ runs, const double* __restrict__ x, int64_t ldx,
                    double* __restrict__ y, int64_t ldy, int32_t k, int32_t cap, int32_t wpad, double alpha, double beta) {
  JP_DYN_SMEM(smem_raw);
  __shared__ __align__(8) uint64_t s_bar[2];  // one per tile buffer; the first phase of s_bar[0] also covers the matrix stage
  double* s_val = reinterpret_cast<double*>(smem_raw);                         // [cap + 8]
  double* xs_base = s_val + cap + 8;                                            // [2][KT][wpad]
  int32_t* s_rp = reinterpret_cast<int32_t*>(xs_base + 2 * KT * wpad);          // [kRunsMaxRows + 8]
  uint16_t* s_lidx = reinterpret_cast<uint16_t*>(s_rp + kRunsMaxRows + 8);      // [cap + 16]
  const BlockRange b = decode_block(blocks, blockIdx.x);
  const int4 aux = __ldg(reinterpret_cast<const int4*>(blocks + blockIdx.x) + 1);
  const int run0 = aux.x, nruns = aux.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ngrp = (b.p1 - b.a0 + 3) >> 2;
  const int al0 = b.p0 & ~7;
  const int nl8 = (b.p1 - al0 + 7) & ~7;
  const int ra0 = b.row0 & ~3;
  const int nrp = (b.row0 + b.nrows + 1 - ra0 + 3) & ~3;
  const int ntiles = (k + KT - 1) / KT;
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
  }
  __syncthreads();
  // warp 0, lane q: run q of the block and its position in the tile (kept for every vector tile)
  int2 my_run = make_int2(0, 0);
  int my_base = 0;
  if (warp == 0 && nruns <= 32) {
    if (lane < nruns) my_run = __ldg(runs + run0 + lane);
    const int stride = lane < nruns ? run_stride(my_run.y) : 0;
    int incl = stride;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    my_base = incl - stride;
  }
  // issue the bulk copies of vector tile t into tile buffer t & 1 (warp 0 only; the barrier has been armed by lane 0)
  auto issue_tile = [&](int t) {
    const int v0 = t * KT;
    const int kt = k - v0 < KT ? k - v0 : KT;
    double* xs = xs_base + (size_t)(t & 1) * KT * wpad;
    uint64_t* bar = &s_bar[t & 1];
    const uint64_t pol_x = policy_evict_last();
    if (nruns <= 32) {
      if (my_run.y > 0) {
        const double* src = x + (size_t)v0 * ldx + my_run.x;
        double* dst = xs + my_base;
        for (int j = 0; j < kt; ++j) {
          bulk_g2s(dst, src, (unsigned)my_run.y * 8, bar, pol_x);
          src += ldx;
          dst += wpad;
        }
      }
    } else {
      int carry = 0;
      for (int q0 = 0; q0 < nruns; q0 += 32) {  // warp-uniform
        const int q = q0 + lane;
        int2 run = make_int2(0, 0);
        if (q < nruns) run = __ldg(runs + run0 + q);
        const int stride = q < nruns ? run_stride(run.y) : 0;
        int incl = stride;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int tt = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += tt;
        }
        const int base = carry + incl - stride;
        carry += __shfl_sync(0xffffffffu, incl, 31);
        if (run.y > 0) {
          const double* src = x + (size_t)v0 * ldx + run.x;
          double* dst = xs + base;
          for (int j = 0; j < kt; ++j) {
            bulk_g2s(dst, src, (unsigned)run.y * 8, bar, pol_x);
            src += ldx;
            dst += wpad;
          }
        }
      }
    }
  };
  const int kt0 = k < KT ? k : KT;
  if (tid == 0) {  // matrix stage + first tile on s_bar[0]
    const uint64_t pol = policy_evict_first();
    mbar_expect_tx(&s_bar[0], (unsigned)(ngrp * 32 + nl8 * 2 + nrp * 4 + aux.w * 8 * kt0));
    if (ngrp > 0) bulk_g2s(s_val, vals + b.a0, (unsigned)ngrp * 32, &s_bar[0], pol);
    if (nl8 > 0) bulk_g2s(s_lidx, lidx + al0, (unsigned)nl8 * 2, &s_bar[0], pol);
    bulk_g2s(s_rp, rowptr + ra0, (unsigned)nrp * 4, &s_bar[0], pol);
  }
  if (warp == 0) {
    __syncwarp();  // s_bar[0] is armed before any lane's copy can complete on it
    issue_tile(0);
  }
  int T = 1;
  while (T < 4 && b.nrows * (T << 1) <= THREADS) T <<= 1;
  const int lane_t = tid & (T - 1), group = tid / T, ngroups = THREADS / T;
  const double* val0 = s_val - b.a0;
  const uint16_t* lidx0 = s_lidx - al0;
  for (int t = 0; t < ntiles; ++t) {
    const int v0 = t * KT;
    const int kt = k - v0 < KT ? k - v0 : KT;
    if (warp == 0 && t + 1 < ntiles) {
      // prefetch the next vector tile into the other buffer: every thread left it at the __syncthreads that ended
      // phase t - 1
      const int ktn = k - (v0 + KT) < KT ? k - (v0 + KT) : KT;
      if (lane == 0) mbar_expect_tx(&s_bar[(t + 1) & 1], (unsigned)(aux.w * 8 * ktn));
      __syncwarp();
      issue_tile(t + 1);
    }
    mbar_wait(&s_bar[t & 1], (unsigned)((t >> 1) & 1));
    const double* xs = xs_base + (size_t)(t & 1) * KT * wpad;
    const int32_t* rp = s_rp + b.rpo;
    for (int rbase = 0; rbase < b.nrows; rbase += ngroups) {  // warp-uniform trip count (shuffles below)
      const int r = rbase + group;
      const bool valid = r < b.nrows;
      const int s = valid ? rp[r] : 0, e = valid ? rp[r + 1] : 0;
      double acc[KT];
#pragma unroll
      for (int j = 0; j < KT; ++j) acc[j] = 0.0;
      runs_row_accumulate<KT>(acc, lidx0, val0, xs, wpad, s + lane_t, e, T);
      for (int o = T >> 1; o > 0; o >>= 1) {
#pragma unroll
        for (int j = 0; j < KT; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
      }
      if (valid && lane_t == 0) {
#pragma unroll
        for (int j = 0; j < KT; ++j)
          if (j < kt) {
            double* yp = y + b.row0 + r + (size_t)(v0 + j) * ldy;
            *yp = combine(beta == 0.0 ? 0.0 : *yp, acc[j], alpha, beta);
          }
      }
    }
    __syncthreads();  // tile buffer t & 1 is free for vector tile t + 2
  }
}

// jp_csr_create, one pass over the uploaded column indices (rows strided over the threads): make them 0-based, find every
// row's largest column (-1 for an empty row) and count the indices outside [0, n_cols) -- the O(nnz) part of the
// validation runs here, not on one host core
__global__ void csr_validate_kernel(const int32_t* __restrict__ rowptr, int32_t* __restrict__ colind, int32_t n_rows, int32_t n_cols, int32_t base,
                                    int32_t* __restrict__ rowmax, unsigned long long* __restrict__ bad) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned long long nbad = 0;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += stride) {
    int32_t mx = -1;
    for (int32_t i = rowptr[r]; i < rowptr[r + 1]; ++i) {
      const int32_t c = colind[i] - base;
      if (base != 0) colind[i] = c;
      if (c < 0 || c >= n_cols) ++nbad;
      mx = c > mx ? c : mx;
    }
    rowmax[r] = mx;
  }
  if (nbad) atomicAdd(bad, nbad);
}

}  // namespace csrk
}  // namespace jp
