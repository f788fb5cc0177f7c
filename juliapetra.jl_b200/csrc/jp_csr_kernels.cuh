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

// KIND_SUBWARP carries log2(lanes per row) in bits 20..23 of BlockDesc::nrows_kind.  (Kinds 1 and 2 were round 1's
// two-phase paths -- products through shared memory, then thread- or warp-per-row sums -- replaced by KIND_SUBWARP:
// profiles/r2_powerlaw10M_subwarp_ab.json.)
enum BlockKind : int { KIND_DIRECT = 0, KIND_LONGROW = 3, KIND_SUBWARP = 4 };
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
    for (int rbase = 0; rbase < b.nrows; rbase += 2 * ngrp) {  // warp-uniform trip count (shuffles below)
      const int ra = rbase + grp, rb = ra + ngrp;
      int ia = 0, ea = 0, ib = 0, eb = 0;
      bool wa = ra < b.nrows, wb = rb < b.nrows;  // this group writes y[ra] / y[rb] in this pass
      if (wa) {
        ia = rp[ra] - b.a0;
        ea = rp[ra + 1] - b.a0;
        if (ea - ia > longlen) { ea = ia; wa = false; }
      }
      if (wb) {
        ib = rp[rb] - b.a0;
        eb = rp[rb + 1] - b.a0;
        if (eb - ib > longlen) { eb = ib; wb = false; }
      }
      double acca = 0.0, accb = 0.0;
      ia += lane_t;
      ib += lane_t;
      // (issuing two entries per row and step -- four gathers in flight per lane -- measured no faster: the kernel is at
      // the memory system's random-sector rate, profiles/r2_ncu_spmv_powerlaw10M_summary.txt)
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
    // second pass, no block barrier in between (rows are independent; a barrier here made every warp wait for the block's
    // longest short row: 18 stall cycles per issued instruction, profiles/r2_ncu_spmv_powerlaw10M_summary.txt): every warp
    // looks for long rows among its share of the block's rows as soon as it is done with the first pass
    {
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
  return dacc;  // (every block kind returns above)
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

static __global__ void __launch_bounds__(kSpmvThreads) sum_partials_kernel(const double* __restrict__ partials, unsigned n, double* __restrict__ out,
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
  const int Tb = (b.kind == KIND_SUBWARP && (1 << b.lg) > T) ? (1 << b.lg) : T;
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
// SpMM on a matrix whose nonzeros lie on a few DIAGONALS (stencils: 5 / 7 / 27 distinct values of col - row).
// What bounds the CSR SpMM kernels above is not HBM but the shared-memory pipe and instruction issue: every FMA needs
// its x operand from shared memory (8 bytes; 128 bytes per clock per SM = 16 FMA per clock per SM) plus, in CSR, a
// position load, a dependent address computation and the value load (profiles/r2_ncu_spmm_runs_summary.txt: issue slots
// 25 % busy, LSU wavefronts 52 %, 12 warps per SM waiting on dependent shared-memory loads).  For a diagonal-structured
// matrix none of the indirection is needed.  At the first k > 1 apply the values are copied once into diagonal-major
// order, vals_dia[d][row] (zero where a row has no entry on diagonal d: domain boundaries; no column indices at all, so
// the matrix stream shrinks from 12 to ~8 bytes per nonzero), and
//   * one thread = one row; lanes = consecutive rows.  The thread's D coefficients are read ONCE per row block with
//     coalesced streaming loads and stay in registers for every vector tile;
//   * the X tile of the block is staged by TMA bulk copies: the columns a block of R rows touches are R + (a few) wide
//     windows, one per group of adjacent diagonals, merged where they overlap; the window list is the same for every
//     block (relative to its first row), so it travels as a kernel parameter -- no run table in memory;
//   * the inner loop is  acc[j] = fma(a[d], xs[j][pos_d + lane], acc[j])  with every address known up front: D * KT
//     independent conflict-free LDS.64, no dependent loads; diagonals are visited in ascending column order, which is
//     the order of the CSR row, so the result is bit-identical to the run-staged kernel's.
// Rows closer to the first / last row than the largest offset (their windows would leave X) are served by the gather
// kernel on the corresponding CSR row blocks.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kDiaThreads = 128;   // rows per block (one per thread)
constexpr int kDiaMaxDiags = 32;
constexpr int kDiaMaxSegs = 16;
constexpr int kDiaTileStride = 1536;  // doubles between the vectors of the X tile: a compile-time constant, so that the KT
                                      // reads of one diagonal are one address + immediate offsets (W <= kDiaTileStride)

struct DiaParams {
  int32_t D;                     // diagonals, ascending offset
  int32_t nseg;                  // tile segments
  int32_t W;                     // tile width in doubles (even)
  int32_t seg_cols;              // sum of the segment lengths (doubles copied per vector)
  int32_t pos[kDiaMaxDiags];     // tile position of x[base + off_d] for the block's first row `base`
  int32_t seg_rel[kDiaMaxSegs];  // first column of segment s relative to `base` (even)
  int32_t seg_len[kDiaMaxSegs];  // its length in columns (even)
  int32_t seg_pos[kDiaMaxSegs];  // where it starts in the tile (even)
};

// issue the bulk copies of one vector tile (called by warp 0): lane = 4 * segment + vector, eight segments per pass
template <int KT>
__device__ __forceinline__ void dia_issue_tile(const DiaParams& P, double* xs, uint64_t* bar, const double* __restrict__ x, int64_t ldx,
                                               int base, int v0, int kt) {
  static_assert(KT == 4, "lane mapping below assumes four vectors per tile");
  const int lane = threadIdx.x & 31;
  if (lane == 0) mbar_expect_tx(bar, (unsigned)(P.seg_cols * 8 * kt));
  __syncwarp();  // the barrier is armed before any lane's copy can complete on it
  const uint64_t pol_x = policy_evict_last();
  const int j = lane & 3;
  for (int sgm = lane >> 2; sgm < P.nseg; sgm += 8)
    if (j < kt)
      bulk_g2s(xs + j * kDiaTileStride + P.seg_pos[sgm], x + (size_t)(v0 + j) * ldx + base + P.seg_rel[sgm], (unsigned)P.seg_len[sgm] * 8, bar,
               pol_x);
}

template <int KT, int DMAX>
__global__ void __launch_bounds__(kDiaThreads, 4)  // <= 128 registers: four CTAs (16 warps) per SM beside the 4 x 38 KB tiles
spmm_dia_kernel(const double* __restrict__ vals_dia, int64_t npad, const double* __restrict__ x, int64_t ldx, double* __restrict__ y,
                int64_t ldy, int32_t k, int32_t row_lo, int32_t row_hi, double alpha, double beta, const DiaParams P) {
  JP_DYN_SMEM(smem_raw);
  __shared__ __align__(8) uint64_t s_bar;
  double* xs = reinterpret_cast<double*>(smem_raw);  // [KT][kDiaTileStride], W columns of each row in use
  const int tid = threadIdx.x, warp = tid >> 5;
  const int base = row_lo + (int)blockIdx.x * kDiaThreads;  // its parity is row_lo's: the plan's windows are aligned for it
  const int r = base + tid;
  const bool valid = r < row_hi;
  if (tid == 0) mbar_init(&s_bar, 1);
  __syncthreads();
  if (warp == 0) dia_issue_tile<KT>(P, xs, &s_bar, x, ldx, base, 0, k < KT ? k : KT);
  // this row's coefficients: coalesced (lanes = rows), streamed once, kept in registers for all vector tiles.  Diagonals
  // d >= D (the template rounds D up) get a zero coefficient and tile position 0: a harmless fma(0, finite, acc).
  double a[DMAX];
  {
    const double* vp = vals_dia + r;
#pragma unroll
    for (int d = 0; d < DMAX; ++d) {
      a[d] = (d < P.D && valid) ? __ldcs(vp) : 0.0;
      vp += npad;
    }
  }
  const double* xt = xs + tid;  // this lane's column of the tile
  unsigned phase = 0;
  for (int v0 = 0; v0 < k; v0 += KT) {
    const int kt = k - v0 < KT ? k - v0 : KT;
    mbar_wait(&s_bar, phase);
    phase ^= 1;
    double acc[KT];
#pragma unroll
    for (int j = 0; j < KT; ++j) acc[j] = 0.0;
#pragma unroll
    for (int d = 0; d < DMAX; ++d) {
      const double* xd = xt + P.pos[d];
#pragma unroll
      for (int j = 0; j < KT; ++j) acc[j] = fma(a[d], xd[j * kDiaTileStride], acc[j]);
    }
    __syncthreads();  // every thread has read the tile: the next one may land
    if (warp == 0 && v0 + KT < k) dia_issue_tile<KT>(P, xs, &s_bar, x, ldx, base, v0 + KT, k - (v0 + KT) < KT ? k - (v0 + KT) : KT);
    if (valid) {
#pragma unroll
      for (int j = 0; j < KT; ++j)
        if (j < kt) {
          double* yp = y + r + (size_t)(v0 + j) * ldy;
          *yp = combine(beta == 0.0 ? 0.0 : *yp, acc[j], alpha, beta);
        }
    }
  }
}

// diagonal-major copy of the values: vals_dia[d * npad + row] = value of the entry (row, row + off[d]); one thread per row
static __global__ void __launch_bounds__(256) dia_fill_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                                                         const double* __restrict__ vals, int32_t n_rows, const DiaParams P,
                                                         const int32_t* __restrict__ offsets, int64_t npad, double* __restrict__ vals_dia,
                                                         unsigned long long* __restrict__ missing) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += stride) {
    for (int32_t i = rowptr[r]; i < rowptr[r + 1]; ++i) {
      const int32_t off = colind[i] - (int32_t)r;
      int lo = 0, hi = P.D - 1, d = -1;
      while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const int32_t o = offsets[mid];
        if (o == off) { d = mid; break; }
        if (o < off) lo = mid + 1;
        else hi = mid - 1;
      }
      if (d < 0) atomicAdd(missing, 1ull);
      else vals_dia[(size_t)d * npad + r] = vals[i];
    }
  }
}

// one bit per distinct value of col - row (+ n_rows - 1)
static __global__ void __launch_bounds__(256) dia_mark_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind, int32_t n_rows,
                                                         uint32_t* __restrict__ bitmap) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += stride) {
    for (int32_t i = rowptr[r]; i < rowptr[r + 1]; ++i) {
      const uint32_t b = (uint32_t)(colind[i] - (int32_t)r + (n_rows - 1));
      const uint32_t w = b >> 5, m = 1u << (b & 31);
      if ((bitmap[w] & m) == 0) atomicOr(bitmap + w, m);  // the plain read filters all but the first setters of a bit
    }
  }
}

// jp_csr_create, one pass over the uploaded column indices (rows strided over the threads): make them 0-based, find every
// row's largest column (-1 for an empty row) and count the indices outside [0, n_cols) -- the O(nnz) part of the
// validation runs here, not on one host core
static __global__ void csr_validate_kernel(const int32_t* __restrict__ rowptr, int32_t* __restrict__ colind, int32_t n_rows, int32_t n_cols, int32_t base,
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
