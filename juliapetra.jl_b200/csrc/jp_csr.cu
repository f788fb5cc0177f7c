// jp_csr.cu -- device-resident LocalCSRMatrix (src/LocalCSRMatrix.jl:3-8, src/LocalCSRGraph.jl:11-14) and
// the localApply kernels (src/CSRMatrix.jl:1123-1162).
//
// Data layout in HBM: rowptr int32[n+1] (0-based), colind int32[nnz+pad], vals f64[nnz+pad]; colind/vals are
// streamed exactly once per apply.  At create time the rows are cut into ROW BLOCKS: contiguous row ranges whose
// nnz fit the shared-memory stage (nnz_cap) and whose row count fits the thread block; a row longer than the
// stage gets a block of its own.  Rows that touch only local columns ("interior") and rows that touch remote
// columns ("boundary") go to separate launch lists so the interior list can run while the halo is in flight.
// Each block carries a KIND chosen from its row-length statistics (the "warp- or subwarp-per-row chosen by
// row-length statistics" of the north star):
//   DIRECT   near-uniform short rows (stencils): thread-per-row straight out of the shared-memory stage
//   TWOPHASE irregular rows: thread-per-nonzero products, then thread-per-row sums
//   WARPROW  blocks holding a long row: thread-per-nonzero products, then warp-per-row sums
//   LONGROW  one row longer than the stage: the whole block streams it from HBM
//
// One CTA per row block.  STAGE: one elected thread issues three TMA bulk copies (cp.async.bulk, SASS UBLKCP)
// global->shared for the block's vals, colind and row pointers, completing on an mbarrier; the copies are
// 16-byte aligned supersets of the block's ranges, fully coalesced irrespective of row boundaries, carry an
// L2 evict-first policy (the matrix is streamed once; x should own the L2) and cost no LSU issue slots.
// DIRECT then has lanes = consecutive rows, so in a banded matrix the gathers x[col] of one warp instruction are
// consecutive doubles (coalesced through the read-only path), shared-memory reads are conflict-free for odd row
// lengths, and the per-row sum runs left to right with unfused multiply/add: bit-identical to the reference.
// SpMM (k > 1): same stage, then T lanes per row accumulate KT vectors at a time in registers; the matrix is
// read from HBM once for all k (the reference re-streams it per vector).
// No tensor cores: this is not a dense contraction.  No atomics in NO_TRANS: results are deterministic.
#include <algorithm>
#include <climits>
#include <cstdlib>
#include <thread>

#include "jp_common.h"
#include "jp_p2p.cuh"

namespace jp {

namespace {

constexpr int kSpmvThreads = 256;
constexpr int kMaxRowsPerBlock = 512;   // rows staged per block (row pointers in shared memory)
constexpr int kLongRowLen = 64;         // blocks holding a row longer than this reduce warp-per-row

enum BlockKind : int { KIND_DIRECT = 0, KIND_TWOPHASE = 1, KIND_WARPROW = 2, KIND_LONGROW = 3 };

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
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n" ::"r"(
                   smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(smem_addr(bar)), "l"(pol)
               : "memory");
}

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
  int a0;   // 4-aligned first staged nonzero: shared index of nonzero i is i - a0
  int rpo;  // shared index of row r's pointer is rpo + r
};

__device__ __forceinline__ BlockRange decode_block(const BlockDesc* __restrict__ blocks, int idx) {
  const int4 blk = __ldg(reinterpret_cast<const int4*>(blocks + idx));
  BlockRange b;
  b.row0 = blk.x;
  b.nrows = blk.y & 0xffff;
  b.kind = blk.y >> 16;
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

// one row block of y = beta*y + alpha*A*x (k = 1); called by every thread of the CTA
template <bool HALO, int TMA>
__device__ __forceinline__ void spmv_block(const BlockRange& b, const StageView& sv, uint64_t* s_bar_p, double* s_red,
                                           const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                                           const double* __restrict__ vals, const double* __restrict__ x,
                                           const double* __restrict__ halo, int32_t nloc, double* __restrict__ y, double alpha, double beta) {
  const int tid = threadIdx.x;
  if (b.kind == KIND_LONGROW) {
    long_row<HALO>(colind, vals, x, 0, halo, 1, nloc, y, 0, 1, b.row0, b.p0, b.p1, alpha, beta, s_red);
    return;
  }
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
      *yp = combine(beta == 0.0 ? 0.0 : *yp, sum, alpha, beta);
    }
    return;
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
      *yp = combine(beta == 0.0 ? 0.0 : *yp, row_sum_sequential(sv.s_val, s, e), alpha, beta);
    }
  } else {
    // WARPROW: the block holds at least one long row.  Short rows still go thread-per-row (sequential sum, the
    // reference's order); only rows longer than kLongRowLen are summed by a whole warp.
    for (int r = tid; r < b.nrows; r += kSpmvThreads) {
      const int s = rp[r] - b.a0, e = rp[r + 1] - b.a0;
      if (e - s <= kLongRowLen) {
        double* yp = y + b.row0 + r;
        *yp = combine(beta == 0.0 ? 0.0 : *yp, row_sum_sequential(sv.s_val, s, e), alpha, beta);
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
      }
    }
  }
}

template <bool HALO, int TMA>
__global__ void __launch_bounds__(kSpmvThreads)
spmv_kernel(const BlockDesc* __restrict__ blocks, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
            const double* __restrict__ vals, const double* __restrict__ x, const double* __restrict__ halo, int32_t nloc,
            double* __restrict__ y, int32_t cap, double alpha, double beta) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ double s_red[kSpmvThreads / 32];
  __shared__ __align__(8) uint64_t s_bar;
  const StageView sv = stage_view(smem_raw, cap);
  const BlockRange b = decode_block(blocks, blockIdx.x);
  spmv_block<HALO, TMA>(b, sv, &s_bar, s_red, rowptr, colind, vals, x, halo, nloc, y, alpha, beta);
}

// ---------------------------------------------------------------------------------------------------------------
// apply! with the halo exchange INSIDE the SpMV kernel (k = 1, peer-memory halo): one launch per apply, no stream
// events, no NCCL call.  The grid is [pack CTAs | interior row blocks | boundary row blocks]:
//   pack CTAs      gather the exported rows of x and store them straight into the neighbours' halo buffers over
//                  NVLink (IPC-mapped peer pointers), then publish ready = epoch with a system-scope release;
//   interior CTAs  the ordinary row-block SpMV on local columns (they are the bulk of the grid and hide the exchange);
//   boundary CTAs  come last in the grid: they acquire the neighbours' ready flags (set ~30 us earlier by the peers'
//                  pack CTAs, which never wait on this GPU's current kernel), run the row-block SpMV on x + halo, and
//                  the last one publishes ack = epoch so the senders may overwrite the halo buffer next time.
// CTAs are dispatched in index order, so pack CTAs run first and a waiting boundary CTA never holds a slot that a
// pack or interior CTA still needs.
// ---------------------------------------------------------------------------------------------------------------
struct FusedSpmvArgs {
  const BlockDesc* interior;
  const BlockDesc* boundary;
  int32_t n_pack, n_interior, n_boundary;
  const int32_t* rowptr;
  const int32_t* colind;
  const double* vals;
  const double* x;
  const double* halo;
  int32_t nloc;
  double* y;
  int32_t cap;
  double alpha, beta;
  FusedHalo h;
};

__global__ void __launch_bounds__(kSpmvThreads) spmv_fused_halo_kernel(const FusedSpmvArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ double s_red[kSpmvThreads / 32];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ bool s_last;
  const int tid = threadIdx.x;
  const int bid = blockIdx.x;
  const uint64_t epoch = a.h.epoch;
  if (bid < a.n_pack) {
    // ---- pack + send ----
    if (tid < a.h.n_sends && epoch > 1)  // back-pressure: the receiver consumed the previous halo
      wait_flag_ge(a.h.sends[tid].ack_flag, epoch - 1, a.h.timeout_flag);
    __syncthreads();
    const int64_t stride = (int64_t)a.n_pack * kSpmvThreads;
    for (int64_t i = (int64_t)bid * kSpmvThreads + tid; i < a.h.n_export; i += stride) {
      int sb = 0;
      while (sb + 1 < a.h.n_sends && i >= a.h.sends[sb].start + a.h.sends[sb].count) ++sb;
      a.h.sends[sb].dst[i - a.h.sends[sb].start] = __ldg(a.x + __ldg(a.h.lids + i));
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(a.h.tickets + 0, 1u) == (unsigned)a.n_pack - 1);
    __syncthreads();
    if (s_last) {
      __threadfence_system();
      if (tid < a.h.n_sends) st_release_sys(a.h.sends[tid].ready_flag, epoch);
      if (tid == 0) a.h.tickets[0] = 0;  // re-arm for the next apply (stream-ordered after this kernel)
    }
  } else if (bid < a.n_pack + a.n_interior) {
    const StageView sv = stage_view(smem_raw, a.cap);
    const BlockRange b = decode_block(a.interior, bid - a.n_pack);
    spmv_block<false, 1>(b, sv, &s_bar, s_red, a.rowptr, a.colind, a.vals, a.x, nullptr, a.nloc, a.y, a.alpha, a.beta);
  } else {
    // ---- boundary rows: wait for the neighbours' halos first ----
    if (tid < a.h.n_sources) wait_flag_ge(a.h.sources[tid].ready_flag, epoch, a.h.timeout_flag);
    __syncthreads();
    const StageView sv = stage_view(smem_raw, a.cap);
    const BlockRange b = decode_block(a.boundary, bid - a.n_pack - a.n_interior);
    spmv_block<true, 1>(b, sv, &s_bar, s_red, a.rowptr, a.colind, a.vals, a.x, a.halo, a.nloc, a.y, a.alpha, a.beta);
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(a.h.tickets + 1, 1u) == (unsigned)a.n_boundary - 1);
    __syncthreads();
    if (s_last) {
      if (tid < a.h.n_sources) st_release_sys(a.h.sources[tid].ack_flag, epoch);
      if (tid == 0) a.h.tickets[1] = 0;
    }
  }
}

template <int KT, bool HALO, int TMA>
__global__ void __launch_bounds__(kSpmvThreads)
spmm_kernel(const BlockDesc* __restrict__ blocks, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
            const double* __restrict__ vals, const double* __restrict__ x, int64_t ldx, const double* __restrict__ halo,
            int32_t halo_k, int32_t nloc, double* __restrict__ y, int64_t ldy, int32_t k, int32_t cap, int32_t T, double alpha,
            double beta) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
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
  const int Tb = b.kind == KIND_WARPROW ? 32 : T;
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
  extern __shared__ __align__(128) unsigned char smem_raw[];
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

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_addr(bar)) : "memory");
}

__host__ __device__ inline size_t pipe_stage_bytes(int cap, int upad, int KT) { return (tile_smem_bytes(cap, upad, KT) + 127) & ~(size_t)127; }

template <int KT, bool HALO>
__global__ void __launch_bounds__(kPipeThreads, 1)
spmm_pipe_kernel(const BlockDesc* __restrict__ blocks, int32_t n_blocks, const int32_t* __restrict__ rowptr,
                 const double* __restrict__ vals, const uint16_t* __restrict__ lidx, const int32_t* __restrict__ ucols,
                 const double* __restrict__ x, int64_t ldx, const double* __restrict__ halo, int32_t halo_k, int32_t nloc,
                 double* __restrict__ y, int64_t ldy, int32_t k, int32_t cap, int32_t upad, double alpha, double beta) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
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

// TRANS / CONJ_TRANS (src/CSRMatrix.jl:1147-1160): Y[ind, v] += (alpha * X[row, v]) * val, after Y *= beta.
// Scatter with fp64 atomics (RED.ADD.F64): the summation order differs from the reference's row sweep, within
// the 1e-12 contract.  Warp per row, lanes over the row's entries.
__global__ void __launch_bounds__(kSpmvThreads)
spmv_trans_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind, const double* __restrict__ vals,
                  const double* __restrict__ x, int64_t ldx, double* __restrict__ y, int64_t ldy, int32_t n_rows, int32_t k, double alpha) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * kSpmvThreads + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * kSpmvThreads) >> 5;
  for (int64_t r = warp; r < n_rows; r += nwarps) {
    const int s = __ldg(rowptr + r), e = __ldg(rowptr + r + 1);
    for (int v = 0; v < k; ++v) {
      const double ax = __dmul_rn(alpha, __ldg(x + r + (size_t)v * ldx));
      for (int i = s + lane; i < e; i += 32) atomicAdd(y + __ldg(colind + i) + (size_t)v * ldy, __dmul_rn(ax, __ldg(vals + i)));
    }
  }
}

__global__ void shift_index_kernel(int32_t* __restrict__ a, int64_t n, int32_t base) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) a[i] -= base;
}

int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return (v && *v) ? std::atoi(v) : dflt;
}

size_t stage_smem_bytes(int cap) { return (size_t)(cap + 8) * 12 + (size_t)(kMaxRowsPerBlock + 8) * 4; }

// cut rows into blocks; cls[r] = 1 when the row touches a remote column
void build_blocks(const std::vector<int32_t>& rp, const std::vector<uint8_t>& cls, int32_t cap, int32_t max_rows, bool allow_direct,
                  std::vector<BlockDesc>& interior, std::vector<BlockDesc>& boundary, std::vector<BlockDesc>& all) {
  const int32_t n = (int32_t)rp.size() - 1;
  int32_t r = 0;
  while (r < n) {
    const uint8_t c = cls[r];
    const int32_t start = r, nnz0 = rp[r];
    int32_t maxlen = 0;
    int kind;
    if (rp[r + 1] - rp[r] > cap) {
      ++r;  // a block of its own, streamed from HBM
      kind = KIND_LONGROW;
    } else {
      while (r < n && cls[r] == c && r - start < max_rows && rp[r + 1] - nnz0 <= cap) {
        maxlen = std::max(maxlen, rp[r + 1] - rp[r]);
        ++r;
      }
      const int64_t nrows = r - start, nnzb = rp[r] - nnz0;
      // near-uniform short rows: max <= 2 * mean  ->  thread-per-row straight from the stage
      if (allow_direct && maxlen <= 32 && (int64_t)maxlen * nrows <= 2 * nnzb + 32) kind = KIND_DIRECT;
      else kind = maxlen > kLongRowLen ? KIND_WARPROW : KIND_TWOPHASE;
    }
    BlockDesc b;
    b.row0 = start;
    b.nrows_kind = (r - start) | (kind << 16);
    b.p0 = nnz0;
    b.p1 = rp[r];
    b.ucol0 = b.ucount = b.pad0 = b.pad1 = 0;
    (c ? boundary : interior).push_back(b);
    all.push_back(b);
  }
}

void upload_blocks(const std::vector<BlockDesc>& v, DeviceBuffer& buf, int32_t& count) {
  count = (int32_t)v.size();
  if (buf.bytes < v.size() * sizeof(BlockDesc) + 16) buf.alloc(v.size() * sizeof(BlockDesc) + 16);
  if (!v.empty()) JP_CUDA(cudaMemcpy(buf.p, v.data(), v.size() * sizeof(BlockDesc), cudaMemcpyHostToDevice));
}

// opt a kernel in to its dynamic shared-memory size once per size (the attribute call is not free)
template <class K>
void set_smem_attr(K kernel, size_t bytes) {
  static size_t configured = 0;  // one static per kernel instantiation
  if (configured == bytes) return;
  JP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  configured = bytes;
}

template <bool HALO, int TMA>
void launch_spmv_impl(const Csr& A, const BlockDesc* b, int32_t n_blocks, size_t smem, const double* x, const double* halo, double* y,
                      double alpha, double beta, cudaStream_t stream) {
  set_smem_attr(spmv_kernel<HALO, TMA>, smem);
  spmv_kernel<HALO, TMA><<<n_blocks, kSpmvThreads, smem, stream>>>(b, A.rowptr.as<int32_t>(), A.colind.as<int32_t>(), A.vals.as<double>(), x,
                                                                    halo, HALO ? A.n_local_cols : A.n_cols, y, A.nnz_cap, alpha, beta);
}

template <bool HALO, int TMA>
void launch_spmm_impl(const Csr& A, const BlockDesc* b, int32_t n_blocks, size_t smem, const double* x, int64_t ldx, const double* halo,
                      int32_t halo_k, double* y, int64_t ldy, int32_t k, double alpha, double beta, cudaStream_t stream) {
  constexpr int KT = 8;
  set_smem_attr(spmm_kernel<KT, HALO, TMA>, smem);
  spmm_kernel<KT, HALO, TMA><<<n_blocks, kSpmvThreads, smem, stream>>>(b, A.rowptr.as<int32_t>(), A.colind.as<int32_t>(),
                                                                        A.vals.as<double>(), x, ldx, halo, halo_k,
                                                                        HALO ? A.n_local_cols : A.n_cols, y, ldy, k, A.nnz_cap,
                                                                        A.threads_per_row, alpha, beta);
}


// Build the SpMM tile tables (lazily, at the first k > 1 apply): per block the sorted unique columns, per nonzero the
// position of its column in that list.  Host threads over blocks; one-time cost.
void build_tiles(Csr& A) {
  constexpr int KT = 8;
  const size_t nb = A.h_blocks_all.size();
  if (A.nnz == 0 || nb == 0 || env_int("JP_SPMM_TILE", 0) == 0) {  // opt-in: measured slower than the gather kernel (profiles/)
    A.tile_state = -1;
    return;
  }
  std::vector<int32_t> ci((size_t)A.nnz);
  JP_CUDA(cudaStreamSynchronize(ctx().compute));
  JP_CUDA(cudaMemcpy(ci.data(), A.colind.p, (size_t)A.nnz * 4, cudaMemcpyDeviceToHost));
  std::vector<int32_t> ubuf((size_t)A.nnz);       // block b's unique list is built in place at [p0, p0 + ucount)
  std::vector<uint16_t> lidx((size_t)A.nnz + 64, 0);
  std::vector<BlockDesc>& all = A.h_blocks_all;
  const int nthreads = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  bool has_longrow = false;
  auto worker = [&](int tix) {
    for (size_t bi = tix; bi < nb; bi += nthreads) {
      BlockDesc& b = all[bi];
      b.ucount = 0;
      if ((b.nrows_kind >> 16) == KIND_LONGROW) {
        has_longrow = true;  // (benign race: only ever set to true)
        continue;
      }
      int32_t* u = ubuf.data() + b.p0;
      const int32_t n = b.p1 - b.p0;
      std::copy(ci.begin() + b.p0, ci.begin() + b.p1, u);
      std::sort(u, u + n);
      const int32_t uc = (int32_t)(std::unique(u, u + n) - u);
      b.ucount = uc;
      for (int32_t i = b.p0; i < b.p1; ++i) lidx[i] = (uint16_t)(std::lower_bound(u, u + uc, ci[i]) - u);
    }
  };
  {
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(worker, t);
    worker(0);
    for (auto& t : th) t.join();
  }
  int64_t sum_u = 0, sum_nnz = 0, off = 0;
  int32_t umax = 0;
  for (BlockDesc& b : all) {
    b.ucol0 = (int32_t)off;
    off += (b.ucount + 3) & ~3;
    umax = std::max(umax, b.ucount);
    if (b.ucount) {
      sum_u += b.ucount;
      sum_nnz += b.p1 - b.p0;
    }
  }
  const int upad = (umax + 3) & ~3;
  // worth it only when columns are shared between the rows of a block, and the tile has to fit beside the stage
  if (has_longrow || sum_nnz == 0 || (double)sum_u > 0.7 * (double)sum_nnz || 2 * pipe_stage_bytes(A.nnz_cap, upad, KT) > 220 * 1024 ||
      off > INT32_MAX - 64) {
    A.tile_state = -1;
    return;
  }
  std::vector<int32_t> ucols((size_t)off + 64, 0);
  for (const BlockDesc& b : all) std::copy(ubuf.begin() + b.p0, ubuf.begin() + b.p0 + b.ucount, ucols.begin() + b.ucol0);
  // the interior / boundary lists are subsequences of the all-list: copy the patched fields over
  auto patch = [&](std::vector<BlockDesc>& lst) {
    size_t j = 0;
    for (BlockDesc& d : lst) {
      while (all[j].row0 != d.row0) ++j;
      d.ucol0 = all[j].ucol0;
      d.ucount = all[j].ucount;
    }
  };
  patch(A.h_blocks_interior);
  patch(A.h_blocks_boundary);
  A.ucols.alloc(ucols.size() * 4);
  A.lidx.alloc(lidx.size() * 2);
  JP_CUDA(cudaMemcpy(A.ucols.p, ucols.data(), ucols.size() * 4, cudaMemcpyHostToDevice));
  JP_CUDA(cudaMemcpy(A.lidx.p, lidx.data(), lidx.size() * 2, cudaMemcpyHostToDevice));
  upload_blocks(A.h_blocks_interior, A.blocks_interior, A.n_blocks_interior);
  upload_blocks(A.h_blocks_boundary, A.blocks_boundary, A.n_blocks_boundary);
  upload_blocks(A.h_blocks_all, A.blocks_all, A.n_blocks_all);
  A.tile_umax = umax;
  A.tile_state = 1;
}

template <bool HALO>
void launch_spmm_tile(const Csr& A, const BlockDesc* b, int32_t n_blocks, const double* x, int64_t ldx, const double* halo, int32_t halo_k,
                      double* y, int64_t ldy, int32_t k, double alpha, double beta, cudaStream_t stream) {
  constexpr int KT = 8;
  const int upad = (A.tile_umax + 3) & ~3;
  static const int pipe = env_int("JP_SPMM_PIPE", 1);
  if (pipe) {
    const size_t smem = 2 * pipe_stage_bytes(A.nnz_cap, upad, KT);
    set_smem_attr(spmm_pipe_kernel<KT, HALO>, smem);
    const int grid = std::min<int>(n_blocks, ctx().sm_count);
    spmm_pipe_kernel<KT, HALO><<<grid, kPipeThreads, smem, stream>>>(b, n_blocks, A.rowptr.as<int32_t>(), A.vals.as<double>(),
                                                                      A.lidx.as<uint16_t>(), A.ucols.as<int32_t>(), x, ldx, halo, halo_k,
                                                                      HALO ? A.n_local_cols : A.n_cols, y, ldy, k, A.nnz_cap, upad, alpha, beta);
    return;
  }
  const size_t smem = tile_smem_bytes(A.nnz_cap, upad, KT);
  set_smem_attr(spmm_tile_kernel<KT, HALO>, smem);
  spmm_tile_kernel<KT, HALO><<<n_blocks, kTileThreads, smem, stream>>>(b, A.rowptr.as<int32_t>(), A.colind.as<int32_t>(), A.vals.as<double>(),
                                                                        A.lidx.as<uint16_t>(), A.ucols.as<int32_t>(), x, ldx, halo, halo_k,
                                                                        HALO ? A.n_local_cols : A.n_cols, y, ldy, k, A.nnz_cap, upad, alpha, beta);
}

}  // namespace

void launch_spmv_fused(const Csr& A, const FusedHalo& h, const double* x, const double* halo, double* y, double alpha, double beta,
                       cudaStream_t stream) {
  FusedSpmvArgs a;
  a.interior = A.blocks_interior.as<BlockDesc>();
  a.boundary = A.blocks_boundary.as<BlockDesc>();
  a.n_interior = A.n_blocks_interior;
  a.n_boundary = A.n_blocks_boundary;
  a.n_pack = std::max(1, std::min(64, (h.n_export + 1023) / 1024));
  a.rowptr = A.rowptr.as<int32_t>();
  a.colind = A.colind.as<int32_t>();
  a.vals = A.vals.as<double>();
  a.x = x;
  a.halo = halo;
  a.nloc = A.n_local_cols;
  a.y = y;
  a.cap = A.nnz_cap;
  a.alpha = alpha;
  a.beta = beta;
  a.h = h;
  const size_t smem = stage_smem_bytes(A.nnz_cap);
  set_smem_attr(spmv_fused_halo_kernel, smem);
  const int grid = a.n_pack + a.n_interior + a.n_boundary;
  spmv_fused_halo_kernel<<<grid, kSpmvThreads, smem, stream>>>(a);
  JP_LAUNCH_CHECK();
}

void launch_spmv(const Csr& A, const DeviceBuffer& blocks, int32_t n_blocks, bool use_halo, const double* x, int64_t ldx,
                 const double* halo, int32_t halo_k, double* y, int64_t ldy, int32_t k, double alpha, double beta,
                 cudaStream_t stream) {
  launch_spmv_blocks(A, blocks.p, n_blocks, use_halo, x, ldx, halo, halo_k, y, ldy, k, alpha, beta, stream);
}

void launch_spmv_blocks(const Csr& A, const void* blocks_int4, int32_t n_blocks, bool use_halo, const double* x, int64_t ldx,
                        const double* halo, int32_t halo_k, double* y, int64_t ldy, int32_t k, double alpha, double beta,
                        cudaStream_t stream) {
  if (n_blocks == 0 || k == 0) return;
  if (k > 1 && A.tile_state == 0) {
    // lazily build the SpMM tile tables; the block lists are re-uploaded in place (same buffers, same order)
    build_tiles(const_cast<Csr&>(A));
  }
  const size_t smem = stage_smem_bytes(A.nnz_cap);
  const BlockDesc* b = static_cast<const BlockDesc*>(blocks_int4);
  static const int tma = env_int("JP_SPMV_TMA", 1);  // 0: cp.async staging (kept for A/B measurements)
  if (k == 1) {
    if (use_halo) {
      if (tma) launch_spmv_impl<true, 1>(A, b, n_blocks, smem, x, halo, y, alpha, beta, stream);
      else launch_spmv_impl<true, 0>(A, b, n_blocks, smem, x, halo, y, alpha, beta, stream);
    } else {
      if (tma) launch_spmv_impl<false, 1>(A, b, n_blocks, smem, x, nullptr, y, alpha, beta, stream);
      else launch_spmv_impl<false, 0>(A, b, n_blocks, smem, x, nullptr, y, alpha, beta, stream);
    }
  } else if (A.tile_state == 1) {
    if (use_halo) launch_spmm_tile<true>(A, b, n_blocks, x, ldx, halo, halo_k, y, ldy, k, alpha, beta, stream);
    else launch_spmm_tile<false>(A, b, n_blocks, x, ldx, nullptr, 1, y, ldy, k, alpha, beta, stream);
  } else {
    if (use_halo) {
      if (tma) launch_spmm_impl<true, 1>(A, b, n_blocks, smem, x, ldx, halo, halo_k, y, ldy, k, alpha, beta, stream);
      else launch_spmm_impl<true, 0>(A, b, n_blocks, smem, x, ldx, halo, halo_k, y, ldy, k, alpha, beta, stream);
    } else {
      if (tma) launch_spmm_impl<false, 1>(A, b, n_blocks, smem, x, ldx, nullptr, 1, y, ldy, k, alpha, beta, stream);
      else launch_spmm_impl<false, 0>(A, b, n_blocks, smem, x, ldx, nullptr, 1, y, ldy, k, alpha, beta, stream);
    }
  }
  JP_LAUNCH_CHECK();
}

void launch_spmv_trans(const Csr& A, const double* x, int64_t ldx, double* y, int64_t ldy, int32_t ny, int32_t k, double alpha,
                       double beta, cudaStream_t stream) {
  if (k == 0) return;
  // rawY[:, :] *= beta  (src/CSRMatrix.jl:1148), beta==0 overwrites
  if (ldy == ny) {
    launch_scale_fill(y, (int64_t)ny * k, beta, stream);
  } else {
    for (int v = 0; v < k; ++v) launch_scale_fill(y + (size_t)v * ldy, ny, beta, stream);
  }
  if (A.n_rows == 0 || A.nnz == 0) return;
  int64_t warps = A.n_rows;
  int64_t blocks = (warps * 32 + kSpmvThreads - 1) / kSpmvThreads;
  int64_t cap = (int64_t)ctx().sm_count * 16;
  if (blocks > cap) blocks = cap;
  spmv_trans_kernel<<<(int)blocks, kSpmvThreads, 0, stream>>>(A.rowptr.as<int32_t>(), A.colind.as<int32_t>(), A.vals.as<double>(), x, ldx,
                                                               y, ldy, A.n_rows, k, alpha);
  JP_LAUNCH_CHECK();
}


// Row blocks, SpMM lanes per row and the host-path chunk table from the row pointers and every row's largest column
// (rowmax[r] = -1 for an empty row).  Shared by jp_csr_create (host arrays) and jp_csr_fill_complete (device arrays).
void csr_finalize_structure(Csr& A, const std::vector<int32_t>& rp, const std::vector<int32_t>& rowmax) {
  const int32_t n_rows = A.n_rows;
  std::vector<uint8_t> cls((size_t)n_rows, 0);
  int32_t maxlen = 0;
  for (int32_t r = 0; r < n_rows; ++r) {
    cls[r] = rowmax[r] >= A.n_local_cols;
    maxlen = std::max(maxlen, rp[r + 1] - rp[r]);
  }
  A.max_row_len = maxlen;
  const double mean = n_rows ? (double)A.nnz / n_rows : 0.0;
  // SpMM lanes per row from the mean row length (power of two)
  int T = 1;
  while (T < 32 && mean > 12.0 * T) T <<= 1;
  T = env_int("JP_SPMM_T", T);
  JP_REQUIRE(T >= 1 && T <= 32 && (T & (T - 1)) == 0, "JP_SPMM_T must be a power of two in [1,32]");
  A.threads_per_row = T;
  int cap = env_int("JP_SPMV_CAP", 2048);
  JP_REQUIRE(cap >= 256 && cap <= 16384 && cap % 4 == 0, "JP_SPMV_CAP must be a multiple of 4 in [256,16384]");
  A.nnz_cap = cap;
  int max_rows = env_int("JP_SPMV_ROWS", kSpmvThreads);
  JP_REQUIRE(max_rows >= 1 && max_rows <= kMaxRowsPerBlock, "JP_SPMV_ROWS must be in [1,512]");
  std::vector<BlockDesc>&bi = A.h_blocks_interior, &bb = A.h_blocks_boundary, &ba = A.h_blocks_all;
  build_blocks(rp, cls, cap, max_rows, env_int("JP_SPMV_DIRECT", 1) != 0, bi, bb, ba);
  upload_blocks(bi, A.blocks_interior, A.n_blocks_interior);
  upload_blocks(bb, A.blocks_boundary, A.n_blocks_boundary);
  upload_blocks(ba, A.blocks_all, A.n_blocks_all);
  // chunks of the all-blocks list for the pipelined host-buffer path
  const int nchunks = std::max(1, std::min<int>(env_int("JP_HOST_CHUNKS", 16), (int)ba.size()));
  int32_t run = -1;
  for (int c = 0; c <= nchunks; ++c) {
    const int32_t bidx = (int32_t)((int64_t)ba.size() * c / nchunks);
    A.chunk_block.push_back(bidx);
    A.chunk_row.push_back(bidx < (int32_t)ba.size() ? ba[bidx].row0 : n_rows);
  }
  for (int c = 0; c < nchunks; ++c) {
    for (int32_t r = A.chunk_row[c]; r < A.chunk_row[c + 1]; ++r) run = std::max(run, rowmax[r]);
    A.chunk_maxcol.push_back(run);
  }
}

}  // namespace jp

using namespace jp;


extern "C" {

int32_t jp_csr_create(int32_t n_rows, int32_t n_cols, int32_t n_local_cols, int64_t nnz, const int32_t* rowptr, const int32_t* colind,
                      const double* vals, int32_t index_base, jp_handle* out) {
  return guarded([&] {
    require_init();
    JP_REQUIRE(out != nullptr, "jp_csr_create: null output");
    JP_REQUIRE(n_rows >= 0 && n_cols >= 0 && nnz >= 0, "jp_csr_create: negative size");
    JP_REQUIRE(n_local_cols >= 0 && n_local_cols <= n_cols, "jp_csr_create: n_local_cols must be in [0, n_cols]");
    JP_REQUIRE(index_base == 0 || index_base == 1, "jp_csr_create: index_base must be 0 or 1");
    JP_REQUIRE(nnz < ((int64_t)1 << 31) - 64, "jp_csr_create: local nnz must fit int32 (LID) row pointers");
    JP_REQUIRE(rowptr != nullptr && (nnz == 0 || (colind && vals)), "jp_csr_create: null array");
    auto A = std::make_unique<Csr>();
    A->n_rows = n_rows;
    A->n_cols = n_cols;
    A->n_local_cols = n_local_cols;
    A->nnz = nnz;
    // host pass: validate, classify rows, row-length statistics
    std::vector<int32_t> rp((size_t)n_rows + 1);
    JP_REQUIRE(rowptr[0] == index_base, "jp_csr_create: rowptr[first] must equal index_base");
    for (int32_t r = 0; r <= n_rows; ++r) {
      rp[r] = rowptr[r] - index_base;
      if (r > 0) JP_REQUIRE(rp[r] >= rp[r - 1], "jp_csr_create: rowptr must be non-decreasing");
    }
    JP_REQUIRE(rp[n_rows] == nnz, "jp_csr_create: rowptr[last] does not match nnz");
    std::vector<int32_t> rowmax((size_t)n_rows, -1);
    for (int32_t r = 0; r < n_rows; ++r) {
      int32_t mx = -1;
      for (int32_t i = rp[r]; i < rp[r + 1]; ++i) {
        const int32_t c = colind[i] - index_base;
        JP_REQUIRE(c >= 0 && c < n_cols, "jp_csr_create: column index outside the column map");
        mx = c > mx ? c : mx;
      }
      rowmax[r] = mx;
    }
    csr_finalize_structure(*A, rp, rowmax);
    // device arrays; 64 elements of zeroed slack so the 16-byte staging may over-read past nnz
    A->rowptr.alloc(((size_t)n_rows + 1 + 16) * 4);  // slack: row pointers are staged in aligned groups of 4
    JP_CUDA(cudaMemset(A->rowptr.p, 0, ((size_t)n_rows + 1 + 16) * 4));
    A->colind.alloc(((size_t)nnz + 64) * 4);
    A->vals.alloc(((size_t)nnz + 64) * 8);
    JP_CUDA(cudaMemcpy(A->rowptr.p, rp.data(), rp.size() * 4, cudaMemcpyHostToDevice));
    JP_CUDA(cudaMemset(A->colind.as<int32_t>() + nnz, 0, 64 * 4));
    JP_CUDA(cudaMemset(A->vals.as<double>() + nnz, 0, 64 * 8));
    if (nnz > 0) {
      JP_CUDA(cudaMemcpy(A->colind.p, colind, (size_t)nnz * 4, cudaMemcpyHostToDevice));
      JP_CUDA(cudaMemcpy(A->vals.p, vals, (size_t)nnz * 8, cudaMemcpyHostToDevice));
      if (index_base != 0) {
        shift_index_kernel<<<ctx().sm_count * 8, 256, 0, ctx().compute>>>(A->colind.as<int32_t>(), nnz, index_base);
        JP_LAUNCH_CHECK();
        JP_CUDA(cudaStreamSynchronize(ctx().compute));
      }
    }
    *out = register_object(std::move(A));
  });
}

int32_t jp_csr_destroy(jp_handle A) {
  return guarded([&] { destroy_object(A, Kind::Csr, "csr matrix"); });
}

int32_t jp_csr_info(jp_handle A, int64_t* info) {
  return guarded([&] {
    Csr* a = get<Csr>(A, Kind::Csr, "csr matrix");
    JP_REQUIRE(info != nullptr, "jp_csr_info: null output");
    info[0] = a->n_rows;
    info[1] = a->n_cols;
    info[2] = a->n_local_cols;
    info[3] = a->nnz;
    info[4] = a->n_blocks_interior;
    info[5] = a->n_blocks_boundary;
    info[6] = a->max_row_len;
    info[7] = a->threads_per_row;
  });
}

int32_t jp_csr_download(jp_handle A, int32_t* rowptr, int32_t* colind, double* vals) {
  return guarded([&] {
    require_init();
    Csr* a = get<Csr>(A, Kind::Csr, "csr matrix");
    JP_CUDA(cudaStreamSynchronize(ctx().compute));
    if (rowptr) JP_CUDA(cudaMemcpy(rowptr, a->rowptr.p, ((size_t)a->n_rows + 1) * 4, cudaMemcpyDeviceToHost));
    if (colind && a->nnz) JP_CUDA(cudaMemcpy(colind, a->colind.p, (size_t)a->nnz * 4, cudaMemcpyDeviceToHost));
    if (vals && a->nnz) JP_CUDA(cudaMemcpy(vals, a->vals.p, (size_t)a->nnz * 8, cudaMemcpyDeviceToHost));
  });
}

int32_t jp_local_apply(jp_handle Y, jp_handle A, jp_handle X, int32_t mode, double alpha, double beta) {
  return guarded([&] {
    require_init();
    Csr* a = get<Csr>(A, Kind::Csr, "csr matrix");
    MultiVector* y = get<MultiVector>(Y, Kind::MultiVector, "multivector");
    MultiVector* x = get<MultiVector>(X, Kind::MultiVector, "multivector");
    JP_REQUIRE(mode == JP_NO_TRANS || mode == JP_TRANS || mode == JP_CONJ_TRANS, "jp_local_apply: bad TransposeMode");
    JP_REQUIRE(x != y, "jp_local_apply: X and Y must be different multivectors");
    JP_REQUIRE(x->k == y->k, "jp_local_apply: X and Y must have the same number of vectors");
    cudaStream_t st = ctx().compute;
    if (mode == JP_NO_TRANS) {
      JP_REQUIRE(y->n == a->n_rows, "jp_local_apply: Y must have one row per matrix row");
      JP_REQUIRE(x->n == a->n_cols, "jp_local_apply: X must be a column-map multivector (n_cols rows)");
      launch_spmv(*a, a->blocks_all, a->n_blocks_all, false, x->d(), x->n, nullptr, 1, y->d(), y->n, y->k, alpha, beta, st);
    } else {
      JP_REQUIRE(x->n == a->n_rows, "jp_local_apply: X must have one row per matrix row in transposed mode");
      JP_REQUIRE(y->n == a->n_cols, "jp_local_apply: Y must be a column-map multivector in transposed mode");
      launch_spmv_trans(*a, x->d(), x->n, y->d(), y->n, y->n, y->k, alpha, beta, st);  // conjugation is the identity for reals
    }
  });
}

}  // extern "C"
