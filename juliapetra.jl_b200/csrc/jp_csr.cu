// jp_csr.cu -- device-resident LocalCSRMatrix (src/LocalCSRMatrix.jl:3-8, src/LocalCSRGraph.jl:11-14) and
// the localApply kernels (src/CSRMatrix.jl:1123-1162).
//
// Data layout in HBM: rowptr int32[n+1] (0-based), colind int32[nnz+pad], vals f64[nnz+pad]; colind/vals are
// streamed exactly once per apply.  At create time the rows are cut into ROW BLOCKS: contiguous row ranges whose
// nnz fit the shared-memory stage (nnz_cap) and whose row count fits the thread block; a row longer than the
// stage gets a block of its own.  Rows that touch only local columns ("interior") and rows that touch remote
// columns ("boundary") go to separate launch lists so the interior list can run while the halo is in flight.
//
// SpMV kernel (k == 1), one CTA per row block:
//   stage   colind/vals of the block are copied global->shared with 16-byte cp.async.cg (LDGSTS, L1 bypass),
//           fully coalesced irrespective of row boundaries; row pointers staged alongside.
//   phase A thread-per-nonzero: s_val[i] *= x[s_col[i]]  -- every gather of the block is independent, so the
//           memory-level parallelism does not depend on row lengths; x is read through the read-only path (L1/L2).
//   phase B thread-per-row (warp-per-row when the block holds long rows): sequential sum of the row's products
//           from shared memory, then y = beta*y + alpha*sum.  For thread-per-row blocks the summation order is
//           the reference's (left to right, unfused multiply/add), so the result is bit-identical to it.
// SpMM kernel (k > 1): same stage, then T lanes per row accumulate KT vectors at a time in registers; lanes of a
// warp map to consecutive rows so the gathers of a banded matrix are coalesced in a column-major X; the matrix
// is read from HBM once for all k (the reference re-streams it per vector).
// No tensor cores: this is not a dense contraction.  No atomics in NO_TRANS: results are deterministic.
#include <algorithm>
#include <cstdlib>

#include "jp_common.h"

namespace jp {

namespace {

constexpr int kSpmvThreads = 256;
constexpr int kMaxRowsPerBlock = 512;   // rows staged per block (row pointers in shared memory)
constexpr int kLongRowLen = 64;         // blocks holding a row longer than this reduce warp-per-row

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n" ::: "memory");
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
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

// copy nnz [a0, p1) of the block and its row pointers into shared memory
__device__ __forceinline__ void stage_block(const StageView& sv, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                                            const double* __restrict__ vals, int row0, int nrows, int a0, int p1) {
  const int tid = threadIdx.x;
  const int ngrp = (p1 - a0 + 3) >> 2;
  for (int g = tid; g < ngrp; g += kSpmvThreads) {
    cp_async16(sv.s_col + 4 * g, colind + a0 + 4 * g);
    cp_async16(sv.s_val + 4 * g, vals + a0 + 4 * g);
    cp_async16(sv.s_val + 4 * g + 2, vals + a0 + 4 * g + 2);
  }
  for (int r = tid; r <= nrows; r += kSpmvThreads) sv.s_rp[r] = __ldg(rowptr + row0 + r) - a0;
  cp_async_wait_all();
  __syncthreads();
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

template <bool HALO>
__global__ void __launch_bounds__(kSpmvThreads)
spmv_kernel(const int4* __restrict__ blocks, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
            const double* __restrict__ vals, const double* __restrict__ x, const double* __restrict__ halo, int32_t nloc,
            double* __restrict__ y, int32_t cap, double alpha, double beta) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double s_red[kSpmvThreads / 32];
  const StageView sv = stage_view(smem_raw, cap);
  const int4 blk = blocks[blockIdx.x];
  const int row0 = blk.x, nrows = blk.y - blk.x, p0 = blk.z;
  const bool longrows = blk.w < 0;
  const int p1 = longrows ? -blk.w - 1 : blk.w;
  const int tid = threadIdx.x;
  if (p1 - p0 > cap) {  // one row, longer than the stage
    long_row<HALO>(colind, vals, x, 0, halo, 1, nloc, y, 0, 1, row0, p0, p1, alpha, beta, s_red);
    return;
  }
  const int a0 = p0 & ~3;
  stage_block(sv, rowptr, colind, vals, row0, nrows, a0, p1);
  // phase A: products, thread-per-nonzero
  const int s0 = p0 - a0, s1 = p1 - a0;
#pragma unroll 4
  for (int i = s0 + tid; i < s1; i += kSpmvThreads) {
    const int c = sv.s_col[i];
    const double xv = (HALO && c >= nloc) ? halo[c - nloc] : __ldg(x + c);
    sv.s_val[i] = __dmul_rn(sv.s_val[i], xv);
  }
  __syncthreads();
  // phase B: row sums
  if (!longrows) {
    for (int r = tid; r < nrows; r += kSpmvThreads) {
      const int s = sv.s_rp[r], e = sv.s_rp[r + 1];
      double sum = 0.0;
      for (int i = s; i < e; ++i) sum = __dadd_rn(sum, sv.s_val[i]);
      double* yp = y + row0 + r;
      *yp = combine(beta == 0.0 ? 0.0 : *yp, sum, alpha, beta);
    }
  } else {
    const int lane = tid & 31, warp = tid >> 5;
    for (int r = warp; r < nrows; r += kSpmvThreads / 32) {
      const int s = sv.s_rp[r], e = sv.s_rp[r + 1];
      double sum = 0.0;
      for (int i = s + lane; i < e; i += 32) sum += sv.s_val[i];
      sum = warp_sum(sum);
      if (lane == 0) {
        double* yp = y + row0 + r;
        *yp = combine(beta == 0.0 ? 0.0 : *yp, sum, alpha, beta);
      }
    }
  }
}

template <int KT, bool HALO>
__global__ void __launch_bounds__(kSpmvThreads)
spmm_kernel(const int4* __restrict__ blocks, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
            const double* __restrict__ vals, const double* __restrict__ x, int64_t ldx, const double* __restrict__ halo,
            int32_t halo_k, int32_t nloc, double* __restrict__ y, int64_t ldy, int32_t k, int32_t cap, int32_t T, double alpha,
            double beta) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double s_red[kSpmvThreads / 32];
  const StageView sv = stage_view(smem_raw, cap);
  const int4 blk = blocks[blockIdx.x];
  const int row0 = blk.x, nrows = blk.y - blk.x, p0 = blk.z;
  const int p1 = blk.w < 0 ? -blk.w - 1 : blk.w;
  const int tid = threadIdx.x;
  if (p1 - p0 > cap) {
    long_row<HALO>(colind, vals, x, ldx, halo, halo_k, nloc, y, ldy, k, row0, p0, p1, alpha, beta, s_red);
    return;
  }
  const int a0 = p0 & ~3;
  stage_block(sv, rowptr, colind, vals, row0, nrows, a0, p1);
  const int lane_t = tid & (T - 1), group = tid / T, ngroups = kSpmvThreads / T;
  for (int v0 = 0; v0 < k; v0 += KT) {
    for (int rbase = 0; rbase < nrows; rbase += ngroups) {  // warp-uniform trip count (shuffles below)
      const int r = rbase + group;
      const bool valid = r < nrows;
      const int s = valid ? sv.s_rp[r] : 0, e = valid ? sv.s_rp[r + 1] : 0;
      double acc[KT];
#pragma unroll
      for (int j = 0; j < KT; ++j) acc[j] = 0.0;
      for (int i = s + lane_t; i < e; i += T) {
        const int c = sv.s_col[i];
        const double v = sv.s_val[i];
        if (HALO && c >= nloc) {
          const double* h = halo + (size_t)(c - nloc) * halo_k + v0;
#pragma unroll
          for (int j = 0; j < KT; ++j)
            if (v0 + j < k) acc[j] = __dadd_rn(acc[j], __dmul_rn(v, h[j]));
        } else {
          const double* xc = x + c + (size_t)v0 * ldx;
#pragma unroll
          for (int j = 0; j < KT; ++j)
            if (v0 + j < k) acc[j] = __dadd_rn(acc[j], __dmul_rn(v, __ldg(xc + (size_t)j * ldx)));
        }
      }
      for (int o = T >> 1; o > 0; o >>= 1) {
#pragma unroll
        for (int j = 0; j < KT; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
      }
      if (valid && lane_t == 0) {
#pragma unroll
        for (int j = 0; j < KT; ++j)
          if (v0 + j < k) {
            double* yp = y + row0 + r + (size_t)(v0 + j) * ldy;
            *yp = combine(beta == 0.0 ? 0.0 : *yp, acc[j], alpha, beta);
          }
      }
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

size_t stage_smem_bytes(int cap) { return (size_t)(cap + 8) * 12 + (size_t)(kMaxRowsPerBlock + 1) * 4; }

// cut rows into blocks; cls[r] = 1 when the row touches a remote column
void build_blocks(const std::vector<int32_t>& rp, const std::vector<uint8_t>& cls, int32_t cap, int32_t max_rows,
                  std::vector<int4>& interior, std::vector<int4>& boundary, std::vector<int4>& all) {
  const int32_t n = (int32_t)rp.size() - 1;
  int32_t r = 0;
  while (r < n) {
    const uint8_t c = cls[r];
    const int32_t start = r, nnz0 = rp[r];
    int32_t maxlen = 0;
    if (rp[r + 1] - rp[r] > cap) {
      maxlen = rp[r + 1] - rp[r];
      ++r;  // a block of its own, streamed from HBM
    } else {
      while (r < n && cls[r] == c && r - start < max_rows && rp[r + 1] - nnz0 <= cap) {
        maxlen = std::max(maxlen, rp[r + 1] - rp[r]);
        ++r;
      }
    }
    int4 b;
    b.x = start;
    b.y = r;
    b.z = nnz0;
    b.w = maxlen > kLongRowLen ? -rp[r] - 1 : rp[r];
    (c ? boundary : interior).push_back(b);
    all.push_back(b);
  }
}

void upload_blocks(const std::vector<int4>& v, DeviceBuffer& buf, int32_t& count) {
  count = (int32_t)v.size();
  buf.alloc(v.size() * sizeof(int4) + 16);
  if (!v.empty()) JP_CUDA(cudaMemcpy(buf.p, v.data(), v.size() * sizeof(int4), cudaMemcpyHostToDevice));
}

// opt a kernel in to its dynamic shared-memory size once per size (the attribute call is not free)
template <class K>
void set_smem_attr(K kernel, size_t bytes) {
  static size_t configured = 0;  // one static per kernel instantiation
  if (configured == bytes) return;
  JP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  configured = bytes;
}

}  // namespace

void launch_spmv(const Csr& A, const DeviceBuffer& blocks, int32_t n_blocks, bool use_halo, const double* x, int64_t ldx,
                 const double* halo, int32_t halo_k, double* y, int64_t ldy, int32_t k, double alpha, double beta,
                 cudaStream_t stream) {
  if (n_blocks == 0 || k == 0) return;
  const size_t smem = stage_smem_bytes(A.nnz_cap);
  const int4* b = blocks.as<int4>();
  const int32_t* rp = A.rowptr.as<int32_t>();
  const int32_t* ci = A.colind.as<int32_t>();
  const double* va = A.vals.as<double>();
  if (k == 1) {
    if (use_halo) {
      set_smem_attr(spmv_kernel<true>, smem);
      spmv_kernel<true><<<n_blocks, kSpmvThreads, smem, stream>>>(b, rp, ci, va, x, halo, A.n_local_cols, y, A.nnz_cap, alpha, beta);
    } else {
      set_smem_attr(spmv_kernel<false>, smem);
      spmv_kernel<false><<<n_blocks, kSpmvThreads, smem, stream>>>(b, rp, ci, va, x, nullptr, A.n_cols, y, A.nnz_cap, alpha, beta);
    }
  } else {
    const int32_t T = A.threads_per_row;
    constexpr int KT = 8;
    if (use_halo) {
      set_smem_attr(spmm_kernel<KT, true>, smem);
      spmm_kernel<KT, true><<<n_blocks, kSpmvThreads, smem, stream>>>(b, rp, ci, va, x, ldx, halo, halo_k, A.n_local_cols, y, ldy, k,
                                                                       A.nnz_cap, T, alpha, beta);
    } else {
      set_smem_attr(spmm_kernel<KT, false>, smem);
      spmm_kernel<KT, false><<<n_blocks, kSpmvThreads, smem, stream>>>(b, rp, ci, va, x, ldx, nullptr, 1, A.n_cols, y, ldy, k, A.nnz_cap,
                                                                        T, alpha, beta);
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
    std::vector<uint8_t> cls((size_t)n_rows, 0);
    int32_t maxlen = 0;
    for (int32_t r = 0; r < n_rows; ++r) {
      int32_t mx = -1;
      for (int32_t i = rp[r]; i < rp[r + 1]; ++i) {
        const int32_t c = colind[i] - index_base;
        JP_REQUIRE(c >= 0 && c < n_cols, "jp_csr_create: column index outside the column map");
        mx = c > mx ? c : mx;
      }
      cls[r] = mx >= n_local_cols;
      maxlen = std::max(maxlen, rp[r + 1] - rp[r]);
    }
    A->max_row_len = maxlen;
    const double mean = n_rows ? (double)nnz / n_rows : 0.0;
    // SpMM lanes per row from the mean row length (power of two); SpMV always uses the two-phase kernel
    int T = 1;
    while (T < 32 && mean > 12.0 * T) T <<= 1;
    T = env_int("JP_SPMM_T", T);
    JP_REQUIRE(T >= 1 && T <= 32 && (T & (T - 1)) == 0, "JP_SPMM_T must be a power of two in [1,32]");
    A->threads_per_row = T;
    int cap = env_int("JP_SPMV_CAP", 2048);
    JP_REQUIRE(cap >= 256 && cap <= 16384 && cap % 4 == 0, "JP_SPMV_CAP must be a multiple of 4 in [256,16384]");
    A->nnz_cap = cap;
    int max_rows = env_int("JP_SPMV_ROWS", kSpmvThreads);
    JP_REQUIRE(max_rows >= 1 && max_rows <= kMaxRowsPerBlock, "JP_SPMV_ROWS must be in [1,512]");
    std::vector<int4> bi, bb, ba;
    build_blocks(rp, cls, cap, max_rows, bi, bb, ba);
    upload_blocks(bi, A->blocks_interior, A->n_blocks_interior);
    upload_blocks(bb, A->blocks_boundary, A->n_blocks_boundary);
    upload_blocks(ba, A->blocks_all, A->n_blocks_all);
    // device arrays; 64 elements of zeroed slack so the 16-byte staging may over-read past nnz
    A->rowptr.alloc(((size_t)n_rows + 1) * 4);
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
