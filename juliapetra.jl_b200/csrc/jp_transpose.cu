// jp_transpose.cu -- the transposed local matrix behind apply!(…, TRANS / CONJ_TRANS) (src/CSRMatrix.jl:1147-1160,
// src/RowMatrix.jl:324-352).
//
// The reference's transposed localApply sweeps the rows of A and scatters  Y[col] += (alpha * X[row]) * val.  A scatter
// needs fp64 atomics on a GPU, is not deterministic, and a warp per 7-entry row leaves 25 of 32 lanes idle.  Instead the
// transposed matrix A^T is built ONCE, at the first transposed apply, as a device CSR of its own (rows = A's column-map
// LIDs, columns = A's rows) with its own row blocks, and every transposed apply is the ordinary row-block SpMV / SpMM on
// it: no atomics, deterministic, HBM-roofline bound like the forward apply.  Inside a row of A^T the entries are in
// ascending row order of A, which is the order in which the reference's row sweep adds them to Y[col].
//
// Construction (one-time, stream-ordered): a STABLE radix sort of the nonzero positions by column index
// (cub::DeviceRadixSort -- CUB ships with the CUDA toolkit; this is setup, not the hot path), then three small kernels:
// row pointers of A^T from the sorted column keys, row index of every nonzero of A, gather of (row, value).
// Cost: one extra copy of the matrix in HBM (12 bytes per nonzero), only for matrices that are applied transposed.
#include <cub/device/device_radix_sort.cuh>

#include "jp_common.h"

namespace jp {

namespace {

constexpr int kThreads = 256;

int grid_for(int64_t n) {
  int64_t want = (n + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)ctx().sm_count * 16;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}

__global__ void __launch_bounds__(kThreads) iota_kernel(uint32_t* __restrict__ a, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) a[i] = (uint32_t)i;
}

// row_of[i] = the row of A that holds nonzero i (thread per row; stencil rows are short, long rows are rare)
__global__ void __launch_bounds__(kThreads) expand_rows_kernel(const int32_t* __restrict__ rowptr, int32_t n_rows, int32_t* __restrict__ row_of) {
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t r = (int64_t)blockIdx.x * kThreads + threadIdx.x; r < n_rows; r += stride)
    for (int32_t i = rowptr[r]; i < rowptr[r + 1]; ++i) row_of[i] = (int32_t)r;
}

// rowptrT[c] = first position of column c in the sorted key array (keys ascending); rowptrT[n_cols] = nnz
__global__ void __launch_bounds__(kThreads) transposed_rowptr_kernel(const int32_t* __restrict__ keys, int64_t nnz, int32_t n_cols,
                                                                       int32_t* __restrict__ rowptrT) {
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i <= nnz; i += stride) {
    const int32_t lo = i == 0 ? -1 : keys[i - 1];
    const int32_t hi = i == nnz ? n_cols : keys[i];
    for (int32_t c = lo + 1; c <= hi; ++c) rowptrT[c] = (int32_t)i;  // columns (lo, hi] start at i (empty ones included)
  }
}

__global__ void __launch_bounds__(kThreads) transposed_gather_kernel(const uint32_t* __restrict__ order, const int32_t* __restrict__ row_of,
                                                                       const double* __restrict__ vals, int64_t nnz,
                                                                       int32_t* __restrict__ colindT, double* __restrict__ valsT) {
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < nnz; i += stride) {
    const uint32_t src = order[i];
    colindT[i] = __ldg(row_of + src);
    valsT[i] = __ldg(vals + src);
  }
}

// rowmax[c] = largest column index of row c of A^T (its last entry: rows of A ascend inside a row of A^T), -1 when empty
__global__ void __launch_bounds__(kThreads) transposed_rowmax_kernel(const int32_t* __restrict__ rowptrT, const int32_t* __restrict__ colindT,
                                                                       int32_t n, int32_t* __restrict__ rowmax) {
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x; c < n; c += stride)
    rowmax[c] = rowptrT[c + 1] > rowptrT[c] ? colindT[rowptrT[c + 1] - 1] : -1;
}

}  // namespace

// A.transposed, built on first use (compute stream)
Csr& csr_transposed(Csr& A) {
  if (A.transposed) return *A.transposed;
  cudaStream_t st = ctx().compute;
  auto T = std::make_unique<Csr>();
  const int64_t nnz = A.nnz;
  T->n_rows = A.n_cols;
  T->n_cols = A.n_rows;
  T->n_local_cols = A.n_rows;  // a transposed apply reads X on A's row map: there is no halo on this side
  T->nnz = nnz;
  T->rowptr.alloc(((size_t)T->n_rows + 1 + 16) * 4);
  JP_CUDA(cudaMemsetAsync(T->rowptr.p, 0, ((size_t)T->n_rows + 1 + 16) * 4, st));
  T->colind.alloc(((size_t)nnz + 64) * 4);
  T->vals.alloc(((size_t)nnz + 64) * 8);
  JP_CUDA(cudaMemsetAsync(T->colind.as<int32_t>() + nnz, 0, 64 * 4, st));
  JP_CUDA(cudaMemsetAsync(T->vals.as<double>() + nnz, 0, 64 * 8, st));
  std::vector<int32_t> rp((size_t)T->n_rows + 1, 0), rowmax((size_t)T->n_rows, -1);
  if (nnz > 0) {
    DeviceBuffer keys_out, order_in, order_out, row_of, tmp, d_rowmax;
    keys_out.alloc((size_t)nnz * 4);
    order_in.alloc((size_t)nnz * 4);
    order_out.alloc((size_t)nnz * 4);
    row_of.alloc((size_t)nnz * 4);
    iota_kernel<<<grid_for(nnz), kThreads, 0, st>>>(order_in.as<uint32_t>(), nnz);
    JP_LAUNCH_CHECK();
    expand_rows_kernel<<<grid_for(A.n_rows), kThreads, 0, st>>>(A.rowptr.as<int32_t>(), A.n_rows, row_of.as<int32_t>());
    JP_LAUNCH_CHECK();
    int end_bit = 1;
    while (end_bit < 31 && ((int64_t)1 << end_bit) < (int64_t)A.n_cols) ++end_bit;
    size_t tmp_bytes = 0;
    JP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, A.colind.as<int32_t>(), keys_out.as<int32_t>(), order_in.as<uint32_t>(),
                                            order_out.as<uint32_t>(), nnz, 0, end_bit, st));
    tmp.alloc(tmp_bytes);
    JP_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, A.colind.as<int32_t>(), keys_out.as<int32_t>(), order_in.as<uint32_t>(),
                                            order_out.as<uint32_t>(), nnz, 0, end_bit, st));
    count_launch();
    transposed_rowptr_kernel<<<grid_for(nnz + 1), kThreads, 0, st>>>(keys_out.as<int32_t>(), nnz, A.n_cols, T->rowptr.as<int32_t>());
    JP_LAUNCH_CHECK();
    transposed_gather_kernel<<<grid_for(nnz), kThreads, 0, st>>>(order_out.as<uint32_t>(), row_of.as<int32_t>(), A.vals.as<double>(), nnz,
                                                                 T->colind.as<int32_t>(), T->vals.as<double>());
    JP_LAUNCH_CHECK();
    d_rowmax.alloc((size_t)T->n_rows * 4 + 16);
    if (T->n_rows > 0) {
      transposed_rowmax_kernel<<<grid_for(T->n_rows), kThreads, 0, st>>>(T->rowptr.as<int32_t>(), T->colind.as<int32_t>(), T->n_rows,
                                                                         d_rowmax.as<int32_t>());
      JP_LAUNCH_CHECK();
      JP_CUDA(cudaMemcpyAsync(rowmax.data(), d_rowmax.p, (size_t)T->n_rows * 4, cudaMemcpyDeviceToHost, st));
    }
    JP_CUDA(cudaMemcpyAsync(rp.data(), T->rowptr.p, rp.size() * 4, cudaMemcpyDeviceToHost, st));
    JP_CUDA(cudaStreamSynchronize(st));  // the scratch buffers go out of scope here
  }
  csr_finalize_structure(*T, rp, rowmax);
  A.transposed = std::move(T);
  return *A.transposed;
}

}  // namespace jp
