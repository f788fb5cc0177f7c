// jp_colblock.cu -- the column-slab copy of a large irregular matrix and the launch sequence of spmv_cb_kernel (see the
// header comment of jp_cb_kernels.cuh for the why and the layout).  Decided once per matrix, at its first k = 1 apply over
// the whole block list without halo columns:
//   * most of the nonzeros are in irregular row blocks (stencil blocks gather coalesced: nothing to gain),
//   * the multivector is larger than two slabs (x of a smaller matrix lives in the L2 anyway).
// One-time cost: a stable radix sort of the nonzero positions by slab (cub::DeviceRadixSort -- setup, not the hot path)
// and three gathers; memory: 16 bytes per nonzero + 24 bytes per 32 nonzeros of carries.
#include <cub/device/device_radix_sort.cuh>

#include <cstdlib>

#include "jp_cb_kernels.cuh"
#include "jp_common.h"
#include "jp_csr_kernels.cuh"

namespace jp {

using namespace cbk;

struct CbData {
  int32_t n_slabs = 0, slab_cols = 0;
  DeviceBuffer row, col, val, carry;
  std::vector<long long> slab_ptr;  // n_slabs + 1
};

void CbDeleter::operator()(CbData* p) const { delete p; }

namespace {
int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return (v && *v) ? std::atoi(v) : dflt;
}
int grid_for(int64_t n) {
  int64_t want = (n + kCbThreads - 1) / kCbThreads;
  const int64_t cap = (int64_t)ctx().sm_count * 16;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}
}  // namespace

void csr_prepare_colblock(Csr& A) {
  if (A.cb_state != 0) return;
  A.cb_state = -1;
  // slab size: 16 / 24 / 32 MB measured 1.285 / 1.259 / 1.224 ms on the 10 M-row power-law matrix (profiles/r2_powerlaw_colblock.json)
  const int slab_mb = env_int("JP_CB_SLAB_MB", 32);  // 0: never (A/B measurements and the tests of the row-block kernel)
  if (slab_mb <= 0 || A.nnz == 0 || A.n_cols != A.n_local_cols) return;
  const int64_t slab_bytes = (int64_t)slab_mb << 20;
  if ((int64_t)A.n_cols * 8 <= 2 * slab_bytes) return;  // x (nearly) fits the L2 as it is
  int64_t irregular = 0;
  for (const BlockDesc& b : A.h_blocks_all)
    if (((b.nrows_kind >> 16) & 0xf) != csrk::KIND_DIRECT) irregular += b.p1 - b.p0;
  if (2 * irregular < A.nnz) return;
  cudaStream_t st = ctx().compute;
  auto cb = std::unique_ptr<CbData, CbDeleter>(new CbData());
  const int64_t want_cols = slab_bytes / 8;
  cb->n_slabs = (int32_t)((A.n_cols + want_cols - 1) / want_cols);
  cb->slab_cols = (int32_t)(((int64_t)A.n_cols + cb->n_slabs - 1) / cb->n_slabs);
  if (cb->n_slabs < 2 || cb->n_slabs > 4096) return;
  const int64_t nnz = A.nnz;
  DeviceBuffer key_in, key_out, order_in, order_out, row_of, tmp, d_slab_ptr;
  key_in.alloc((size_t)nnz * 4);
  key_out.alloc((size_t)nnz * 4);
  order_in.alloc((size_t)nnz * 4);
  order_out.alloc((size_t)nnz * 4);
  row_of.alloc((size_t)nnz * 4);
  cb_keys_kernel<<<grid_for(nnz), kCbThreads, 0, st>>>(A.colind.as<int32_t>(), nnz, cb->slab_cols, key_in.as<uint32_t>(), order_in.as<uint32_t>());
  JP_LAUNCH_CHECK();
  cb_expand_rows_kernel<<<grid_for(A.n_rows), kCbThreads, 0, st>>>(A.rowptr.as<int32_t>(), A.n_rows, row_of.as<int32_t>());
  JP_LAUNCH_CHECK();
  int end_bit = 1;
  while (end_bit < 31 && (1 << end_bit) < cb->n_slabs) ++end_bit;
  size_t tmp_bytes = 0;
  JP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key_in.as<uint32_t>(), key_out.as<uint32_t>(), order_in.as<uint32_t>(),
                                          order_out.as<uint32_t>(), nnz, 0, end_bit, st));
  tmp.alloc(tmp_bytes);
  JP_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, key_in.as<uint32_t>(), key_out.as<uint32_t>(), order_in.as<uint32_t>(),
                                          order_out.as<uint32_t>(), nnz, 0, end_bit, st));
  count_launch();
  cb->row.alloc((size_t)nnz * 4 + 64);
  cb->col.alloc((size_t)nnz * 4 + 64);
  cb->val.alloc((size_t)nnz * 8 + 64);
  cb_gather_kernel<<<grid_for(nnz), kCbThreads, 0, st>>>(order_out.as<uint32_t>(), row_of.as<int32_t>(), A.colind.as<int32_t>(), A.vals.as<double>(),
                                                        nnz, cb->row.as<int32_t>(), cb->col.as<int32_t>(), cb->val.as<double>());
  JP_LAUNCH_CHECK();
  d_slab_ptr.alloc(((size_t)cb->n_slabs + 1) * 8);
  cb_slab_ptr_kernel<<<grid_for(nnz + 1), kCbThreads, 0, st>>>(key_out.as<uint32_t>(), nnz, cb->n_slabs, d_slab_ptr.as<long long>());
  JP_LAUNCH_CHECK();
  cb->slab_ptr.resize((size_t)cb->n_slabs + 1);
  JP_CUDA(cudaMemcpyAsync(cb->slab_ptr.data(), d_slab_ptr.p, ((size_t)cb->n_slabs + 1) * 8, cudaMemcpyDeviceToHost, st));
  JP_CUDA(cudaStreamSynchronize(st));  // the scratch buffers go out of scope here
  int64_t max_slab = 0;
  for (int s = 0; s < cb->n_slabs; ++s) max_slab = std::max<int64_t>(max_slab, cb->slab_ptr[s + 1] - cb->slab_ptr[s]);
  cb->carry.alloc((size_t)(2 * ((max_slab + 31) / 32) + 2) * sizeof(CbCarry));
  A.cb = std::move(cb);
  A.cb_state = 1;
}

// y = beta*y + alpha*A*x (k = 1, the whole matrix, no halo columns), slab by slab
void launch_spmv_colblock(const Csr& A, const double* x, double* y, double alpha, double beta, cudaStream_t stream) {
  const CbData& cb = *A.cb;
  launch_scale_fill(y, A.n_rows, beta, stream);
  for (int s = 0; s < cb.n_slabs; ++s) {
    const int64_t e0 = cb.slab_ptr[s], e1 = cb.slab_ptr[s + 1];
    if (e1 <= e0) continue;
    const int64_t passes = (e1 - e0 + 31) / 32;
    const int grid = (int)((e1 - e0 + kCbPerBlock - 1) / kCbPerBlock);
    spmv_cb_kernel<<<grid, kCbThreads, 0, stream>>>(cb.row.as<int32_t>(), cb.col.as<int32_t>(), cb.val.as<double>(), e0, e1, x, y, alpha,
                                                    cb.carry.as<CbCarry>());
    JP_LAUNCH_CHECK();
    spmv_cb_carry_kernel<<<grid_for(2 * passes), kCbThreads, 0, stream>>>(cb.carry.as<CbCarry>(), 2 * passes, y, alpha);
    JP_LAUNCH_CHECK();
  }
}

int64_t csr_colblock_info(const Csr& A, int what) {
  if (A.cb_state != 1) return 0;
  return what == 0 ? A.cb->n_slabs : A.cb->slab_cols;
}

}  // namespace jp
