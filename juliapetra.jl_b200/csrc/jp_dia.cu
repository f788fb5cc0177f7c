// jp_dia.cu -- the diagonal-major copy of a stencil-like matrix and the launch of spmm_dia_kernel (see the kernel's
// header comment in jp_csr_kernels.cuh for the why).  Built lazily at the first k > 1 apply of a matrix that
//   * has no remote columns in the launch (single rank, or the matrix is applied through its whole-block list),
//   * has at most kDiaMaxDiags distinct values of col - row, found on the device with a bitmap,
//   * would be padded by less than 25 % (n_rows * D <= 1.25 * nnz).
// One-time cost: one pass over the column indices, one over the values; memory: 8 bytes per (row, diagonal).
#include <cstdlib>

#include "jp_common.h"
#include "jp_dia_plan.h"

namespace jp {

using namespace csrk;

namespace {
int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return (v && *v) ? std::atoi(v) : dflt;
}

template <class K>
void set_smem_attr(K kernel, size_t bytes) {
  static size_t configured = 0;
  if (configured == bytes) return;
  JP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  configured = bytes;
}
}  // namespace

struct DiaData {
  DiaParams P;
  DeviceBuffer vals;  // [D][npad]
  int64_t npad = 0;
  int32_t row_lo = 0, row_hi = 0;  // rows served by the diagonal kernel
  int32_t b_lo = 0, b_hi = 0;      // all-blocks list: [0, b_lo) and [b_hi, n) go to the gather kernel (edge rows)
};

void DiaDeleter::operator()(DiaData* p) const { delete p; }

void csr_prepare_dia(Csr& A) {
  if (A.dia_state != 0) return;
  A.dia_state = -1;
  if (env_int("JP_SPMM_DIA", 1) == 0) return;  // A/B measurements and the tests of the run-staged path
  if (A.nnz == 0 || A.n_rows < 4 * kDiaThreads || A.n_cols != A.n_local_cols || A.max_row_len > kDiaMaxDiags) return;
  if ((double)A.n_rows * A.max_row_len > 1.25 * (double)A.nnz) return;  // D >= max_row_len: too much padding already
  cudaStream_t st = ctx().compute;
  // 1. distinct offsets
  const size_t nbits = (size_t)A.n_rows + A.n_cols, nwords = (nbits + 31) / 32;
  DeviceBuffer bitmap;
  bitmap.alloc(nwords * 4);
  JP_CUDA(cudaMemsetAsync(bitmap.p, 0, nwords * 4, st));
  const int grid = (int)std::min<int64_t>(((int64_t)A.n_rows + 255) / 256, (int64_t)ctx().sm_count * 16);
  dia_mark_kernel<<<grid, 256, 0, st>>>(A.rowptr.as<int32_t>(), A.colind.as<int32_t>(), A.n_rows, bitmap.as<uint32_t>());
  JP_LAUNCH_CHECK();
  std::vector<uint32_t> h(nwords);
  JP_CUDA(cudaMemcpyAsync(h.data(), bitmap.p, nwords * 4, cudaMemcpyDeviceToHost, st));
  JP_CUDA(cudaStreamSynchronize(st));
  std::vector<int32_t> offsets;
  for (size_t w = 0; w < nwords; ++w) {
    uint32_t bits = h[w];
    while (bits) {
      const int b = __builtin_ctz(bits);
      bits &= bits - 1;
      offsets.push_back((int32_t)((int64_t)w * 32 + b - (A.n_rows - 1)));
      if ((int)offsets.size() > kDiaMaxDiags) return;
    }
  }
  const int D = (int)offsets.size();
  if (D == 0 || (double)A.n_rows * D > 1.25 * (double)A.nnz) return;
  auto dia = std::unique_ptr<DiaData, DiaDeleter>(new DiaData());
  // The diagonal kernel serves whole CSR row blocks [b_lo, b_hi) of the all-blocks list; the edge blocks go to the gather
  // kernel.  The first served row fixes the parity of every kernel block's first row, hence the alignment of the windows.
  const std::vector<BlockDesc>& ba = A.h_blocks_all;
  int32_t b_lo = 0, b_hi = (int32_t)ba.size();
  const int32_t need_lo = std::max(0, -offsets.front() + 1);  // + 1: a window may start one column early for alignment
  while (b_lo < (int32_t)ba.size() && ba[b_lo].row0 < need_lo) ++b_lo;
  if (b_lo >= (int32_t)ba.size()) return;
  const int parity = ba[b_lo].row0 & 1;
  if (!dia_plan(offsets, dia->P, parity)) return;
  if (dia_smem_bytes(dia->P, 4) > 100 * 1024) return;  // at least two CTAs per SM
  dia_row_range(dia->P, A.n_rows, A.n_cols, dia->row_lo, dia->row_hi, parity);
  if (ba[b_lo].row0 < dia->row_lo) return;  // (cannot happen: need_lo covers the alignment slack)
  while (b_hi > b_lo && ba[b_hi - 1].row0 + (ba[b_hi - 1].nrows_kind & 0xffff) > dia->row_hi) --b_hi;
  if (b_hi - b_lo < 4) return;
  dia->b_lo = b_lo;
  dia->b_hi = b_hi;
  dia->row_lo = ba[b_lo].row0;
  dia->row_hi = ba[b_hi - 1].row0 + (ba[b_hi - 1].nrows_kind & 0xffff);
  // 2. diagonal-major values
  dia->npad = (((int64_t)A.n_rows + kDiaThreads - 1) / kDiaThreads + 1) * kDiaThreads;
  dia->vals.alloc((size_t)D * dia->npad * 8);
  JP_CUDA(cudaMemsetAsync(dia->vals.p, 0, (size_t)D * dia->npad * 8, st));
  DeviceBuffer d_off, d_missing;
  d_off.alloc((size_t)D * 4);
  d_missing.alloc(8);
  JP_CUDA(cudaMemcpyAsync(d_off.p, offsets.data(), (size_t)D * 4, cudaMemcpyHostToDevice, st));
  JP_CUDA(cudaMemsetAsync(d_missing.p, 0, 8, st));
  dia_fill_kernel<<<grid, 256, 0, st>>>(A.rowptr.as<int32_t>(), A.colind.as<int32_t>(), A.vals.as<double>(), A.n_rows, dia->P,
                                        d_off.as<int32_t>(), dia->npad, dia->vals.as<double>(), d_missing.as<unsigned long long>());
  JP_LAUNCH_CHECK();
  unsigned long long missing = 0;
  JP_CUDA(cudaMemcpyAsync(&missing, d_missing.p, 8, cudaMemcpyDeviceToHost, st));
  JP_CUDA(cudaStreamSynchronize(st));
  if (missing != 0) return;  // cannot happen (the offsets come from the same indices); never trust a partial copy
  A.dia = std::move(dia);
  A.dia_state = 1;
}

template <int DMAX>
static void launch_dia_impl(const DiaData& d, const double* x, int64_t ldx, double* y, int64_t ldy, int32_t k, double alpha, double beta,
                            cudaStream_t stream) {
  constexpr int KT = 4;
  const size_t smem = dia_smem_bytes(d.P, KT);
  set_smem_attr(spmm_dia_kernel<KT, DMAX>, smem);
  const int grid = (d.row_hi - d.row_lo + kDiaThreads - 1) / kDiaThreads;
  spmm_dia_kernel<KT, DMAX><<<grid, kDiaThreads, smem, stream>>>(d.vals.as<double>(), d.npad, x, ldx, y, ldy, k, d.row_lo, d.row_hi, alpha, beta,
                                                                  d.P);
  JP_LAUNCH_CHECK();
}

// y = beta*y + alpha*A*x for the WHOLE matrix (k > 1, no halo columns): the diagonal kernel on the rows whose windows lie
// inside X, the gather kernel (via `edge`) on the CSR row blocks of the first and last rows
void launch_spmm_dia(const Csr& A, const double* x, int64_t ldx, double* y, int64_t ldy, int32_t k, double alpha, double beta,
                     cudaStream_t stream) {
  const DiaData& d = *A.dia;
  if (d.P.D <= 8) launch_dia_impl<8>(d, x, ldx, y, ldy, k, alpha, beta, stream);
  else if (d.P.D <= 16) launch_dia_impl<16>(d, x, ldx, y, ldy, k, alpha, beta, stream);
  else if (d.P.D <= 28) launch_dia_impl<28>(d, x, ldx, y, ldy, k, alpha, beta, stream);
  else launch_dia_impl<32>(d, x, ldx, y, ldy, k, alpha, beta, stream);
  const char* blocks = static_cast<const char*>(A.blocks_all.p);
  const int32_t nb = A.n_blocks_all;
  if (d.b_lo > 0) launch_spmm_gather_blocks(A, blocks, d.b_lo, x, ldx, y, ldy, k, alpha, beta, stream);
  if (d.b_hi < nb) launch_spmm_gather_blocks(A, blocks + (size_t)d.b_hi * kBlockDescBytes, nb - d.b_hi, x, ldx, y, ldy, k, alpha, beta, stream);
}

int64_t csr_dia_info(const Csr& A, int what) {
  if (A.dia_state != 1) return 0;
  return what == 0 ? A.dia->P.D : what == 1 ? (A.dia->row_hi - A.dia->row_lo) : A.dia->P.W;
}

}  // namespace jp
