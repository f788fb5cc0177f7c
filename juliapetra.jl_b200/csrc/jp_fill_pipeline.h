// jp_fill_pipeline.h -- host orchestration of the device fillComplete (kernels: jp_fill_kernels.cuh).
// Compiled by nvcc into libjpetra_b200.so (included from jp_fill.cu after jp_common.h) and, unchanged, by g++ against
// tests/cuda_emul/emul_runtime.h so the same allocation sizes, launch geometry and scans are exercised on the CPU.
//
// Steps (all stream-ordered on the compute stream; two 32-byte downloads synchronise):
//   1. fill_mark_kernel          one bit per referenced global column id
//   2. scan (popcount)           exclusive popcount prefix of the bitmap words
//   3. fill_counts_kernel        rank(lo), used local columns, all columns, numSameIDs          -> host (32 B)
//   4. fill_colmap_kernel        GIDs of the column map in LID order
//   5. fill_rows_*_kernel        per row: GID -> LID, stable rank sort, duplicate merge, count, largest LID
//   6. scan (counts)             row pointers of the packed CSR; total                            -> host (4 B)
//   7. fill_pack_kernel          only when duplicates were merged; otherwise the sorted arrays ARE the CSR
#pragma once
#include <vector>

#include "jp_fill_kernels.cuh"

namespace jp {

struct FillOutput {
  int32_t n_rows = 0, n_cols = 0, n_local_used = 0, n_same = 0;
  int64_t nnz = 0;
  DeviceBuffer rowptr;          // int32[n_rows + 1 + 16], 0-based
  DeviceBuffer colind, vals;    // nnz + 64 elements (zeroed slack, like jp_csr_create)
  DeviceBuffer colmap;          // int64[n_cols]: GIDs of the column map
  std::vector<int32_t> rp_host, rowmax_host;
};

inline unsigned fill_grid(int64_t work_items, int per_block) {
  int64_t want = (work_items + per_block - 1) / per_block;
  const int64_t cap = (int64_t)ctx().sm_count * 16;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (unsigned)want;
}

// out[i] = sum_{j<i} load(j) for i < n; the total goes to total_dev[0] and, when `last` is given, to *last
template <class Load>
void exclusive_scan_u32(Load load, int64_t n, uint32_t* out, uint32_t* total_dev, uint32_t* last, DeviceBuffer& tmp, cudaStream_t st) {
  using namespace fillk;
  const int64_t nb = (n + kScanChunk - 1) / kScanChunk;
  tmp.ensure((size_t)(nb + 1) * 4);
  if (nb > 0) JP_KLAUNCH(scan_block_sums_kernel<Load>, (unsigned)nb, kThreads, 0, st, load, n, tmp.as<uint32_t>());
  JP_KLAUNCH(scan_block_offsets_kernel, 1u, kThreads, 0, st, tmp.as<uint32_t>(), nb, total_dev, last);
  if (nb > 0) JP_KLAUNCH(scan_write_kernel<Load>, (unsigned)nb, kThreads, 0, st, load, n, tmp.as<uint32_t>(), out);
}

// h_ptr: n_rows + 1 offsets (0-based) of every row's entries in insertion order; h_gids / h_vals: the entries.
// [gmin, gmax] = minAllGID..maxAllGID of the domain map; the domain map owns dom_n contiguous GIDs from dom_min.
inline void fill_complete_device(int32_t n_rows, const int64_t* h_ptr, const int64_t* h_gids, const double* h_vals, int64_t gmin,
                                 int64_t gmax, int64_t dom_min, int32_t dom_n, FillOutput& out) {
  using namespace fillk;
  cudaStream_t st = ctx().compute;
  JP_REQUIRE(n_rows >= 0 && dom_n >= 0, "fillComplete: negative size");
  JP_REQUIRE(h_ptr != nullptr && h_ptr[0] == 0, "fillComplete: entry offsets must start at 0");
  int64_t maxlen = 0;
  for (int32_t r = 0; r < n_rows; ++r) {
    JP_REQUIRE(h_ptr[r + 1] >= h_ptr[r], "fillComplete: entry offsets must be non-decreasing");
    maxlen = std::max(maxlen, h_ptr[r + 1] - h_ptr[r]);
  }
  const int64_t nnz_in = h_ptr[n_rows];
  JP_REQUIRE(nnz_in < ((int64_t)1 << 31) - 64, "fillComplete: local entries must fit int32 (LID) row pointers");
  JP_REQUIRE(nnz_in == 0 || (h_gids && h_vals), "fillComplete: null array");
  JP_REQUIRE(gmax >= gmin - 1, "fillComplete: bad global id range");
  const int64_t G = gmax - gmin + 1;
  JP_REQUIRE(dom_n == 0 || (dom_min >= gmin && dom_min + dom_n - 1 <= gmax), "fillComplete: the domain map's ids lie outside the global range");
  const int64_t lo = dom_n ? dom_min - gmin : 0, hi = lo + dom_n;
  const int64_t nwords = (G >> 5) + 1;  // one spare word so that rank(G) is a plain lookup

  out.n_rows = n_rows;
  DeviceBuffer d_ptr, d_gids, d_vals, bitmap, prefix, small, scan_tmp, k0, s_lid, s_val, cnt, rowmax;
  d_ptr.alloc(((size_t)n_rows + 1) * 8);
  d_gids.alloc((size_t)nnz_in * 8);
  d_vals.alloc((size_t)nnz_in * 8);
  JP_CUDA(cudaMemcpyAsync(d_ptr.p, h_ptr, ((size_t)n_rows + 1) * 8, cudaMemcpyHostToDevice, st));
  if (nnz_in) {
    JP_CUDA(cudaMemcpyAsync(d_gids.p, h_gids, (size_t)nnz_in * 8, cudaMemcpyHostToDevice, st));
    JP_CUDA(cudaMemcpyAsync(d_vals.p, h_vals, (size_t)nnz_in * 8, cudaMemcpyHostToDevice, st));
  }
  bitmap.alloc((size_t)nwords * 4);
  prefix.alloc((size_t)nwords * 4);
  small.alloc(64);  // [0..3] counts, [4] scan total, byte 32: bad-GID counter (uint64)
  JP_CUDA(cudaMemsetAsync(bitmap.p, 0, (size_t)nwords * 4, st));
  JP_CUDA(cudaMemsetAsync(small.p, 0, 64, st));
  uint32_t* d_small = small.as<uint32_t>();
  unsigned long long* d_bad = reinterpret_cast<unsigned long long*>(small.as<unsigned char>() + 32);

  // 1-3: column map numbering
  if (nnz_in) JP_KLAUNCH(fill_mark_kernel, fill_grid(nnz_in, kThreads * 4), kThreads, 0, st, d_gids.as<int64_t>(), nnz_in, gmin, gmax,
                         bitmap.as<uint32_t>(), d_bad);
  exclusive_scan_u32(LoadPopc{bitmap.as<uint32_t>()}, nwords, prefix.as<uint32_t>(), d_small + 4, nullptr, scan_tmp, st);
  JP_KLAUNCH(fill_counts_kernel, 1u, kThreads, 0, st, bitmap.as<uint32_t>(), prefix.as<uint32_t>(), G, lo, hi, d_small);
  uint32_t h_small[16];
  JP_CUDA(cudaMemcpyAsync(h_small, small.p, 64, cudaMemcpyDeviceToHost, st));
  JP_CUDA(cudaStreamSynchronize(st));
  unsigned long long bad = 0;
  std::memcpy(&bad, (const unsigned char*)h_small + 32, 8);
  if (bad != 0)
    fail(JP_INVALID_ARGUMENT, "fillComplete: " + std::to_string(bad) + " column indices lie outside the domain map's global id range");
  JP_REQUIRE(h_small[2] < 0x7fffffffu, "fillComplete: the column map does not fit int32 LIDs");
  ColNumbering c;
  c.bitmap = bitmap.as<uint32_t>();
  c.prefix = prefix.as<uint32_t>();
  c.gmin = gmin;
  c.lo = lo;
  c.hi = hi;
  c.before_lo = h_small[0];
  c.n_local_used = h_small[1];
  out.n_local_used = (int32_t)h_small[1];
  out.n_cols = (int32_t)h_small[2];
  out.n_same = (int32_t)h_small[3];

  // 4: the column map's GIDs
  out.colmap.alloc((size_t)out.n_cols * 8);
  if (out.n_cols) JP_KLAUNCH(fill_colmap_kernel, fill_grid(nwords, kThreads), kThreads, 0, st, c, nwords, out.colmap.as<int64_t>());

  // 5: rows
  k0.alloc((size_t)nnz_in * 4);
  s_lid.alloc(((size_t)nnz_in + 64) * 4);
  s_val.alloc(((size_t)nnz_in + 64) * 8);
  cnt.alloc(((size_t)n_rows + 1) * 4);
  rowmax.alloc((size_t)n_rows * 4);
  RowJob j;
  j.in_ptr = d_ptr.as<int64_t>();
  j.gids = d_gids.as<int64_t>();
  j.vals = d_vals.as<double>();
  j.k0 = k0.as<int32_t>();
  j.s_lid = s_lid.as<int32_t>();
  j.s_val = s_val.as<double>();
  j.cnt = cnt.as<uint32_t>();
  j.rowmax = rowmax.as<int32_t>();
  j.n_rows = n_rows;
  j.c = c;
  if (n_rows) {
    JP_KLAUNCH(fill_rows_warp_kernel, fill_grid(n_rows, kWarps), kThreads, 0, st, j);
    if (maxlen > kWarpRowMax) JP_KLAUNCH(fill_rows_block_kernel, fill_grid(n_rows, kThreads), kThreads, 0, st, j);
  }

  // 6: row pointers
  out.rowptr.alloc(((size_t)n_rows + 1 + 16) * 4);
  JP_CUDA(cudaMemsetAsync(out.rowptr.p, 0, ((size_t)n_rows + 1 + 16) * 4, st));
  exclusive_scan_u32(LoadU32{cnt.as<uint32_t>()}, n_rows, out.rowptr.as<uint32_t>(), d_small + 4, out.rowptr.as<uint32_t>() + n_rows,
                     scan_tmp, st);
  uint32_t total = 0;
  JP_CUDA(cudaMemcpyAsync(&total, d_small + 4, 4, cudaMemcpyDeviceToHost, st));
  JP_CUDA(cudaStreamSynchronize(st));
  out.nnz = total;

  // 7: packed CSR
  if (out.nnz == nnz_in) {
    out.colind.swap(s_lid);  // no duplicates anywhere: every row already sits at its final offset
    out.vals.swap(s_val);
  } else {
    out.colind.alloc(((size_t)out.nnz + 64) * 4);
    out.vals.alloc(((size_t)out.nnz + 64) * 8);
    JP_KLAUNCH(fill_pack_kernel, fill_grid(n_rows, kWarps), kThreads, 0, st, d_ptr.as<int64_t>(), out.rowptr.as<uint32_t>(),
               s_lid.as<int32_t>(), s_val.as<double>(), n_rows, out.colind.as<int32_t>(), out.vals.as<double>());
  }
  JP_CUDA(cudaMemsetAsync(out.colind.as<int32_t>() + out.nnz, 0, 64 * 4, st));
  JP_CUDA(cudaMemsetAsync(out.vals.as<double>() + out.nnz, 0, 64 * 8, st));

  // what the row-block builder needs on the host
  out.rp_host.resize((size_t)n_rows + 1);
  out.rowmax_host.resize((size_t)n_rows);
  JP_CUDA(cudaMemcpyAsync(out.rp_host.data(), out.rowptr.p, ((size_t)n_rows + 1) * 4, cudaMemcpyDeviceToHost, st));
  if (n_rows) JP_CUDA(cudaMemcpyAsync(out.rowmax_host.data(), rowmax.p, (size_t)n_rows * 4, cudaMemcpyDeviceToHost, st));
  JP_CUDA(cudaStreamSynchronize(st));  // also: the scratch buffers above are released when this function returns
}

}  // namespace jp
