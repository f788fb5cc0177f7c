// jp_csr_blocks.h -- host-side construction of the row blocks the SpMV/SpMM kernels consume (pure C++: shared by
// jp_csr.cu and by the CPU execution model of the kernels in tests/cuda_emul).
#pragma once
#include <algorithm>
#include <cstdint>
#include <thread>
#include <vector>

#include "jp_csr_kernels.cuh"

namespace jp {
namespace csrk {

inline size_t stage_smem_bytes(int cap) { return (size_t)(cap + 8) * 12 + (size_t)(kMaxRowsPerBlock + 8) * 4; }

// cut rows into blocks; cls[r] = 1 when the row touches a remote column
inline void build_blocks(const std::vector<int32_t>& rp, const std::vector<uint8_t>& cls, int32_t cap, int32_t max_rows, bool allow_direct,
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

// SpMM tile tables: per block the sorted unique columns it references, per nonzero the position of its column in that
// list.  Returns false (tables unusable) when the matrix has a LONGROW block, when columns are hardly shared between
// the rows of a block, or when two stages do not fit the shared memory.
struct TileTables {
  std::vector<int32_t> ucols;
  std::vector<uint16_t> lidx;
  int32_t umax = 0;
};

inline bool build_tile_tables(const std::vector<int32_t>& ci, std::vector<BlockDesc>& all, int32_t nnz_cap, int KT, TileTables& out) {
  const size_t nb = all.size();
  const size_t nnz = ci.size();
  std::vector<int32_t> ubuf(nnz);  // block b's unique list is built in place at [p0, p0 + ucount)
  out.lidx.assign(nnz + 64, 0);
  const int nthreads = (int)std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
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
      for (int32_t i = b.p0; i < b.p1; ++i) out.lidx[i] = (uint16_t)(std::lower_bound(u, u + uc, ci[i]) - u);
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
  if (has_longrow || sum_nnz == 0 || (double)sum_u > 0.7 * (double)sum_nnz || 2 * pipe_stage_bytes(nnz_cap, upad, KT) > 220 * 1024 ||
      off > INT32_MAX - 64)
    return false;
  out.ucols.assign((size_t)off + 64, 0);
  for (const BlockDesc& b : all) std::copy(ubuf.begin() + b.p0, ubuf.begin() + b.p0 + b.ucount, out.ucols.begin() + b.ucol0);
  out.umax = umax;
  return true;
}

}  // namespace csrk
}  // namespace jp
