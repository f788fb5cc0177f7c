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
// Which blocks go thread-per-row straight from the stage (KIND_DIRECT): longest row <= direct_maxlen and
// longest row <= direct_ratio * mean row length of the block.  32 and 2 are what round 1 measured on stencils; both are
// tuning knobs (JP_SPMV_DIRECT_MAXLEN, JP_SPMV_DIRECT_RATIO).
struct DirectRule {
  int32_t maxlen = 32;
  double ratio = 2.0;
  // every other block is KIND_SUBWARP with T = the largest power of two <= mean row length of the block / subwarp_div
  // lanes per row (2 <= T <= 32); 1.5 = 8 lanes at mean length 16, measured best of 4 / 8 / 16 on the power-law matrix
  double subwarp_div = 1.5;
};

inline void build_blocks(const std::vector<int32_t>& rp, const std::vector<uint8_t>& cls, int32_t cap, int32_t max_rows, bool allow_direct,
                  std::vector<BlockDesc>& interior, std::vector<BlockDesc>& boundary, std::vector<BlockDesc>& all, const DirectRule& rule = DirectRule()) {
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
      if (allow_direct && maxlen <= rule.maxlen && (double)maxlen * (double)nrows <= rule.ratio * (double)nnzb + 32.0) kind = KIND_DIRECT;
      else {
        const double target = (double)nnzb / (double)std::max<int64_t>(nrows, 1) / rule.subwarp_div;
        int lg = 1;
        while (lg < 5 && (double)(2 << lg) <= target) ++lg;
        kind = KIND_SUBWARP | (lg << 4);  // shifted by 16 below: the lane count lands in bits 20..23
      }
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

// SpMM "runs" path (spmm_runs_kernel): per block the columns it references, merged into 16-byte ALIGNED RUNS
// (start even, length even) so that the block's X tile can be copied global -> shared with 16-byte cp.async, and per
// nonzero the uint16 position of its column in the concatenated runs.  Returns false when the matrix is not banded
// enough for this to pay (tile wider than 0.7 nnz on average), when a block's tile does not fit, or with LONGROW blocks.
struct RunTables {
  std::vector<int32_t> runs;   // (start column, length) pairs, block b's at [2*ucol0, 2*(ucol0 + ucount))
  std::vector<uint16_t> lidx;  // per nonzero
  int32_t wmax = 0;            // widest tile (columns)
};

inline bool build_run_tables(const std::vector<int32_t>& ci, std::vector<BlockDesc>& all, int32_t max_tile_cols, RunTables& out) {
  const size_t nb = all.size();
  const size_t nnz = ci.size();
  out.lidx.assign(nnz + 64, 0);
  std::vector<std::vector<int32_t>> block_runs(nb);
  std::vector<int32_t> width(nb, 0), ncopied(nb, 0);
  const int nthreads = (int)std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  for (const BlockDesc& b : all)
    if (((b.nrows_kind >> 16) & 0xf) == KIND_LONGROW) return false;
  // cheap early answer for matrices that are not banded (a random 10^9-nonzero matrix must not pay a full sort of
  // every block just to be turned away): distinct aligned column pairs of up to 256 evenly spaced blocks
  {
    int64_t s_w = 0, s_nnz = 0;
    std::vector<int32_t> u;
    const size_t step = std::max<size_t>(1, nb / 256);
    for (size_t bi = 0; bi < nb; bi += step) {
      const BlockDesc& b = all[bi];
      u.assign(ci.begin() + b.p0, ci.begin() + b.p1);
      for (int32_t& c : u) c >>= 1;
      std::sort(u.begin(), u.end());
      s_w += 2 * (int64_t)(std::unique(u.begin(), u.end()) - u.begin());
      s_nnz += b.p1 - b.p0;
    }
    if (s_nnz == 0 || (double)s_w > 0.7 * (double)s_nnz) return false;
  }
  bool reject = false;  // (benign race: only ever set to true)
  auto worker = [&](int tix) {
    std::vector<int32_t> u, base;
    for (size_t bi = tix; bi < nb; bi += nthreads) {
      const BlockDesc& b = all[bi];
      if (reject) return;
      u.assign(ci.begin() + b.p0, ci.begin() + b.p1);
      std::sort(u.begin(), u.end());
      u.erase(std::unique(u.begin(), u.end()), u.end());
      std::vector<int32_t>& runs = block_runs[bi];
      int32_t cur_end = -1;
      for (int32_t c : u) {
        if (c < cur_end) continue;  // covered by the aligned tail of the current run
        if (!runs.empty() && (c & ~1) == cur_end) {
          runs.back() += 2;  // extend
        } else {
          runs.push_back(c & ~1);
          runs.push_back(2);
        }
        cur_end = (c | 1) + 1;
      }
      const size_t nr = runs.size() / 2;
      base.assign(nr, 0);
      int32_t w = 0, copied = 0;
      for (size_t q = 0; q < nr; ++q) {  // tile position of every run (padded strides: run_stride)
        base[q] = w;
        w += run_stride(runs[2 * q + 1]);
        copied += runs[2 * q + 1];
      }
      width[bi] = w;
      ncopied[bi] = copied;
      if (w > max_tile_cols || w > 65534) {
        reject = true;
        continue;
      }
      for (int32_t i = b.p0; i < b.p1; ++i) {
        const int32_t c = ci[i];
        size_t lo = 0, hi = nr;  // last run with start <= c
        while (hi - lo > 1) {
          const size_t mid = (lo + hi) / 2;
          if (runs[2 * mid] <= c) lo = mid;
          else hi = mid;
        }
        out.lidx[i] = (uint16_t)(base[lo] + (c - runs[2 * lo]));
      }
    }
  };
  {
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(worker, t);
    worker(0);
    for (auto& t : th) t.join();
  }
  if (reject) return false;
  int64_t sum_w = 0, sum_nnz = 0, off = 0;
  int32_t wmax = 0;
  for (size_t bi = 0; bi < nb; ++bi) {
    BlockDesc& b = all[bi];
    b.ucol0 = (int32_t)off;
    b.ucount = (int32_t)(block_runs[bi].size() / 2);
    b.pad0 = width[bi];    // tile columns incl. the padding between runs
    b.pad1 = ncopied[bi];  // columns actually copied (sum of the run lengths)
    off += b.ucount;
    wmax = std::max(wmax, width[bi]);
    sum_w += ncopied[bi];
    sum_nnz += b.p1 - b.p0;
  }
  if (sum_nnz == 0 || (double)sum_w > 0.7 * (double)sum_nnz || off > INT32_MAX / 2 - 64) return false;
  out.runs.assign((size_t)off * 2 + 64, 0);
  for (size_t bi = 0; bi < nb; ++bi) std::copy(block_runs[bi].begin(), block_runs[bi].end(), out.runs.begin() + 2 * (size_t)all[bi].ucol0);
  out.wmax = wmax;
  return true;
}

}  // namespace csrk
}  // namespace jp
