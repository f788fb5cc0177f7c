// jp_dia_plan.h -- host-side planning of the diagonal SpMM kernel (spmm_dia_kernel, jp_csr_kernels.cuh): pure C++, shared
// by jp_dia.cu and by the CPU execution model of the kernels in tests/cuda_emul.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

#include "jp_csr_kernels.cuh"

namespace jp {
namespace csrk {

// From the ascending diagonal offsets: the X windows a block of kDiaThreads rows needs (relative to its first row),
// merged where they overlap, and every diagonal's tile position.  The bulk copies need 16-byte aligned sources: with the
// first row of every block congruent to `parity` (mod 2), a window starts at a relative column of the same parity and
// has an even length.  Returns false when there are too many diagonals / windows for the kernel's parameter block.
inline bool dia_plan(const std::vector<int32_t>& offsets, DiaParams& P, int parity = 0) {
  const int D = (int)offsets.size();
  if (D == 0 || D > kDiaMaxDiags) return false;
  P = DiaParams();
  P.D = D;
  struct Win { int32_t lo, hi; };  // columns [lo, hi) relative to the block's first row
  std::vector<Win> wins;
  for (int d = 0; d < D; ++d) {
    int32_t lo = offsets[d], hi = offsets[d] + kDiaThreads;
    lo -= (lo - parity) & 1;   // floor to the blocks' parity (two's complement: also right for negative values)
    hi += (hi - lo) & 1;       // even length
    if (!wins.empty() && lo <= wins.back().hi) wins.back().hi = std::max(wins.back().hi, hi);
    else wins.push_back({lo, hi});
  }
  if ((int)wins.size() > kDiaMaxSegs) return false;
  P.nseg = (int32_t)wins.size();
  int32_t pos = 0;
  for (int s = 0; s < P.nseg; ++s) {
    P.seg_rel[s] = wins[s].lo;
    P.seg_len[s] = wins[s].hi - wins[s].lo;
    P.seg_pos[s] = pos;
    pos += P.seg_len[s];
  }
  if (pos > kDiaTileStride) return false;  // the tile rows have a fixed stride (jp_csr_kernels.cuh)
  P.W = pos;
  P.seg_cols = pos;
  for (int d = 0; d < D; ++d) {
    int s = 0;
    while (!(offsets[d] >= wins[s].lo && offsets[d] + kDiaThreads <= wins[s].hi)) ++s;
    P.pos[d] = P.seg_pos[s] + (offsets[d] - wins[s].lo);
  }
  return true;
}

inline size_t dia_smem_bytes(const DiaParams& P, int KT) { return ((size_t)(KT - 1) * kDiaTileStride + P.W) * 8; }

// rows [row_lo, row_hi) whose windows -- for EVERY block of kDiaThreads rows starting at row_lo, the last, partial one
// included -- stay inside [0, n_cols); row_lo has the parity the plan was made for.  Empty (row_lo >= row_hi) when the
// matrix is too small.
inline void dia_row_range(const DiaParams& P, int32_t n_rows, int32_t n_cols, int32_t& row_lo, int32_t& row_hi, int parity = 0) {
  const int32_t min_rel = P.seg_rel[0], max_end = P.seg_rel[P.nseg - 1] + P.seg_len[P.nseg - 1];  // window extremes relative to a block's first row
  row_lo = std::max(0, -min_rel);
  row_lo += (row_lo - parity) & 1;
  // a block starting at b reads columns [b + min_rel, b + max_end); the last block starts below row_hi
  const int64_t hi = (int64_t)n_cols - max_end + 1;  // exclusive bound on block starts
  row_hi = (int32_t)std::max<int64_t>(0, std::min<int64_t>(n_rows, hi));
}

}  // namespace csrk
}  // namespace jp
