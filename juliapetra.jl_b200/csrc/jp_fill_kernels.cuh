// jp_fill_kernels.cuh -- device side of CSRMatrix fillComplete (src/CSRMatrix.jl:706-747) for a contiguous domain map:
//   __makeColMap                  src/CSRGraphInternalMethods.jl:852-982   column map = [used local GIDs in domain order |
//                                                                          remote GIDs grouped by owner, ascending GID]
//   makeIndicesLocal              src/CSRGraphInternalMethods.jl:636-712   GID -> LID of the column map
//   sortAndMergeIndicesAndValues  src/CSRMatrix.jl:549-606                 per row: stable sort by LID, duplicates summed
//                                                                          left to right in insertion order
//   fillLocalGraphAndMatrix       src/CSRMatrix.jl:403-484                 packed CSR (row pointers, LIDs, values)
// All of it is integer work whose results must be bit-identical to the host path / the oracle.
//
// Sort-free column map: one bit per GLOBAL column id (N/8 bytes of HBM: 16 MB for 512^3).  Marking the bitmap is one
// streaming pass over the entries; a popcount prefix over its words turns "how many referenced GIDs precede g" into two
// loads, and because a contiguous (linear) map gives ascending owner PIDs for ascending GIDs, rank order IS the
// reference's remote order (owner PID, then GID -- SURVEY.md 8c deviation 1).  With the domain's own GIDs occupying
// bits [lo, hi) and rank(b) = number of set bits below b:
//     local column  b in [lo, hi):  LID = rank(b) - rank(lo)
//     remote column b <  lo      :  LID = n_local_used + rank(b)
//     remote column b >= hi      :  LID = rank(b)                  (every used local bit precedes it)
// Per-row sort: rank sort (position = number of keys that are smaller, or equal and inserted earlier), which is stable
// by construction, needs no scratch beyond the output and is O(L^2/32) per warp -- rows are short (mean 7..27) and the
// power-law tail (L up to 10^4) goes to a whole thread block per row.
//
// This file holds device code only, in the CUDA subset that tests/cuda_emul/cuda_emul.h can execute on the CPU.
#pragma once
#include <cstdint>

namespace jp {
namespace fillk {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kScanItems = 8;                      // consecutive elements per thread in the scan kernels
constexpr int kScanChunk = kThreads * kScanItems;  // elements per block
constexpr int kWarpRowMax = 1024;                  // longer rows are sorted by a whole block

// ---- block-wide exclusive scan of one uint32 per thread --------------------------------------------------------
__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// returns the exclusive prefix of v over the block; *total = block sum (every thread).  s_warp has kWarps + 1 slots.
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* s_warp, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t incl = warp_inclusive_scan(v, lane);
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const uint32_t w = lane < kWarps ? s_warp[lane] : 0u;
    const uint32_t wi = warp_inclusive_scan(w, lane);
    if (lane < kWarps) s_warp[lane] = wi - w;
    if (lane == kWarps - 1) s_warp[kWarps] = wi;
  }
  __syncthreads();
  const uint32_t r = s_warp[warp] + incl - v;
  *total = s_warp[kWarps];
  __syncthreads();  // s_warp may be reused by the caller
  return r;
}

struct LoadU32 {
  const uint32_t* p;
  __device__ __forceinline__ uint32_t operator()(int64_t i) const { return p[i]; }
};
struct LoadPopc {
  const uint32_t* p;
  __device__ __forceinline__ uint32_t operator()(int64_t i) const { return (uint32_t)__popc(p[i]); }
};

// phase 1: one sum per block of kScanChunk elements
template <class Load>
__global__ void __launch_bounds__(kThreads) scan_block_sums_kernel(Load load, int64_t n, uint32_t* __restrict__ block_sums) {
  __shared__ uint32_t s_warp[kWarps + 1];
  const int64_t base = (int64_t)blockIdx.x * kScanChunk;
  uint32_t v = 0;
#pragma unroll
  for (int q = 0; q < kScanItems; ++q) {
    const int64_t i = base + (int64_t)q * kThreads + threadIdx.x;  // coalesced; the order does not matter for a sum
    if (i < n) v += load(i);
  }
  uint32_t total;
  block_exclusive_scan(v, s_warp, &total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// phase 2 (ONE block): block_sums -> exclusive offsets in place; the grand total goes to total_out[0] and, when given,
// to *last (the n-th slot of the element-wise output, i.e. rowptr[n_rows])
__global__ void __launch_bounds__(kThreads) scan_block_offsets_kernel(uint32_t* __restrict__ block_sums, int64_t nb,
                                                                        uint32_t* __restrict__ total_out, uint32_t* __restrict__ last) {
  __shared__ uint32_t s_warp[kWarps + 1];
  uint32_t carry = 0;
  for (int64_t base = 0; base < nb; base += kThreads) {  // block-uniform trip count
    const int64_t i = base + threadIdx.x;
    const uint32_t v = i < nb ? block_sums[i] : 0u;
    uint32_t total;
    const uint32_t ex = block_exclusive_scan(v, s_warp, &total);
    if (i < nb) block_sums[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) {
    total_out[0] = carry;
    if (last) *last = carry;
  }
}

// phase 3: element-wise exclusive prefix
template <class Load>
__global__ void __launch_bounds__(kThreads) scan_write_kernel(Load load, int64_t n, const uint32_t* __restrict__ block_offsets,
                                                                uint32_t* __restrict__ out) {
  __shared__ uint32_t s_warp[kWarps + 1];
  const int64_t base = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * kScanItems;
  uint32_t item[kScanItems];
  uint32_t v = 0;
#pragma unroll
  for (int q = 0; q < kScanItems; ++q) {
    const int64_t i = base + q;
    item[q] = i < n ? load(i) : 0u;
    v += item[q];
  }
  uint32_t total;
  uint32_t ex = block_exclusive_scan(v, s_warp, &total) + block_offsets[blockIdx.x];
#pragma unroll
  for (int q = 0; q < kScanItems; ++q) {
    const int64_t i = base + q;
    if (i < n) out[i] = ex;
    ex += item[q];
  }
}

// ---- column map --------------------------------------------------------------------------------------------------
// one bit per global column id referenced by an entry; GIDs outside [gmin, gmax] are counted in *bad
__global__ void __launch_bounds__(kThreads) fill_mark_kernel(const int64_t* __restrict__ gids, int64_t nnz, int64_t gmin, int64_t gmax,
                                                               uint32_t* bitmap, unsigned long long* bad) {
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < nnz; i += stride) {
    const int64_t g = gids[i];
    if (g < gmin || g > gmax) {
      atomicAdd(bad, 1ull);
      continue;
    }
    const int64_t b = g - gmin;
    const uint32_t m = 1u << (b & 31);
    uint32_t* w = bitmap + (b >> 5);
    if (!(*(volatile uint32_t*)w & m)) atomicOr(w, m);  // most columns are referenced several times: skip the atomic
  }
}

struct ColNumbering {
  const uint32_t* bitmap;  // G/32 + 1 words
  const uint32_t* prefix;  // exclusive popcount prefix of the words
  int64_t gmin;            // GID of bit 0
  int64_t lo, hi;          // bits of the domain map's own GIDs: [lo, hi)
  uint32_t before_lo;      // rank(lo)
  uint32_t n_local_used;   // rank(hi) - rank(lo)
};

__device__ __forceinline__ uint32_t bit_rank(const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ prefix, int64_t b) {
  const int64_t w = b >> 5;
  return prefix[w] + (uint32_t)__popc(bitmap[w] & ((1u << (b & 31)) - 1u));
}

__device__ __forceinline__ int32_t lid_of_rank(const ColNumbering& c, int64_t b, uint32_t r) {
  if (b >= c.lo && b < c.hi) return (int32_t)(r - c.before_lo);
  if (b < c.lo) return (int32_t)(c.n_local_used + r);
  return (int32_t)r;
}

// 0-based LID of a GID in the column map
__device__ __forceinline__ int32_t col_lid(const ColNumbering& c, int64_t gid) {
  const int64_t b = gid - c.gmin;
  return lid_of_rank(c, b, bit_rank(c.bitmap, c.prefix, b));
}

// ONE block.  out[0] = rank(lo), out[1] = used local columns, out[2] = all columns, out[3] = numSameIDs of
// Import(domainMap, colMap) = length of the run of set bits starting at lo (src/Import.jl:97-118)
__global__ void __launch_bounds__(kThreads) fill_counts_kernel(const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ prefix,
                                                                 int64_t G, int64_t lo, int64_t hi, uint32_t* __restrict__ out) {
  __shared__ unsigned long long s_first;  // first unset bit of [lo, hi), relative to lo
  if (threadIdx.x == 0) s_first = (unsigned long long)(hi - lo);
  __syncthreads();
  if (hi > lo) {
    const int64_t w0 = lo >> 5, w1 = (hi - 1) >> 5;
    for (int64_t w = w0 + threadIdx.x; w <= w1; w += kThreads) {
      uint32_t mask = 0xffffffffu;
      if (w == w0) mask &= ~((1u << (lo & 31)) - 1u);
      if (w == w1 && (hi & 31)) mask &= (1u << (hi & 31)) - 1u;
      const uint32_t zeros = ~bitmap[w] & mask;
      if (zeros) atomicMin(&s_first, (unsigned long long)((w << 5) + (__ffs((int)zeros) - 1) - lo));
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t rlo = bit_rank(bitmap, prefix, lo), rhi = bit_rank(bitmap, prefix, hi);
    out[0] = rlo;
    out[1] = rhi - rlo;
    out[2] = bit_rank(bitmap, prefix, G);
    out[3] = (uint32_t)s_first;
  }
}

// GIDs of the column map, in LID order
__global__ void __launch_bounds__(kThreads) fill_colmap_kernel(ColNumbering c, int64_t nwords, int64_t* __restrict__ colmap_gids) {
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t w = (int64_t)blockIdx.x * kThreads + threadIdx.x; w < nwords; w += stride) {
    uint32_t word = c.bitmap[w];
    uint32_t r = c.prefix[w];
    while (word) {
      const int bit = __ffs((int)word) - 1;
      word &= word - 1u;
      const int64_t b = (w << 5) + bit;
      colmap_gids[lid_of_rank(c, b, r)] = c.gmin + b;
      ++r;
    }
  }
}

// ---- per-row remap + stable sort + duplicate merge ---------------------------------------------------------------
struct RowJob {
  const int64_t* in_ptr;  // [n_rows + 1] 0-based offsets of the rows' entries (insertion order)
  const int64_t* gids;    // column GIDs
  const double* vals;
  int32_t* k0;            // scratch: LIDs in insertion order
  int32_t* s_lid;         // out: sorted + merged LIDs of row r at [in_ptr[r], in_ptr[r] + cnt[r])
  double* s_val;
  uint32_t* cnt;          // out: entries of row r after the merge
  int32_t* rowmax;        // out: largest LID of row r (-1 for an empty row)
  int32_t n_rows;
  ColNumbering c;
};

template <bool BLOCKWIDE>
__device__ __forceinline__ void group_sync() {
  if (BLOCKWIDE) __syncthreads();
  else __syncwarp();
}

// one row by one group of threads (a warp, or the whole block); tid = index in the group
template <bool BLOCKWIDE>
__device__ void sort_merge_row(const RowJob& j, int32_t r, int tid, int nthreads) {
  const int64_t s = j.in_ptr[r];
  const int64_t L = j.in_ptr[r + 1] - s;
  // 1. makeIndicesLocal
  for (int64_t i = tid; i < L; i += nthreads) j.k0[s + i] = col_lid(j.c, j.gids[s + i]);
  group_sync<BLOCKWIDE>();
  // 2. rank sort: position = #(smaller keys) + #(equal keys inserted earlier)  -> stable
  for (int64_t i = tid; i < L; i += nthreads) {
    const int32_t ki = j.k0[s + i];
    int64_t rank = 0;
    for (int64_t t = 0; t < L; ++t) {
      const int32_t kt = j.k0[s + t];
      rank += (kt < ki) || (kt == ki && t < i);
    }
    j.s_lid[s + rank] = ki;
    j.s_val[s + rank] = j.vals[s + i];
  }
  group_sync<BLOCKWIDE>();
  // 3. merge duplicates: every run of equal LIDs is summed left to right (insertion order) into its first entry
  //    (src/CSRMatrix.jl:588-597) and the row is compacted in place, 32 sorted positions at a time
  if (tid < 32) {
    uint32_t nout = 0;
    for (int64_t base = 0; base < L; base += 32) {  // warp-uniform trip count
      const int64_t p = base + tid;
      const bool in = p < L;
      const int32_t k = in ? j.s_lid[s + p] : 0;
      const bool head = in && (p == 0 || j.s_lid[s + p - 1] != k);
      double sum = 0.0;
      if (head) {
        sum = j.s_val[s + p];
        for (int64_t q = p + 1; q < L && j.s_lid[s + q] == k; ++q) sum = __dadd_rn(sum, j.s_val[s + q]);
      }
      const unsigned heads = __ballot_sync(0xffffffffu, head);
      __syncwarp();  // every read of this chunk (and of the run tails behind it) precedes the compaction writes
      if (head) {
        const uint32_t o = nout + (uint32_t)__popc(heads & ((1u << tid) - 1u));  // o <= p: never ahead of the reads
        j.s_lid[s + o] = k;
        j.s_val[s + o] = sum;
      }
      nout += (uint32_t)__popc(heads);
      __syncwarp();
    }
    if (tid == 0) {
      j.cnt[r] = nout;
      j.rowmax[r] = nout ? j.s_lid[s + nout - 1] : -1;
    }
  }
}

// rows of at most kWarpRowMax entries: one warp per row
__global__ void __launch_bounds__(kThreads) fill_rows_warp_kernel(RowJob j) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * kThreads) >> 5;
  for (int64_t r = warp; r < j.n_rows; r += nwarps) {
    if (j.in_ptr[r + 1] - j.in_ptr[r] > kWarpRowMax) continue;  // warp-uniform
    sort_merge_row<false>(j, (int32_t)r, lane, 32);
  }
}

// rows longer than kWarpRowMax: the block scans 256 rows at a time for them and sorts each with all its threads
__global__ void __launch_bounds__(kThreads) fill_rows_block_kernel(RowJob j) {
  __shared__ int s_n;
  __shared__ int32_t s_rows[kThreads];
  for (int64_t base = (int64_t)blockIdx.x * kThreads; base < j.n_rows; base += (int64_t)gridDim.x * kThreads) {  // block-uniform
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    const int64_t r = base + threadIdx.x;
    if (r < j.n_rows && j.in_ptr[r + 1] - j.in_ptr[r] > kWarpRowMax) s_rows[atomicAdd(&s_n, 1)] = (int32_t)r;
    __syncthreads();
    const int nl = s_n;
    for (int q = 0; q < nl; ++q) sort_merge_row<true>(j, s_rows[q], (int)threadIdx.x, kThreads);
    __syncthreads();  // s_n / s_rows are rewritten by the next batch
  }
}

// merged rows [in_ptr[r], +cnt) -> packed CSR at out_ptr[r]; only needed when duplicates were merged
__global__ void __launch_bounds__(kThreads) fill_pack_kernel(const int64_t* __restrict__ in_ptr, const uint32_t* __restrict__ out_ptr,
                                                               const int32_t* __restrict__ s_lid, const double* __restrict__ s_val,
                                                               int32_t n_rows, int32_t* __restrict__ colind, double* __restrict__ vals) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * kThreads) >> 5;
  for (int64_t r = warp; r < n_rows; r += nwarps) {
    const int64_t s = in_ptr[r];
    const uint32_t o = out_ptr[r], n = out_ptr[r + 1] - o;
    for (uint32_t i = lane; i < n; i += 32) {
      colind[o + i] = s_lid[s + i];
      vals[o + i] = s_val[s + i];
    }
  }
}

}  // namespace fillk
}  // namespace jp
