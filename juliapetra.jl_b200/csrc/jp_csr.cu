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
#include "jp_csr_blocks.h"
#include "jp_csr_kernels.cuh"
#include "jp_p2p.cuh"

namespace jp {

namespace {
using namespace csrk;

// ---------------------------------------------------------------------------------------------------------------
// apply! with the halo exchange INSIDE the SpMV kernel (k = 1, peer-memory halo): one launch per apply, no stream
// events, no NCCL call.  The grid is [pack CTAs | interior row blocks | boundary row blocks]; the protocol (two halo
// buffers selected by epoch parity, ready / ack flags) is described in jp_p2p.cuh.
//   pack ITEMS     gather the exported rows of x and store them straight into the neighbours' halo buffers over
//                  NVLink (IPC-mapped peer pointers); whoever completes the last item publishes ready = epoch with a
//                  system-scope release.  Items are CLAIMED from a ticket counter -- by the dedicated pack CTAs at the
//                  head of the grid and by every boundary CTA before it starts to wait;
//   interior CTAs  the ordinary row-block SpMV on local columns (the bulk of the grid; they hide the exchange);
//   boundary CTAs  sit behind the first three quarters of the interior blocks: claim pack items until none is left,
//                  acquire the neighbours' ready flags (set ~20 us earlier by the peers' pack items), run the row-block
//                  SpMV on x + halo, and the last one publishes ack = epoch so the senders may reuse this halo buffer two
//                  applies later.  The last quarter of the interior blocks runs behind them, so the claim / acquire
//                  latency of the boundary CTAs is not the tail of the kernel (at 8 GPUs they were 43 % of the last wave).
// Forward progress does not depend on the order in which the hardware dispatches CTAs: a CTA only ever spins on a
// remote flag AFTER every pack item of its own kernel has been claimed by a running CTA, pack items wait only for
// acks of apply e - 2 (published by kernels that precede the neighbours' current one in stream order), and interior
// CTAs never wait.  So on every rank the pack items of apply e complete, hence every boundary wait of apply e ends.
// ---------------------------------------------------------------------------------------------------------------
struct FusedSpmvArgs {
  const BlockDesc* interior;
  const BlockDesc* boundary;
  int32_t n_pack, n_interior, n_boundary;
  int32_t n_first;  // interior blocks in front of the boundary blocks; grid = [pack | first | boundary | rest of the interior]
  int32_t n_items;  // pack items (>= 1; item t packs export entries t*T + tid, strided by n_items*T)
  const int32_t* rowptr;
  const int32_t* colind;
  const double* vals;
  const double* x;
  const double* halo;  // this apply's halo buffer (parity already applied)
  int32_t nloc;
  double* y;
  int32_t cap;
  double alpha, beta;
  FusedHalo h;
};

// one pack item; called by every thread of the CTA
__device__ __forceinline__ void fused_pack_item(const FusedSpmvArgs& a, unsigned item, bool* s_last) {
  const int tid = threadIdx.x;
  const uint64_t epoch = a.h.epoch;
  if (tid < a.h.n_sends && epoch > 2)  // back-pressure: the receiver has consumed what this buffer held two applies ago
    wait_flag_ge(a.h.sends[tid].ack_flag, epoch - 2, a.h.timeout_flag);
  __syncthreads();
  const int64_t stride = (int64_t)a.n_items * kSpmvThreads;
  for (int64_t i = (int64_t)item * kSpmvThreads + tid; i < a.h.n_export; i += stride) {
    int sb = 0;
    while (sb + 1 < a.h.n_sends && i >= a.h.sends[sb].start + a.h.sends[sb].count) ++sb;
    const P2PSend& s = a.h.sends[sb];
    s.dst_base[(epoch & 1) * s.dst_buf_elems + s.dst_off + (i - s.start)] = __ldg(a.x + __ldg(a.h.lids + i));
  }
  __threadfence_system();
  __syncthreads();
  if (tid == 0) *s_last = (atomicAdd(a.h.tickets + 0, 1u) == (unsigned)a.n_items - 1);
  __syncthreads();
  if (*s_last) {
    __threadfence_system();
    if (tid < a.h.n_sends) st_release_sys(a.h.sends[tid].ready_flag, epoch);
    // a rank whose matrix has no boundary rows still owes its sources the ack
    if (a.n_boundary == 0 && tid < a.h.n_sources) st_release_sys(a.h.sources[tid].ack_flag, epoch);
    if (tid == 0) a.h.tickets[0] = 0;  // re-arm for the next apply (stream-ordered after this kernel)
  }
  __syncthreads();  // s_last is reused by the next item
}

// claim and run pack items until none is left.  n_pack + n_boundary CTAs call this once per kernel and each makes
// exactly one unsuccessful claim, so the CTA that draws the last ticket of the kernel re-arms the counter.
__device__ __forceinline__ void fused_pack_claims(const FusedSpmvArgs& a, unsigned* s_claim, bool* s_last) {
  const unsigned last_ticket = (unsigned)a.n_items + (unsigned)a.n_pack + (unsigned)a.n_boundary - 1u;
  for (;;) {
    if (threadIdx.x == 0) {
      const unsigned t = atomicAdd(a.h.tickets + 2, 1u);
      if (t == last_ticket) a.h.tickets[2] = 0;
      *s_claim = t;
    }
    __syncthreads();
    const unsigned t = *s_claim;
    __syncthreads();
    if (t >= (unsigned)a.n_items) return;
    fused_pack_item(a, t, s_last);
  }
}

// DOT: every CTA also leaves sum_r w[r] * y_new[r] over its rows in partials[blockIdx.x] (0 for the pack CTAs), the
// multi-GPU form of spmv_dot_kernel (jp_apply_dot with a peer-memory halo)
template <bool DOT>
__device__ __forceinline__ void fused_halo_body(const FusedSpmvArgs& a, const double* __restrict__ w, double* __restrict__ partials) {
  JP_DYN_SMEM(smem_raw);
  __shared__ double s_red[kSpmvThreads / 32];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ bool s_last;
  __shared__ unsigned s_claim;
  const int tid = threadIdx.x;
  const int bid = blockIdx.x;
  double dacc = 0.0;
  if (bid < a.n_pack) {
    fused_pack_claims(a, &s_claim, &s_last);
  } else if (bid < a.n_pack + a.n_first || bid >= a.n_pack + a.n_first + a.n_boundary) {
    const StageView sv = stage_view(smem_raw, a.cap);
    const BlockRange b = decode_block(a.interior, bid < a.n_pack + a.n_first ? bid - a.n_pack : bid - a.n_pack - a.n_boundary);
    if (DOT) dacc = spmv_block<false, 1, true>(b, sv, &s_bar, s_red, a.rowptr, a.colind, a.vals, a.x, nullptr, a.nloc, a.y, a.alpha, a.beta, w);
    else spmv_block<false, 1>(b, sv, &s_bar, s_red, a.rowptr, a.colind, a.vals, a.x, nullptr, a.nloc, a.y, a.alpha, a.beta);
  } else {
    // ---- boundary rows: no pack item may be left unclaimed before this CTA starts to wait for the neighbours ----
    fused_pack_claims(a, &s_claim, &s_last);
    if (tid < a.h.n_sources) wait_flag_ge(a.h.sources[tid].ready_flag, a.h.epoch, a.h.timeout_flag);
    __syncthreads();
    const StageView sv = stage_view(smem_raw, a.cap);
    const BlockRange b = decode_block(a.boundary, bid - a.n_pack - a.n_first);
    if (DOT) dacc = spmv_block<true, 1, true>(b, sv, &s_bar, s_red, a.rowptr, a.colind, a.vals, a.x, a.halo, a.nloc, a.y, a.alpha, a.beta, w);
    else spmv_block<true, 1>(b, sv, &s_bar, s_red, a.rowptr, a.colind, a.vals, a.x, a.halo, a.nloc, a.y, a.alpha, a.beta);
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(a.h.tickets + 1, 1u) == (unsigned)a.n_boundary - 1);
    __syncthreads();
    if (s_last) {
      if (tid < a.h.n_sources) st_release_sys(a.h.sources[tid].ack_flag, a.h.epoch);
      if (tid == 0) a.h.tickets[1] = 0;
    }
  }
  if (DOT) warp_partial_store(dacc, partials, blockIdx.x);
}

__global__ void __launch_bounds__(kSpmvThreads) spmv_fused_halo_kernel(const FusedSpmvArgs a) { fused_halo_body<false>(a, nullptr, nullptr); }

__global__ void __launch_bounds__(kSpmvThreads) spmv_fused_halo_dot_kernel(const FusedSpmvArgs a, const double* __restrict__ w,
                                                                             double* __restrict__ partials) {
  fused_halo_body<true>(a, w, partials);
}

int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return (v && *v) ? std::atoi(v) : dflt;
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


// Build the tables of the SpMM runs path (lazily, at the first k > 1 apply); a one-time host cost.  One configuration:
// 128 rows per block, one lane per row, 4 vectors per tile -- the best of the twelve measured in round 1
// (profiles/r1_spmm_ab_runs_tma_27pt128.json) and of the double-buffered variant measured in round 2
// (profiles/r2_spmm192_variants.json); the others are gone.
constexpr int kRunsCap = 3456, kRunsKT = 4, kRunsThreads = 128;
void build_runs(Csr& A) {
  if (A.nnz == 0 || A.n_rows == 0 || env_int("JP_SPMM_RUNS", 1) == 0) {  // on by default; only banded matrices pass build_run_tables
    A.runs_state = -1;
    return;
  }
  const int cap = kRunsCap, kt = kRunsKT, max_rows = kRunsMaxRows;
  if (A.max_row_len > cap) {  // a row longer than the stage would be a LONGROW block: not this kernel's matrix (no download needed)
    A.runs_state = -1;
    return;
  }
  std::vector<int32_t> ci((size_t)A.nnz), rp((size_t)A.n_rows + 1);
  JP_CUDA(cudaStreamSynchronize(ctx().compute));
  JP_CUDA(cudaMemcpy(ci.data(), A.colind.p, (size_t)A.nnz * 4, cudaMemcpyDeviceToHost));
  JP_CUDA(cudaMemcpy(rp.data(), A.rowptr.p, rp.size() * 4, cudaMemcpyDeviceToHost));
  std::vector<uint8_t> cls((size_t)A.n_rows, 0);
  for (int32_t r = 0; r < A.n_rows; ++r) {
    int32_t mx = -1;
    for (int32_t i = rp[r]; i < rp[r + 1]; ++i) mx = std::max(mx, ci[i]);
    cls[r] = mx >= A.n_local_cols;
  }
  std::vector<BlockDesc> bi, bb, ba;
  build_blocks(rp, cls, cap, max_rows, true, bi, bb, ba);
  // the widest tile that still leaves two CTAs per SM
  const int max_tile_cols = (int)((110 * 1024 - (int64_t)runs_smem_bytes(cap, 0, kt)) / (8 * kt)) & ~1;
  RunTables t;
  if (max_tile_cols < 64 || !build_run_tables(ci, ba, max_tile_cols, t)) {
    A.runs_state = -1;
    return;
  }
  size_t j = 0;  // the interior list is a subsequence of the all-list
  for (BlockDesc& d : bi) {
    while (ba[j].row0 != d.row0) ++j;
    d = ba[j];
  }
  A.runs.alloc(t.runs.size() * 4);
  A.runs_lidx.alloc(t.lidx.size() * 2);
  JP_CUDA(cudaMemcpy(A.runs.p, t.runs.data(), t.runs.size() * 4, cudaMemcpyHostToDevice));
  JP_CUDA(cudaMemcpy(A.runs_lidx.p, t.lidx.data(), t.lidx.size() * 2, cudaMemcpyHostToDevice));
  upload_blocks(ba, A.runs_blocks_all, A.n_runs_blocks_all);
  upload_blocks(bi, A.runs_blocks_interior, A.n_runs_blocks_interior);
  A.runs_wmax = t.wmax;
  A.runs_cap = cap;
  A.runs_state = 1;
}

template <int KT, int THREADS>
void launch_spmm_runs_impl(const Csr& A, const BlockDesc* b, int32_t n_blocks, const double* x, int64_t ldx, double* y, int64_t ldy, int32_t k,
                           double alpha, double beta, cudaStream_t stream) {
  const int wpad = (A.runs_wmax + 1) & ~1;
  const size_t smem = runs_smem_bytes(A.runs_cap, wpad, KT);
  set_smem_attr(spmm_runs_kernel<KT, THREADS>, smem);
  spmm_runs_kernel<KT, THREADS><<<n_blocks, THREADS, smem, stream>>>(b, A.rowptr.as<int32_t>(), A.vals.as<double>(), A.runs_lidx.as<uint16_t>(),
                                                                      A.runs.as<int2>(), x, ldx, y, ldy, k, A.runs_cap, wpad, alpha, beta);
}

}  // namespace

// the argument block shared by the two fused launches; `halo` is this apply's halo buffer (epoch parity applied)
static FusedSpmvArgs fused_args(const Csr& A, const FusedHalo& h, const double* x, const double* halo, double* y, double alpha, double beta) {
  FusedSpmvArgs a;
  a.interior = A.blocks_interior.as<BlockDesc>();
  a.boundary = A.blocks_boundary.as<BlockDesc>();
  a.n_interior = A.n_blocks_interior;
  a.n_boundary = A.n_blocks_boundary;
  static const int pack_rows = std::max(64, env_int("JP_FUSED_PACK_ROWS", 1024));  // exported rows per pack item
  a.n_items = std::max(1, std::min(64, (h.n_export + pack_rows - 1) / pack_rows));
  a.n_pack = a.n_items;  // one dedicated CTA per item at the head of the grid (boundary CTAs claim whatever is left)
  static const int first_pct = std::min(100, std::max(0, env_int("JP_FUSED_BOUNDARY_AT", 75)));  // % of the interior blocks in front
  a.n_first = (int32_t)((int64_t)a.n_interior * first_pct / 100);
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
  return a;
}

void launch_spmv_fused(const Csr& A, const FusedHalo& h, const double* x, const double* halo, double* y, double alpha, double beta,
                       cudaStream_t stream) {
  const FusedSpmvArgs a = fused_args(A, h, x, halo, y, alpha, beta);
  const size_t smem = stage_smem_bytes(A.nnz_cap);
  set_smem_attr(spmv_fused_halo_kernel, smem);
  const int grid = a.n_pack + a.n_interior + a.n_boundary;
  spmv_fused_halo_kernel<<<grid, kSpmvThreads, smem, stream>>>(a);
  JP_LAUNCH_CHECK();
}

// decide (once per matrix) which SpMM path k > 1 applies take
void csr_prepare_spmm(Csr& A) {
  csr_prepare_dia(A);
  if (A.runs_state == 0) {
    // a single-rank diagonal-structured matrix is served by the diagonal kernel alone: no run tables needed
    if (A.dia_state == 1) A.runs_state = -1;
    else build_runs(A);
  }
}

void launch_spmm_gather_blocks(const Csr& A, const void* blocks, int32_t n_blocks, const double* x, int64_t ldx, double* y, int64_t ldy,
                               int32_t k, double alpha, double beta, cudaStream_t stream) {
  if (n_blocks == 0) return;
  launch_spmm_impl<false, 1>(A, static_cast<const BlockDesc*>(blocks), n_blocks, stage_smem_bytes(A.nnz_cap), x, ldx, nullptr, 1, y, ldy, k, alpha,
                             beta, stream);
  JP_LAUNCH_CHECK();
}

// jp_apply_dot workspace of `n` CTA partials: [partials n | level-2 partials | result | ticket]; zeroed when (re)allocated
struct DotWorkspace {
  double* partials;
  double* level2;
  double* out;
  unsigned int* ticket;
};
static DotWorkspace dot_workspace(Csr& A, int32_t n, cudaStream_t stream) {
  const unsigned g2 = sum_partials_grid((unsigned)n);
  const size_t bytes = ((size_t)n + g2 + 4) * 8;
  if (A.dot_ws.bytes < bytes) {
    A.dot_ws.alloc(bytes);
    JP_CUDA(cudaMemsetAsync(A.dot_ws.p, 0, bytes, stream));
  }
  DotWorkspace w;
  w.partials = A.dot_ws.as<double>();
  w.level2 = w.partials + n;
  w.out = w.level2 + g2;
  w.ticket = reinterpret_cast<unsigned int*>(w.out + 1);
  return w;
}

double* launch_spmv_dot(Csr& A, const double* x, double* y, double alpha, double beta, const double* w, cudaStream_t stream) {
  const int32_t nb = A.n_blocks_all;
  const int32_t np = nb * kDotPartialsPerBlock;  // one partial per warp
  const DotWorkspace ws = dot_workspace(A, np, stream);
  if (nb == 0) {
    JP_CUDA(cudaMemsetAsync(ws.out, 0, 8, stream));
    return ws.out;
  }
  const size_t smem = stage_smem_bytes(A.nnz_cap);
  const BlockDesc* b = A.blocks_all.as<BlockDesc>();
  set_smem_attr(spmv_dot_kernel<false, 1>, smem);
  spmv_dot_kernel<false, 1><<<nb, kSpmvThreads, smem, stream>>>(b, A.rowptr.as<int32_t>(), A.colind.as<int32_t>(), A.vals.as<double>(), x,
                                                                 nullptr, A.n_cols, y, A.nnz_cap, alpha, beta, w, ws.partials);
  JP_LAUNCH_CHECK();
  sum_partials_kernel<<<sum_partials_grid((unsigned)np), kSpmvThreads, 0, stream>>>(ws.partials, (unsigned)np, ws.out, ws.level2, ws.ticket);
  JP_LAUNCH_CHECK();
  return ws.out;
}

// the fused halo + SpMV launch of launch_spmv_fused with the dot product of jp_apply_dot accumulated on the way
double* launch_spmv_fused_dot(Csr& A, const FusedHalo& h, const double* x, const double* halo, double* y, double alpha, double beta,
                              const double* w, cudaStream_t stream) {
  const FusedSpmvArgs a = fused_args(A, h, x, halo, y, alpha, beta);
  const int grid = a.n_pack + a.n_interior + a.n_boundary;
  const int32_t np = grid * kDotPartialsPerBlock;  // one partial per warp
  const DotWorkspace ws = dot_workspace(A, np, stream);
  const size_t smem = stage_smem_bytes(A.nnz_cap);
  set_smem_attr(spmv_fused_halo_dot_kernel, smem);
  spmv_fused_halo_dot_kernel<<<grid, kSpmvThreads, smem, stream>>>(a, w, ws.partials);
  JP_LAUNCH_CHECK();
  sum_partials_kernel<<<sum_partials_grid((unsigned)np), kSpmvThreads, 0, stream>>>(ws.partials, (unsigned)np, ws.out, ws.level2, ws.ticket);
  JP_LAUNCH_CHECK();
  return ws.out;
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
  // lazily decide the SpMM path of this matrix
  if (k > 1) csr_prepare_spmm(const_cast<Csr&>(A));
  // the run-staged kernel serves launches over the WHOLE all-blocks or interior list (it has its own block lists with a
  // different geometry; a chunk of a list, as jp_apply_host launches, stays with the gather kernel)
  const bool whole_all = blocks_int4 == A.blocks_all.p && n_blocks == A.n_blocks_all;
  // (the interior-only launch is the compute-stream half of the multi-GPU two-stream path)
  const bool whole_interior = blocks_int4 == A.blocks_interior.p && n_blocks == A.n_blocks_interior;
  if (k == 1 && !use_halo && whole_all) {  // a large irregular matrix sweeps its nonzeros column slab by column slab
    csr_prepare_colblock(const_cast<Csr&>(A));
    if (A.cb_state == 1) {
      launch_spmv_colblock(A, x, y, alpha, beta, stream);
      return;
    }
  }
  const bool tma_ok = (ldx & 1) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;  // bulk copies need 16-byte aligned sources
  if (k > 1 && A.dia_state == 1 && !use_halo && tma_ok && whole_all) {
    launch_spmm_dia(A, x, ldx, y, ldy, k, alpha, beta, stream);  // diagonal-structured matrix
    return;
  }
  if (k > 1 && A.runs_state == 1 && !use_halo && tma_ok && (whole_all || whole_interior)) {
    // banded matrix, no halo columns in this launch
    const bool all = whole_all;
    const BlockDesc* rb = (all ? A.runs_blocks_all : A.runs_blocks_interior).as<BlockDesc>();
    const int32_t rn = all ? A.n_runs_blocks_all : A.n_runs_blocks_interior;
    if (rn > 0) {
      launch_spmm_runs_impl<kRunsKT, kRunsThreads>(A, rb, rn, x, ldx, y, ldy, k, alpha, beta, stream);
      JP_LAUNCH_CHECK();
    }
    return;
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

// y (A's column map, ny rows) = beta*y + alpha * A^T * x (x on A's row map): the ordinary row-block kernels on the
// transposed matrix (jp_transpose.cu) -- no atomics, deterministic, the entries of a column added in ascending row order
// like the reference's row sweep (src/CSRMatrix.jl:1147-1160)
void launch_spmv_trans(Csr& A, const double* x, int64_t ldx, double* y, int64_t ldy, int32_t ny, int32_t k, double alpha, double beta,
                       cudaStream_t stream) {
  if (k == 0 || ny == 0) return;
  Csr& T = csr_transposed(A);
  JP_REQUIRE(ny == T.n_rows, "transposed apply: Y must have one row per column-map entry");
  launch_spmv(T, T.blocks_all, T.n_blocks_all, false, x, ldx, nullptr, 1, y, ldy, k, alpha, beta, stream);
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
  DirectRule rule;
  rule.maxlen = env_int("JP_SPMV_DIRECT_MAXLEN", rule.maxlen);
  if (const char* v = std::getenv("JP_SPMV_DIRECT_RATIO"))
    if (*v) rule.ratio = std::atof(v);
  JP_REQUIRE(rule.maxlen >= 1 && rule.ratio >= 1.0, "JP_SPMV_DIRECT_MAXLEN must be >= 1 and JP_SPMV_DIRECT_RATIO >= 1");
  if (const char* v = std::getenv("JP_SPMV_SUBWARP_DIV"))  // lanes per irregular row = mean row length / this (tuning knob)
    if (*v) rule.subwarp_div = std::atof(v);
  JP_REQUIRE(rule.subwarp_div > 0.0, "JP_SPMV_SUBWARP_DIV must be > 0");
  build_blocks(rp, cls, cap, max_rows, env_int("JP_SPMV_DIRECT", 1) != 0, bi, bb, ba, rule);
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
  // chunks of the interior list (host-buffer path with an importer)
  if (!bi.empty() && !bb.empty()) {
    const int nic = std::max(1, std::min<int>(env_int("JP_HOST_CHUNKS", 16), (int)bi.size()));
    for (int c = 0; c <= nic; ++c) {
      const int32_t bidx = (int32_t)((int64_t)bi.size() * c / nic);
      A.ichunk_block.push_back(bidx);
      A.ichunk_row.push_back(c == 0 ? 0 : (bidx < (int32_t)bi.size() ? bi[bidx].row0 : n_rows));
    }
    int32_t irun = -1;
    for (int c = 0; c < nic; ++c) {
      uint8_t hasb = 0;
      for (int32_t r = A.ichunk_row[c]; r < A.ichunk_row[c + 1]; ++r) {
        if (cls[r]) hasb = 1;
        else irun = std::max(irun, rowmax[r]);
      }
      A.ichunk_maxcol.push_back(irun);
      A.ichunk_has_boundary.push_back(hasb);
    }
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
    // host pass over the row pointers only (n_rows + 1 integers)
    std::vector<int32_t> rp((size_t)n_rows + 1);
    JP_REQUIRE(rowptr[0] == index_base, "jp_csr_create: rowptr[first] must equal index_base");
    for (int32_t r = 0; r <= n_rows; ++r) {
      rp[r] = rowptr[r] - index_base;
      if (r > 0) JP_REQUIRE(rp[r] >= rp[r - 1], "jp_csr_create: rowptr must be non-decreasing");
    }
    JP_REQUIRE(rp[n_rows] == nnz, "jp_csr_create: rowptr[last] does not match nnz");
    // device arrays; 64 elements of zeroed slack so the 16-byte staging may over-read past nnz
    cudaStream_t st = ctx().compute;
    A->rowptr.alloc(((size_t)n_rows + 1 + 16) * 4);  // slack: row pointers are staged in aligned groups of 4
    JP_CUDA(cudaMemsetAsync(A->rowptr.p, 0, ((size_t)n_rows + 1 + 16) * 4, st));
    A->colind.alloc(((size_t)nnz + 64) * 4);
    A->vals.alloc(((size_t)nnz + 64) * 8);
    JP_CUDA(cudaMemcpyAsync(A->rowptr.p, rp.data(), rp.size() * 4, cudaMemcpyHostToDevice, st));
    JP_CUDA(cudaMemsetAsync(A->colind.as<int32_t>() + nnz, 0, 64 * 4, st));
    JP_CUDA(cudaMemsetAsync(A->vals.as<double>() + nnz, 0, 64 * 8, st));
    if (nnz > 0) {
      JP_CUDA(cudaMemcpyAsync(A->colind.p, colind, (size_t)nnz * 4, cudaMemcpyHostToDevice, st));
      JP_CUDA(cudaMemcpyAsync(A->vals.p, vals, (size_t)nnz * 8, cudaMemcpyHostToDevice, st));
    }
    // device pass over the column indices: 0-based, range check, every row's largest column
    std::vector<int32_t> rowmax((size_t)n_rows, -1);
    if (n_rows > 0) {
      DeviceBuffer d_rowmax, d_bad;
      d_rowmax.alloc((size_t)n_rows * 4);
      d_bad.alloc(8);
      JP_CUDA(cudaMemsetAsync(d_bad.p, 0, 8, st));
      const int64_t want = ((int64_t)n_rows + 255) / 256;
      csr_validate_kernel<<<(int)std::min<int64_t>(want, (int64_t)ctx().sm_count * 16), 256, 0, st>>>(
          A->rowptr.as<int32_t>(), A->colind.as<int32_t>(), n_rows, n_cols, index_base, d_rowmax.as<int32_t>(), d_bad.as<unsigned long long>());
      JP_LAUNCH_CHECK();
      unsigned long long bad = 0;
      JP_CUDA(cudaMemcpyAsync(rowmax.data(), d_rowmax.p, (size_t)n_rows * 4, cudaMemcpyDeviceToHost, st));
      JP_CUDA(cudaMemcpyAsync(&bad, d_bad.p, 8, cudaMemcpyDeviceToHost, st));
      JP_CUDA(cudaStreamSynchronize(st));
      JP_REQUIRE(bad == 0, "jp_csr_create: column index outside the column map");
    } else {
      JP_CUDA(cudaStreamSynchronize(st));  // the host arrays are not retained
    }
    csr_finalize_structure(*A, rp, rowmax);
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

int32_t jp_csr_spmm_info(jp_handle A, int64_t* info) {
  return guarded([&] {
    Csr* a = get<Csr>(A, Kind::Csr, "csr matrix");
    JP_REQUIRE(info != nullptr, "jp_csr_spmm_info: null output");
    info[0] = -1;  // (the shared-memory tile tables of round 1 are gone)
    info[1] = a->runs_state;
    info[2] = a->n_runs_blocks_all;
    info[3] = a->runs_wmax;
    info[4] = a->dia_state;
    info[5] = csr_dia_info(*a, 0);
    info[6] = csr_dia_info(*a, 1);
    info[7] = csr_dia_info(*a, 2);
    info[8] = a->cb_state;
    info[9] = csr_colblock_info(*a, 0);
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
