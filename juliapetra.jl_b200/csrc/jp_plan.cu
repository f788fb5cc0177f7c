// jp_plan.cu -- Import/Export plans on the device, the halo exchange, and apply!.
//
// Replaces, for MultiVectors:
//   doTransfer            src/DistObject.jl:200-258
//   copyAndPermute        src/MultiVector.jl:308-325   (copy kernel; skipped entirely on the fused apply path)
//   packAndPrepare        src/MultiVector.jl:327-349   (pack kernel -> LID-major, vector-minor send buffer)
//   resolvePosts/Waits    src/MPIDistributor.jl:451-590 (one ncclGroup of send/recv per neighbour on the comm
//                         stream; message sizes are static per plan, so the reference's 4 barriers, 2 size
//                         handshakes and byte (de)serialisation per exchange disappear)
//   unpackAndCombine      src/MultiVector.jl:351-371   (unpack kernel; not needed on the fused path because the
//                         receives land directly in the halo buffer the boundary kernel reads)
//   apply!                src/RowMatrix.jl:261-357
//
// apply! NO_TRANS with a plan whose remote LIDs are the contiguous tail of the column map (what __makeColMap
// produces, src/CSRGraphInternalMethods.jl:946-981) never copies the local part of X (the reference's
// copyAndPermute K8 is gone) and has three halo transports, chosen at run time:
//   1. k == 1, peers IPC-mappable: ONE kernel (spmv_fused_halo_kernel, jp_csr.cu) = pack + NVLink peer stores +
//      interior rows + flag wait + boundary rows + ack
//   2. peer memory, two streams: comm   : wait(X ready) -> pack_p2p -> halo_wait -> boundary rows -> halo_ack -> record
//                                compute: interior rows, concurrently -> wait(done)
//   3. NCCL, two streams:        comm   : wait(X ready) -> pack -> ncclGroup{recv, send} -> boundary rows -> record
// TRANS with a plan runs the exchange backwards with ADD (apply_trans_distributed).
#include <cstdlib>

#include "jp_common.h"
#include "jp_p2p.cuh"
#include "jp_plan_kernels.cuh"

namespace jp {

namespace {
using namespace plank;

int grid_for(int64_t n) {
  int64_t want = (n + kThreads - 1) / kThreads;
  int64_t cap = (int64_t)ctx().sm_count * 8;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}

void upload_lids(DeviceBuffer& buf, const int32_t* host, int32_t n, int32_t base, int32_t limit, const char* what) {
  std::vector<int32_t> v((size_t)n);
  for (int32_t i = 0; i < n; ++i) {
    v[i] = host[i] - base;
    JP_REQUIRE(v[i] >= 0 && v[i] < limit, std::string("jp_plan_create: ") + what + " out of range");
  }
  buf.alloc((size_t)n * 4 + 16);
  if (n > 0) JP_CUDA(cudaMemcpy(buf.p, v.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
}

// the exchange proper, on the comm stream: sendbuf -> peers' recvbuf
void exchange(Plan& p, int32_t k, bool reverse, cudaStream_t st) {
  Comm& c = *p.comm;
  const std::vector<int32_t>& to = reverse ? p.procs_from : p.procs_to;
  const std::vector<int64_t>& lto = reverse ? p.lengths_from : p.lengths_to;
  const std::vector<int64_t>& sto = reverse ? p.starts_from : p.starts_to;
  const std::vector<int32_t>& from = reverse ? p.procs_to : p.procs_from;
  const std::vector<int64_t>& lfrom = reverse ? p.lengths_to : p.lengths_from;
  const std::vector<int64_t>& sfrom = reverse ? p.starts_to : p.starts_from;
  double* sb = p.sendbuf.as<double>();
  double* rb = p.recvbuf.as<double>();
  const int me = c.rank0 + 1;
  bool any_remote = false;
  for (size_t i = 0; i < from.size(); ++i) any_remote |= (from[i] != me && lfrom[i] > 0);
  for (size_t i = 0; i < to.size(); ++i) any_remote |= (to[i] != me && lto[i] > 0);
  if (any_remote) {
    JP_REQUIRE_STATE(!c.serial, "plan names remote ranks but its comm is serial");
    JP_NCCL(ncclGroupStart());
    for (size_t i = 0; i < from.size(); ++i)
      if (from[i] != me && lfrom[i] > 0)
        JP_NCCL(ncclRecv(rb + (size_t)sfrom[i] * k, (size_t)lfrom[i] * k, ncclDouble, from[i] - 1, c.nccl, st));
    for (size_t i = 0; i < to.size(); ++i)
      if (to[i] != me && lto[i] > 0)
        JP_NCCL(ncclSend(sb + (size_t)sto[i] * k, (size_t)lto[i] * k, ncclDouble, to[i] - 1, c.nccl, st));
    JP_NCCL(ncclGroupEnd());
  }
  // self message: aliased in the reference (src/MPIDistributor.jl:559-561), a device copy here
  for (size_t i = 0; i < to.size(); ++i) {
    if (to[i] != me || lto[i] == 0) continue;
    for (size_t r = 0; r < from.size(); ++r)
      if (from[r] == me)
        JP_CUDA(cudaMemcpyAsync(rb + (size_t)sfrom[r] * k, sb + (size_t)sto[i] * k, (size_t)lto[i] * k * 8, cudaMemcpyDeviceToDevice, st));
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Peer-to-peer halo for the fused apply.  The halo of one apply is a few hundred KiB per neighbour (512 KiB for
// 256^3 slabs), so its cost is latency, not bandwidth: an NCCL send/recv group costs ~25 us per exchange, the same
// order as the whole interior SpMV at 8 GPUs.  Here the pack kernel IS the send: it gathers the exported rows of X
// and stores them straight into the receiver's halo buffer through an IPC-mapped peer pointer (NVLink stores), then
// publishes an epoch flag; the receiver's comm stream waits on the flag in a one-warp kernel right before the
// boundary-row kernel.  Every rank has TWO halo buffers, selected by the parity of the epoch, and an ack flag per
// source, so a sender can never overwrite a halo that is still being read and neighbours are not lock-stepped
// (protocol: jp_p2p.cuh).
// ---------------------------------------------------------------------------------------------------------------
// fused pack + send: dst[(off + i - start)*k + j] = src[lid[i] + j*ld] for every export entry i, into the receivers' buffers.
__global__ void __launch_bounds__(kThreads) pack_p2p_kernel(const double* __restrict__ src, int64_t ld, const int32_t* __restrict__ lids,
                                                              int32_t n, int32_t k, const P2PSend* __restrict__ table, int32_t n_blocks,
                                                              uint64_t epoch, unsigned int* __restrict__ ticket,
                                                              unsigned long long* timeout_flag) {
  __shared__ bool is_last;
  // back-pressure: the receiver must have consumed what this buffer held two applies ago
  if (threadIdx.x < n_blocks && epoch > 2) wait_flag_ge(table[threadIdx.x].ack_flag, epoch - 2, timeout_flag);
  __syncthreads();
  const int64_t total = (int64_t)n * k;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += stride) {
    const int32_t j = (int32_t)(t / n), i = (int32_t)(t - (int64_t)j * n);
    int b = 0;
    while (b + 1 < n_blocks && i >= table[b].start + table[b].count) ++b;
    const P2PSend& s = table[b];
    s.dst_base[(epoch & 1) * s.dst_buf_elems + (size_t)(s.dst_off + i - s.start) * k + j] = __ldg(src + __ldg(lids + i) + (size_t)j * ld);
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (is_last) {
    __threadfence_system();
    if (threadIdx.x == 0) *ticket = 0;
    if (threadIdx.x < n_blocks) st_release_sys(table[threadIdx.x].ready_flag, epoch);
  }
}

// one warp: wait until every source has published this epoch's halo
__global__ void halo_wait_kernel(const P2PSource* __restrict__ sources, int32_t n_sources, uint64_t epoch,
                                 unsigned long long* timeout_flag) {
  if ((int)threadIdx.x < n_sources) wait_flag_ge(sources[threadIdx.x].ready_flag, epoch, timeout_flag);
}

// one warp: tell every source that this epoch's halo has been consumed
__global__ void halo_ack_kernel(const P2PSource* __restrict__ sources, int32_t n_sources, uint64_t epoch) {
  if ((int)threadIdx.x < n_sources) st_release_sys(sources[threadIdx.x].ack_flag, epoch);
}

void close_peer_mappings(Plan& p) {
  const int me = p.comm ? p.comm->rank0 : -1;
  for (int q = 0; q < (int)p.peer_arena.size(); ++q)
    if (p.peer_arena[q] && q != me) cudaIpcCloseMemHandle(p.peer_arena[q]);
  p.peer_arena.clear();
}

// Collective over the plan's comm: build an arena with room for k_cap vectors per halo buffer, exchange IPC handles
// and receive offsets, map the neighbours.  Returns false ON EVERY RANK when the peer-memory halo cannot be used
// (a rank cannot export or map memory, more ranks / sources / destinations than the flag block and the kernels'
// one-thread-per-flag loops hold); the caller then uses NCCL send/recv.  Every feasibility condition is evaluated
// BEFORE the agreement all-reduce, so the ranks never disagree.
bool p2p_setup(Plan& p, int32_t k_cap) {
  Comm& c = *p.comm;
  Context& x = ctx();
  if (c.serial) return false;
  static const bool disabled = []() {
    const char* v = std::getenv("JP_HALO");
    return v && std::string(v) == "nccl";
  }();
  if (disabled) return false;
  JP_CUDA(cudaStreamSynchronize(x.comm));
  JP_CUDA(cudaStreamSynchronize(x.compute));
  const int P = c.nproc, me = c.rank0;
  // a previous, smaller arena: every rank closes its mappings, and only after all of them have done so is the memory
  // retired (an exported allocation must not be freed while an importer still has it open)
  if (!p.peer_arena.empty()) {
    close_peer_mappings(p);
    allreduce_min_host(c, 1);
    retire_arena(p.arena);
  }
  int ok = 1;
  size_t n_sends = 0, n_sources = 0;
  for (size_t i = 0; i < p.procs_to.size(); ++i) n_sends += p.lengths_to[i] > 0;
  for (size_t i = 0; i < p.procs_from.size(); ++i) n_sources += p.lengths_from[i] > 0;
  if (P > kP2PMaxRanks || n_sends > (size_t)kP2PMaxSends || n_sources > (size_t)kP2PMaxSources) ok = 0;
  p.peer_arena.assign(P, nullptr);
  p.arena_buf_bytes = (((size_t)std::max(p.n_remote, 1) * k_cap * 8) + 255) & ~(size_t)255;
  const size_t arena_bytes = kArenaFlagBytes + 2 * p.arena_buf_bytes;
  p.arena.alloc(arena_bytes);
  JP_CUDA(cudaMemset(p.arena.p, 0, arena_bytes));
  p.p2p_tickets.alloc(64);  // uint32 tickets: [0] pack items, [1] boundary blocks, [2] pack-item claims
  JP_CUDA(cudaMemset(p.p2p_tickets.p, 0, 64));
  p.epoch = 0;
  // what every rank publishes: IPC handle, buffer size, and for each source rank the row offset of its data
  struct Info {
    cudaIpcMemHandle_t handle;
    uint64_t buf_bytes;
    int64_t recv_off[kP2PMaxRanks];
  };
  std::vector<Info> all(P);
  Info mine;
  std::memset(&mine, 0, sizeof mine);
  if (cudaIpcGetMemHandle(&mine.handle, p.arena.p) != cudaSuccess) {
    cudaGetLastError();
    ok = 0;
  }
  mine.buf_bytes = p.arena_buf_bytes;
  for (int q = 0; q < kP2PMaxRanks; ++q) mine.recv_off[q] = -1;
  for (size_t i = 0; i < p.procs_from.size(); ++i)
    if (p.procs_from[i] - 1 < kP2PMaxRanks) mine.recv_off[p.procs_from[i] - 1] = p.starts_from[i];
  allgather_host_bytes(c, &mine, sizeof(Info), all.data());
  // map every neighbour (ranks I send to or receive from)
  std::vector<char> need(P, 0);
  for (int32_t q : p.procs_to) need[q - 1] = 1;
  for (int32_t q : p.procs_from) need[q - 1] = 1;
  p.peer_arena[me] = p.arena.p;
  for (int q = 0; q < P && ok; ++q) {
    if (!need[q] || q == me) continue;
    void* ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, all[q].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      ok = 0;
      break;
    }
    p.peer_arena[q] = ptr;
  }
  // every receiver must expect what I send
  for (size_t i = 0; i < p.procs_to.size() && ok; ++i)
    if (p.lengths_to[i] > 0 && (p.procs_to[i] - 1 >= kP2PMaxRanks || all[p.procs_to[i] - 1].recv_off[me] < 0)) ok = 0;
  // everybody or nobody
  if (!allreduce_min_host(c, ok)) {
    close_peer_mappings(p);
    allreduce_min_host(c, 1);  // all mappings are closed before any arena is released
    p.arena.release();
    return false;
  }
  // device tables
  std::vector<P2PSend> sends;
  for (size_t i = 0; i < p.procs_to.size(); ++i) {
    if (p.lengths_to[i] == 0) continue;
    const int q = p.procs_to[i] - 1;
    char* base = (char*)p.peer_arena[q];
    P2PSend s;
    s.start = (int32_t)p.starts_to[i];
    s.count = (int32_t)p.lengths_to[i];
    s.dst_base = (double*)(base + kArenaFlagBytes);
    s.dst_buf_elems = (int64_t)(all[q].buf_bytes / 8);
    s.dst_off = all[q].recv_off[me];
    s.ready_flag = (uint64_t*)base + me;
    s.ack_flag = (const uint64_t*)((char*)p.arena.p + 2048) + q;
    sends.push_back(s);
  }
  std::vector<P2PSource> sources;
  for (size_t i = 0; i < p.procs_from.size(); ++i) {
    if (p.lengths_from[i] == 0) continue;
    const int q = p.procs_from[i] - 1;
    P2PSource s;
    s.ready_flag = (const uint64_t*)p.arena.p + q;
    s.ack_flag = (uint64_t*)((char*)p.peer_arena[q] + 2048) + me;
    sources.push_back(s);
  }
  p.n_send_blocks = (int32_t)sends.size();
  p.n_sources = (int32_t)sources.size();
  p.send_table.alloc(sends.size() * sizeof(P2PSend) + 16);
  p.wait_table.alloc(sources.size() * sizeof(P2PSource) + 16);
  if (!sends.empty()) JP_CUDA(cudaMemcpy(p.send_table.p, sends.data(), sends.size() * sizeof(P2PSend), cudaMemcpyHostToDevice));
  if (!sources.empty()) JP_CUDA(cudaMemcpy(p.wait_table.p, sources.data(), sources.size() * sizeof(P2PSource), cudaMemcpyHostToDevice));
  // nobody may publish into an arena that is not mapped everywhere yet
  allreduce_min_host(c, 1);
  p.p2p_kcap = k_cap;
  return true;
}

// Distributed TRANS / CONJ_TRANS: src/RowMatrix.jl:324-352 with the semantics of
// Tpetra::CrsMatrix::applyTranspose (the reference's own lines cannot run: undefined names, ADD ignored):
//   YColMap = alpha * A_local^T * X  (scatter kernel into the cached column-map multivector, beta = 0)
//   Y = beta * Y;  doExport(YColMap, Y, importer, ADD) in reverse mode: local columns are added in place, every
//   remote column's partial sum travels back to its owner (the halo exchange run backwards) and is added there.
}  // namespace

void apply_trans_distributed(Csr* a, Plan* p, MultiVector* y, MultiVector* x, double alpha, double beta) {
  Context& c = ctx();
  const int32_t k = x->k;
  JP_REQUIRE(x->n == a->n_rows, "apply!: transposed apply needs X on the row map");
  JP_REQUIRE(y->n == p->n_src && p->n_tgt == a->n_cols, "apply!: transposed apply needs Y on the domain map of the importer");
  if (!a->import_mv || a->import_mv->k != k) {
    JP_CUDA(cudaStreamSynchronize(c.compute));
    JP_CUDA(cudaStreamSynchronize(c.comm));
    a->import_mv = std::make_unique<MultiVector>();
    a->import_mv->n = a->n_cols;
    a->import_mv->k = k;
    a->import_mv->data.alloc((size_t)a->n_cols * k * 8 + 64);
  }
  MultiVector* yc = a->import_mv.get();
  plan_ensure_buffers(*p, k);
  launch_spmv_trans(*a, x->d(), x->n, yc->d(), yc->n, yc->n, k, alpha, 0.0, c.compute);
  launch_scale_fill(y->d(), (int64_t)y->n * k, beta, c.compute);
  if (p->num_same > 0 && k > 0) {
    add_rows_kernel<<<grid_for((int64_t)p->num_same * k), kThreads, 0, c.compute>>>(y->d(), y->n, yc->d(), yc->n, p->num_same, k);
    JP_LAUNCH_CHECK();
  }
  if (p->n_permute > 0 && k > 0) {
    add_permute_kernel<<<grid_for((int64_t)p->n_permute * k), kThreads, 0, c.compute>>>(y->d(), y->n, yc->d(), yc->n, p->permute_from.as<int32_t>(),
                                                                                       p->permute_to.as<int32_t>(), p->n_permute, k);
    JP_LAUNCH_CHECK();
  }
  // the halo exchange run backwards: what was imported forward (remote LIDs) is exported now, on the comm stream
  JP_CUDA(cudaEventRecord(c.ev_x_ready, c.compute));
  JP_CUDA(cudaStreamWaitEvent(c.comm, c.ev_x_ready, 0));
  if (p->n_remote > 0 && k > 0) {
    pack_kernel<<<grid_for((int64_t)p->n_remote * k), kThreads, 0, c.comm>>>(yc->d(), yc->n, p->remote_lids.as<int32_t>(), p->n_remote, k,
                                                                            p->sendbuf.as<double>());
    JP_LAUNCH_CHECK();
  }
  exchange(*p, k, /*reverse=*/true, c.comm);
  JP_CUDA(cudaEventRecord(p->ev_done, c.comm));
  JP_CUDA(cudaStreamWaitEvent(c.compute, p->ev_done, 0));
  if (p->n_export > 0 && k > 0) {
    scatter_add_kernel<<<grid_for((int64_t)p->n_export * k), kThreads, 0, c.compute>>>(y->d(), y->n, p->export_lids.as<int32_t>(), p->n_export,
                                                                                      k, p->recvbuf.as<double>());
    JP_LAUNCH_CHECK();
  }
  JP_CUDA(cudaEventRecord(c.ev_tmp, c.compute));
  JP_CUDA(cudaStreamWaitEvent(c.comm, c.ev_tmp, 0));
}


void plan_ensure_buffers(Plan& p, int32_t k) {
  if (k <= p.buf_k) return;
  JP_CUDA(cudaStreamSynchronize(ctx().comm));
  JP_CUDA(cudaStreamSynchronize(ctx().compute));
  int64_t ns = std::max<int64_t>(p.n_export, 0), nr = std::max<int64_t>(p.n_remote, 0);
  // reverse mode swaps the roles of the two buffers' capacities
  int64_t cap = std::max(ns, nr);
  p.sendbuf.alloc((size_t)cap * k * 8 + 64);
  p.recvbuf.alloc((size_t)cap * k * 8 + 64);
  p.buf_k = k;
}

// comm stream: [copyAndPermute] + pack + exchange; records p.ev_done.  halo_only skips copyAndPermute (the fused
// apply reads the local part from X directly).
void plan_post(Plan& p, const MultiVector& src, MultiVector* tgt, bool halo_only, int32_t combine_mode, bool reverse) {
  Context& x = ctx();
  const int32_t k = src.k;
  plan_ensure_buffers(p, k);
  JP_CUDA(cudaEventRecord(x.ev_x_ready, x.compute));
  JP_CUDA(cudaStreamWaitEvent(x.comm, x.ev_x_ready, 0));
  if (!halo_only) {
    // copyAndPermute (src/MultiVector.jl:308-325)
    if (p.num_same > 0 && k > 0 && tgt->d() != src.d())
      JP_CUDA(cudaMemcpy2DAsync(tgt->d(), (size_t)tgt->n * 8, src.d(), (size_t)src.n * 8, (size_t)p.num_same * 8, (size_t)k,
                                cudaMemcpyDeviceToDevice, x.comm));
    if (p.n_permute > 0 && k > 0) {
      permute_kernel<<<grid_for((int64_t)p.n_permute * k), kThreads, 0, x.comm>>>(tgt->d(), tgt->n, src.d(), src.n,
                                                                                 p.permute_to.as<int32_t>(), p.permute_from.as<int32_t>(),
                                                                                 p.n_permute, k);
      JP_LAUNCH_CHECK();
    }
  }
  if (combine_mode != JP_ZERO && k > 0) {  // src/DistObject.jl:238
    if (p.n_export > 0) {
      pack_kernel<<<grid_for((int64_t)p.n_export * k), kThreads, 0, x.comm>>>(src.d(), src.n, p.export_lids.as<int32_t>(), p.n_export, k,
                                                                             p.sendbuf.as<double>());
      JP_LAUNCH_CHECK();
    }
    exchange(p, k, reverse, x.comm);
  }
  JP_CUDA(cudaEventRecord(p.ev_done, x.comm));
}

// tgt[remote_lids] = recvbuf on `stream` (must already be ordered after ev_done)
void plan_unpack(Plan& p, MultiVector& tgt, cudaStream_t stream) {
  if (p.n_remote == 0 || tgt.k == 0) return;
  unpack_kernel<<<grid_for((int64_t)p.n_remote * tgt.k), kThreads, 0, stream>>>(tgt.d(), tgt.n, p.remote_lids.as<int32_t>(), p.n_remote, tgt.k,
                                                                               p.recvbuf.as<double>());
  JP_LAUNCH_CHECK();
}

// Collective: make the peer-memory halo of `p` usable for k vectors (first use, or k exceeds the arena's capacity).
// Every rank reaches the same decision because applies are collective and carry the same k everywhere.
bool plan_p2p_ensure(Plan& p, int32_t k) {
  if (p.p2p_state == 0 || (p.p2p_state == 1 && k > p.p2p_kcap)) {
    int32_t cap = std::max(k, 4);  // room for small multivectors without a second collective setup
    if (p.p2p_state == 1) cap = std::max(cap, 2 * p.p2p_kcap);
    p.p2p_state = p2p_setup(p, cap) ? 1 : -1;
  }
  return p.p2p_state == 1;
}

// this apply's halo arguments for the fused kernel; advances the epoch
FusedHalo plan_p2p_next_epoch(Plan& p, const double** halo) {
  FusedHalo h;
  h.lids = p.export_lids.as<int32_t>();
  h.n_export = p.n_export;
  h.sends = p.send_table.as<P2PSend>();
  h.n_sends = p.n_send_blocks;
  h.sources = p.wait_table.as<P2PSource>();
  h.n_sources = p.n_sources;
  h.epoch = ++p.epoch;
  h.tickets = p.p2p_tickets.as<unsigned int>();
  h.timeout_flag = p2p_timeout_flag_device();
  *halo = reinterpret_cast<const double*>(static_cast<const char*>(p.arena.p) + kArenaFlagBytes + (p.epoch & 1) * p.arena_buf_bytes);
  return h;
}

Plan::~Plan() {
  if (ev_done) cudaEventDestroy(ev_done);
  // a peer may still have this rank's arena mapped (plans are destroyed in no particular order across ranks), so the
  // arena is retired -- kept until jp_finalize -- instead of freed
  for (int q = 0; q < (int)peer_arena.size(); ++q)
    if (peer_arena[q] && comm && q != comm->rank0) cudaIpcCloseMemHandle(peer_arena[q]);
  retire_arena(arena);
}

// Two-stream halo over peer memory (k > 1): returns the halo buffer of this epoch, or nullptr when the plan has to use
// the NCCL path.  Enqueues pack+send and the wait on the comm stream; the caller launches the boundary rows on
// the comm stream and then calls plan_p2p_ack().
const double* plan_post_p2p(Plan& p, const MultiVector& src) {
  Context& x = ctx();
  const int32_t k = src.k;
  if (!plan_p2p_ensure(p, k)) return nullptr;
  p.epoch += 1;
  JP_CUDA(cudaEventRecord(x.ev_x_ready, x.compute));
  JP_CUDA(cudaStreamWaitEvent(x.comm, x.ev_x_ready, 0));
  // always launched: it also advances the epoch
  pack_p2p_kernel<<<grid_for((int64_t)p.n_export * k), kThreads, 0, x.comm>>>(src.d(), src.n, p.export_lids.as<int32_t>(), p.n_export, k,
                                                                             p.send_table.as<P2PSend>(), p.n_send_blocks, p.epoch,
                                                                             p.p2p_tickets.as<unsigned int>(), p2p_timeout_flag_device());
  JP_LAUNCH_CHECK();
  if (p.n_sources > 0) {
    halo_wait_kernel<<<1, 32, 0, x.comm>>>(p.wait_table.as<P2PSource>(), p.n_sources, p.epoch, p2p_timeout_flag_device());
    JP_LAUNCH_CHECK();
  }
  return reinterpret_cast<const double*>((const char*)p.arena.p + kArenaFlagBytes + (p.epoch & 1) * p.arena_buf_bytes);
}

void plan_p2p_ack(Plan& p) {
  Context& x = ctx();
  if (p.n_sources > 0) {
    halo_ack_kernel<<<1, 32, 0, x.comm>>>(p.wait_table.as<P2PSource>(), p.n_sources, p.epoch);
    JP_LAUNCH_CHECK();
  }
}

}  // namespace jp

using namespace jp;

namespace {

void check_transfer_args(Plan* p, MultiVector* s, MultiVector* t, int32_t combine_mode, int32_t reverse) {
  JP_REQUIRE(combine_mode >= JP_ADD && combine_mode <= JP_ZERO, "bad CombineMode");
  // checkSizes, src/MultiVector.jl:301-306
  JP_REQUIRE(s->k == t->k, "checkSize() indicates that the destination object is not a legal target for redistribution from the "
                           "source object: MultiVectors must have the same number of columns");
  JP_REQUIRE(s->n == p->n_src && t->n == p->n_tgt, "multivector local lengths do not match the plan's source/target maps");
  if (reverse) {
    int64_t tot_to = 0, tot_from = 0;
    for (int64_t l : p->lengths_to) tot_to += l;
    for (int64_t l : p->lengths_from) tot_from += l;
    JP_REQUIRE(p->n_export == tot_from && p->n_remote == tot_to,
               "reverse-mode transfer with non-matching export/remote lists is undefined in the reference "
               "(src/DistObject.jl:104-157 pass the forward lists unswapped)");
  }
}

// the column map is [domain prefix | remote tail] (what __makeColMap produces): the local part of X is read in place and
// the received rows are addressed as column - n_local_cols, so no column-map copy of X is needed
bool halo_in_place(const Csr* a, const Plan* p, const MultiVector* x) {
  return p->remote_is_tail && p->n_permute == 0 && p->num_same == a->n_local_cols && p->num_same == x->n &&
         p->num_same + p->n_remote == a->n_cols;
}

bool fused_kernel_enabled() {
  static const bool on = []() {
    const char* v = std::getenv("JP_FUSED");  // JP_FUSED=0: two-stream sequence instead of the one-kernel apply (A/B measurements)
    return !(v && *v == '0');
  }();
  return on;
}

// apply!'s NO_TRANS body on device objects; X is a domain-map multivector
void apply_no_trans(Csr* a, Plan* p, MultiVector* y, MultiVector* x, double alpha, double beta) {
  Context& c = ctx();
  JP_REQUIRE(y->n == a->n_rows, "apply!: Y must have one row per local matrix row");
  if (p == nullptr) {
    // importer === nothing: XColMap = X (src/RowMatrix.jl:290-291)
    JP_REQUIRE(x->n == a->n_cols, "apply!: X does not match the column map and no importer was given");
    launch_spmv(*a, a->blocks_all, a->n_blocks_all, false, x->d(), x->n, nullptr, 1, y->d(), y->n, y->k, alpha, beta, c.compute);
    return;
  }
  JP_REQUIRE(x->n == p->n_src, "apply!: X does not match the importer's source map");
  JP_REQUIRE(p->n_tgt == a->n_cols, "apply!: the importer's target map is not the matrix's column map");
  const bool fused = halo_in_place(a, p, x);
  if (!fused) {
    // general column map (unused local columns => permutes): import into the cached column-map multivector
    // exactly as the reference does (src/RowMatrix.jl:294-295), then one launch over all row blocks
    if (!a->import_mv || a->import_mv->k != x->k) {  // createColumnMapMultiVector, src/RowMatrix.jl:210-233
      JP_CUDA(cudaStreamSynchronize(c.compute));
      JP_CUDA(cudaStreamSynchronize(c.comm));
      a->import_mv = std::make_unique<MultiVector>();
      a->import_mv->n = a->n_cols;
      a->import_mv->k = x->k;
      a->import_mv->data.alloc((size_t)a->n_cols * x->k * 8 + 64);
      JP_CUDA(cudaMemsetAsync(a->import_mv->d(), 0, (size_t)a->n_cols * x->k * 8 + 64, c.compute));
    }
    MultiVector* xc = a->import_mv.get();
    plan_post(*p, *x, xc, /*halo_only=*/false, JP_INSERT, false);
    JP_CUDA(cudaStreamWaitEvent(c.compute, p->ev_done, 0));
    plan_unpack(*p, *xc, c.compute);
    launch_spmv(*a, a->blocks_all, a->n_blocks_all, false, xc->d(), xc->n, nullptr, 1, y->d(), y->n, y->k, alpha, beta, c.compute);
    JP_CUDA(cudaEventRecord(c.ev_tmp, c.compute));
    JP_CUDA(cudaStreamWaitEvent(c.comm, c.ev_tmp, 0));
    return;
  }
  // k == 1 with a peer-memory halo: ONE kernel does pack + send + interior rows + wait + boundary rows + ack
  // (spmv_fused_halo_kernel, csrc/jp_csr.cu)
  if (fused_kernel_enabled() && x->k == 1 && plan_p2p_ensure(*p, 1)) {
    const double* halo = nullptr;
    const FusedHalo h = plan_p2p_next_epoch(*p, &halo);
    launch_spmv_fused(*a, h, x->d(), halo, y->d(), alpha, beta, c.compute);
    return;
  }
  // Two-stream path.  comm stream (high priority): pack -> exchange -> BOUNDARY rows; compute stream: INTERIOR rows,
  // concurrently.  The two launches write disjoint rows of Y; the compute stream joins on the comm stream at the end, so
  // there is no serial boundary-kernel latency after the interior kernel's tail.  (Replaying this sequence as a CUDA
  // graph was measured slower than eager enqueue at 2 and 8 GPUs in round 1 -- profiles/r1_scale_*graph* -- and is gone.)
  plan_ensure_buffers(*p, x->k);
  const double* halo = plan_post_p2p(*p, *x);  // peer-memory halo; nullptr -> NCCL send/recv
  const bool p2p = halo != nullptr;
  if (!p2p) {
    plan_post(*p, *x, nullptr, /*halo_only=*/true, JP_INSERT, false);
    halo = p->recvbuf.as<double>();
  }
  launch_spmv(*a, a->blocks_boundary, a->n_blocks_boundary, true, x->d(), x->n, halo, y->k, y->d(), y->n, y->k, alpha, beta, c.comm);
  if (p2p) plan_p2p_ack(*p);
  JP_CUDA(cudaEventRecord(p->ev_done, c.comm));
  launch_spmv(*a, a->blocks_interior, a->n_blocks_interior, false, x->d(), x->n, nullptr, 1, y->d(), y->n, y->k, alpha, beta, c.compute);
  JP_CUDA(cudaStreamWaitEvent(c.compute, p->ev_done, 0));
}

// apply! when rangeMap != rowMap (src/RowMatrix.jl:298-308 and 324-330).  The reference's lines
// cannot run (undefined names; ADD ignored by unpackAndCombine; reverse mode passes the forward lists unswapped), so
// this follows Tpetra::CrsMatrix::apply, which they transcribe -- and the oracle's exportAdd / exportReverseInsert:
//   NO_TRANS: YRowMap = alpha * A * XColMap (the ordinary apply, into the cached row-map multivector, beta = 0);
//             Y = beta * Y; every row-map entry is ADDED to the range-map entry with the same GID: in place for same /
//             permuted ids, through the Export's distributor (pack -> send/recv -> scatter-add) for the others.
//   TRANS:    X (range map) is brought onto the row map first -- the Export run backwards with INSERT -- then the
//             ordinary transposed apply runs from the row-map multivector.
// Only kernels of the importer / transposed paths are involved (add_rows, add_permute, pack, scatter_add, permute,
// unpack); `ex` is the plan made from Export(rowMap, rangeMap).
void apply_impl(Csr* a, Plan* p, MultiVector* y, MultiVector* x, int32_t mode, double alpha, double beta);

void apply_with_exporter(Csr* a, Plan* imp, Plan* ex, MultiVector* y, MultiVector* x, int32_t mode, double alpha, double beta) {
  Context& c = ctx();
  JP_REQUIRE(mode == JP_NO_TRANS || mode == JP_TRANS || mode == JP_CONJ_TRANS, "apply!: bad TransposeMode");
  JP_REQUIRE(x != y, "apply!: X and Y must be different multivectors");
  JP_REQUIRE(x->k == y->k, "apply!: X and Y must have the same number of vectors");
  JP_REQUIRE(ex->n_src == a->n_rows, "apply!: the exporter's source map is not the matrix's row map");
  const int32_t k = x->k;
  if (alpha == 0.0) {  // src/RowMatrix.jl:271-278
    launch_scale_fill(y->d(), (int64_t)y->n * k, beta, c.compute);
    return;
  }
  if (!a->export_mv || a->export_mv->k != k) {  // createRowMapMultiVector, src/RowMatrix.jl:236-259
    JP_CUDA(cudaStreamSynchronize(c.compute));
    JP_CUDA(cudaStreamSynchronize(c.comm));
    a->export_mv = std::make_unique<MultiVector>();
    a->export_mv->n = a->n_rows;
    a->export_mv->k = k;
    a->export_mv->data.alloc((size_t)a->n_rows * k * 8 + 64);
    JP_CUDA(cudaMemsetAsync(a->export_mv->d(), 0, (size_t)a->n_rows * k * 8 + 64, c.compute));
  }
  MultiVector* rowmv = a->export_mv.get();
  plan_ensure_buffers(*ex, k);
  if (mode == JP_NO_TRANS) {
    JP_REQUIRE(y->n == ex->n_tgt, "apply!: Y does not match the exporter's target (range) map");
    apply_impl(a, imp, rowmv, x, JP_NO_TRANS, alpha, 0.0);  // leaves the compute stream ordered after all of it
    launch_scale_fill(y->d(), (int64_t)y->n * k, beta, c.compute);
    if (k == 0) return;
    if (ex->num_same > 0) {
      add_rows_kernel<<<grid_for((int64_t)ex->num_same * k), kThreads, 0, c.compute>>>(y->d(), y->n, rowmv->d(), rowmv->n, ex->num_same, k);
      JP_LAUNCH_CHECK();
    }
    if (ex->n_permute > 0) {
      add_permute_kernel<<<grid_for((int64_t)ex->n_permute * k), kThreads, 0, c.compute>>>(y->d(), y->n, rowmv->d(), rowmv->n,
                                                                                          ex->permute_to.as<int32_t>(),
                                                                                          ex->permute_from.as<int32_t>(), ex->n_permute, k);
      JP_LAUNCH_CHECK();
    }
    JP_CUDA(cudaEventRecord(c.ev_x_ready, c.compute));
    JP_CUDA(cudaStreamWaitEvent(c.comm, c.ev_x_ready, 0));
    if (ex->n_export > 0) {
      pack_kernel<<<grid_for((int64_t)ex->n_export * k), kThreads, 0, c.comm>>>(rowmv->d(), rowmv->n, ex->export_lids.as<int32_t>(),
                                                                               ex->n_export, k, ex->sendbuf.as<double>());
      JP_LAUNCH_CHECK();
    }
    exchange(*ex, k, /*reverse=*/false, c.comm);
    JP_CUDA(cudaEventRecord(ex->ev_done, c.comm));
    JP_CUDA(cudaStreamWaitEvent(c.compute, ex->ev_done, 0));
    if (ex->n_remote > 0) {
      scatter_add_kernel<<<grid_for((int64_t)ex->n_remote * k), kThreads, 0, c.compute>>>(y->d(), y->n, ex->remote_lids.as<int32_t>(),
                                                                                         ex->n_remote, k, ex->recvbuf.as<double>());
      JP_LAUNCH_CHECK();
    }
    JP_CUDA(cudaEventRecord(c.ev_tmp, c.compute));
    JP_CUDA(cudaStreamWaitEvent(c.comm, c.ev_tmp, 0));
    return;
  }
  // TRANS / CONJ_TRANS: XRowMap = the Export run backwards with INSERT
  JP_REQUIRE(x->n == ex->n_tgt, "apply!: X does not match the exporter's target (range) map");
  if (k > 0) {
    if (ex->num_same > 0)
      JP_CUDA(cudaMemcpy2DAsync(rowmv->d(), (size_t)rowmv->n * 8, x->d(), (size_t)x->n * 8, (size_t)ex->num_same * 8, (size_t)k,
                                cudaMemcpyDeviceToDevice, c.compute));
    if (ex->n_permute > 0) {  // rowmv[permuteFrom] = x[permuteTo]: the forward permutation inverted
      permute_kernel<<<grid_for((int64_t)ex->n_permute * k), kThreads, 0, c.compute>>>(rowmv->d(), rowmv->n, x->d(), x->n,
                                                                                      ex->permute_from.as<int32_t>(),
                                                                                      ex->permute_to.as<int32_t>(), ex->n_permute, k);
      JP_LAUNCH_CHECK();
    }
    JP_CUDA(cudaEventRecord(c.ev_x_ready, c.compute));
    JP_CUDA(cudaStreamWaitEvent(c.comm, c.ev_x_ready, 0));
    if (ex->n_remote > 0) {  // what the forward export delivers here travels back to the ranks that hold the row
      pack_kernel<<<grid_for((int64_t)ex->n_remote * k), kThreads, 0, c.comm>>>(x->d(), x->n, ex->remote_lids.as<int32_t>(), ex->n_remote, k,
                                                                               ex->sendbuf.as<double>());
      JP_LAUNCH_CHECK();
    }
    exchange(*ex, k, /*reverse=*/true, c.comm);
    JP_CUDA(cudaEventRecord(ex->ev_done, c.comm));
    JP_CUDA(cudaStreamWaitEvent(c.compute, ex->ev_done, 0));
    if (ex->n_export > 0) {
      unpack_kernel<<<grid_for((int64_t)ex->n_export * k), kThreads, 0, c.compute>>>(rowmv->d(), rowmv->n, ex->export_lids.as<int32_t>(),
                                                                                    ex->n_export, k, ex->recvbuf.as<double>());
      JP_LAUNCH_CHECK();
    }
    JP_CUDA(cudaEventRecord(c.ev_tmp, c.compute));
    JP_CUDA(cudaStreamWaitEvent(c.comm, c.ev_tmp, 0));
  }
  apply_impl(a, imp, y, rowmv, mode, alpha, beta);
}

void apply_impl(Csr* a, Plan* p, MultiVector* y, MultiVector* x, int32_t mode, double alpha, double beta) {
  JP_REQUIRE(mode == JP_NO_TRANS || mode == JP_TRANS || mode == JP_CONJ_TRANS, "apply!: bad TransposeMode");
  JP_REQUIRE(x != y, "apply!: X and Y must be different multivectors");
  JP_REQUIRE(x->k == y->k, "apply!: X and Y must have the same number of vectors");
  if (alpha == 0.0) {  // src/RowMatrix.jl:271-278
    launch_scale_fill(y->d(), (int64_t)y->n * y->k, beta, ctx().compute);
    return;
  }
  if (mode == JP_NO_TRANS) {
    apply_no_trans(a, p, y, x, alpha, beta);
  } else {
    if (p == nullptr) {
      JP_REQUIRE(x->n == a->n_rows && y->n == a->n_cols, "apply!: transposed apply needs X on the row map and Y on the column map");
      launch_spmv_trans(*a, x->d(), x->n, y->d(), y->n, y->n, y->k, alpha, beta, ctx().compute);
    } else {
      apply_trans_distributed(a, p, y, x, alpha, beta);
    }
  }
}

// device staging for jp_apply_host, keyed by shape
struct HostPathCache {
  std::unique_ptr<MultiVector> x, y;
};
HostPathCache g_host_cache;
std::vector<cudaEvent_t> g_host_events;

MultiVector* cached_mv(std::unique_ptr<MultiVector>& slot, int32_t n, int32_t k) {
  if (!slot || slot->n != n || slot->k != k) {
    JP_CUDA(cudaStreamSynchronize(ctx().compute));
    JP_CUDA(cudaStreamSynchronize(ctx().comm));
    slot = std::make_unique<MultiVector>();
    slot->n = n;
    slot->k = k;
    slot->data.alloc((size_t)n * k * 8 + 64);
  }
  return slot.get();
}

}  // namespace

namespace jp {
// y = beta*y + alpha*op(A)*x, and the LOCAL part of sum_r w[r, v] * y[r, v] for every vector v left in a device array of
// k doubles (compute stream; the caller adds the ranks' parts).  One kernel when the apply itself is one launch:
//   no importer, k = 1, NO_TRANS       spmv_dot_kernel
//   peer-memory halo, k = 1, NO_TRANS  spmv_fused_halo_dot_kernel (halo exchange + SpMV + dot)
// otherwise the apply followed, on the same stream, by the dot kernel (no host round trip in between).
double* apply_dot_local(Csr* a, Plan* p, MultiVector* y, MultiVector* x, MultiVector* w, int32_t mode, double alpha, double beta) {
  JP_REQUIRE(w->n == y->n && w->k == y->k, "jp_apply_dot: W must have the shape of Y");
  JP_REQUIRE(x != y, "apply!: X and Y must be different multivectors");
  JP_REQUIRE(x->k == y->k, "apply!: X and Y must have the same number of vectors");
  const bool one_launch = mode == JP_NO_TRANS && alpha != 0.0 && y->k == 1 && w != y && y->n == a->n_rows;
  if (one_launch && p == nullptr && x->n == a->n_cols) return launch_spmv_dot(*a, x->d(), y->d(), alpha, beta, w->d(), ctx().compute);
  if (one_launch && p != nullptr && fused_kernel_enabled() && x->n == p->n_src && p->n_tgt == a->n_cols && halo_in_place(a, p, x) &&
      plan_p2p_ensure(*p, 1)) {
    const double* halo = nullptr;
    const FusedHalo h = plan_p2p_next_epoch(*p, &halo);
    return launch_spmv_fused_dot(*a, h, x->d(), halo, y->d(), alpha, beta, w->d(), ctx().compute);
  }
  apply_impl(a, p, y, x, mode, alpha, beta);
  return mv_dot_local(*w, *y);
}

void release_host_path_cache() {
  g_host_cache.x.reset();
  g_host_cache.y.reset();
  for (cudaEvent_t e : g_host_events) cudaEventDestroy(e);
  g_host_events.clear();
}
}  // namespace jp

extern "C" {

int32_t jp_plan_create(jp_handle comm, int32_t n_src, int32_t n_tgt, int32_t num_same, const int32_t* permute_to,
                       const int32_t* permute_from, int32_t n_permute, const int32_t* remote_lids, int32_t n_remote,
                       const int32_t* export_lids, int32_t n_export, const int32_t* procs_to, const int64_t* lengths_to,
                       int32_t n_sends, const int32_t* procs_from, const int64_t* lengths_from, int32_t n_recvs,
                       int32_t index_base, jp_handle* out) {
  return guarded([&] {
    require_init();
    JP_REQUIRE(out != nullptr, "jp_plan_create: null output");
    Comm* c = get<Comm>(comm, Kind::Comm, "comm");
    JP_REQUIRE(index_base == 0 || index_base == 1, "jp_plan_create: index_base must be 0 or 1");
    JP_REQUIRE(n_src >= 0 && n_tgt >= 0 && n_permute >= 0 && n_remote >= 0 && n_export >= 0 && n_sends >= 0 && n_recvs >= 0,
               "jp_plan_create: negative size");
    JP_REQUIRE(num_same >= 0 && num_same <= n_src && num_same <= n_tgt, "jp_plan_create: numSameIDs out of range");
    auto p = std::make_unique<Plan>();
    p->comm = c;
    p->n_src = n_src;
    p->n_tgt = n_tgt;
    p->num_same = num_same;
    p->n_permute = n_permute;
    p->n_remote = n_remote;
    p->n_export = n_export;
    upload_lids(p->permute_to, permute_to, n_permute, index_base, n_tgt, "permuteToLIDs");
    upload_lids(p->permute_from, permute_from, n_permute, index_base, n_src, "permuteFromLIDs");
    upload_lids(p->remote_lids, remote_lids, n_remote, index_base, n_tgt, "remoteLIDs");
    upload_lids(p->export_lids, export_lids, n_export, index_base, n_src, "exportLIDs");
    p->remote_is_tail = true;
    for (int32_t i = 0; i < n_remote; ++i)
      if (remote_lids[i] - index_base != num_same + i) {
        p->remote_is_tail = false;
        break;
      }
    int64_t tot_to = 0, tot_from = 0;
    for (int32_t i = 0; i < n_sends; ++i) {
      JP_REQUIRE(procs_to[i] >= 1 && procs_to[i] <= c->nproc, "jp_plan_create: procs_to out of range");
      JP_REQUIRE(lengths_to[i] >= 0, "jp_plan_create: negative send length");
      p->procs_to.push_back(procs_to[i]);
      p->lengths_to.push_back(lengths_to[i]);
      p->starts_to.push_back(tot_to);
      tot_to += lengths_to[i];
    }
    for (int32_t i = 0; i < n_recvs; ++i) {
      JP_REQUIRE(procs_from[i] >= 1 && procs_from[i] <= c->nproc, "jp_plan_create: procs_from out of range");
      JP_REQUIRE(lengths_from[i] >= 0, "jp_plan_create: negative receive length");
      p->procs_from.push_back(procs_from[i]);
      p->lengths_from.push_back(lengths_from[i]);
      p->starts_from.push_back(tot_from);
      tot_from += lengths_from[i];
    }
    JP_REQUIRE(tot_to == n_export, "jp_plan_create: sum(lengths_to) must equal the number of export LIDs");
    JP_REQUIRE(tot_from == n_remote, "jp_plan_create: sum(lengths_from) must equal the number of remote LIDs");
    JP_CUDA(cudaEventCreateWithFlags(&p->ev_done, cudaEventDisableTiming));
    *out = register_object(std::move(p));
  });
}

int32_t jp_plan_destroy(jp_handle plan) {
  return guarded([&] { destroy_object(plan, Kind::Plan, "plan"); });
}

int32_t jp_transfer_post(jp_handle plan, jp_handle src, jp_handle tgt, int32_t combine_mode, int32_t reverse) {
  return guarded([&] {
    require_init();
    Plan* p = get<Plan>(plan, Kind::Plan, "plan");
    MultiVector* s = get<MultiVector>(src, Kind::MultiVector, "multivector");
    MultiVector* t = get<MultiVector>(tgt, Kind::MultiVector, "multivector");
    check_transfer_args(p, s, t, combine_mode, reverse);
    plan_post(*p, *s, t, false, combine_mode, reverse != 0);
    p->posted = true;
    p->posted_tgt = tgt;
    p->posted_k = combine_mode == JP_ZERO ? 0 : s->k;
  });
}

int32_t jp_transfer_wait(jp_handle plan) {
  return guarded([&] {
    require_init();
    Plan* p = get<Plan>(plan, Kind::Plan, "plan");
    // src/MPIDistributor.jl:574-576 / src/SerialComm.jl:52-54
    JP_REQUIRE_STATE(p->posted, "Cannot resolve waits when no posts have been made");
    p->posted = false;
    MultiVector* t = get<MultiVector>(p->posted_tgt, Kind::MultiVector, "multivector");
    Context& x = ctx();
    JP_CUDA(cudaStreamWaitEvent(x.compute, p->ev_done, 0));
    if (p->posted_k > 0) plan_unpack(*p, *t, x.compute);
    JP_CUDA(cudaEventRecord(x.ev_tmp, x.compute));
    JP_CUDA(cudaStreamWaitEvent(x.comm, x.ev_tmp, 0));
  });
}

int32_t jp_transfer(jp_handle plan, jp_handle src, jp_handle tgt, int32_t combine_mode, int32_t reverse) {
  int32_t rc = jp_transfer_post(plan, src, tgt, combine_mode, reverse);
  if (rc != JP_OK) return rc;
  return jp_transfer_wait(plan);
}

int32_t jp_apply(jp_handle comm, jp_handle Y, jp_handle A, jp_handle X, jp_handle plan, int32_t mode, double alpha, double beta) {
  return guarded([&] {
    require_init();
    (void)get<Comm>(comm, Kind::Comm, "comm");
    Csr* a = get<Csr>(A, Kind::Csr, "csr matrix");
    MultiVector* y = get<MultiVector>(Y, Kind::MultiVector, "multivector");
    MultiVector* x = get<MultiVector>(X, Kind::MultiVector, "multivector");
    Plan* p = plan ? get<Plan>(plan, Kind::Plan, "plan") : nullptr;
    apply_impl(a, p, y, x, mode, alpha, beta);
  });
}

int32_t jp_apply_export(jp_handle comm, jp_handle Y, jp_handle A, jp_handle X, jp_handle import_plan, jp_handle export_plan, int32_t mode,
                        double alpha, double beta) {
  return guarded([&] {
    require_init();
    (void)get<Comm>(comm, Kind::Comm, "comm");
    Csr* a = get<Csr>(A, Kind::Csr, "csr matrix");
    MultiVector* y = get<MultiVector>(Y, Kind::MultiVector, "multivector");
    MultiVector* x = get<MultiVector>(X, Kind::MultiVector, "multivector");
    Plan* imp = import_plan ? get<Plan>(import_plan, Kind::Plan, "plan") : nullptr;
    Plan* ex = get<Plan>(export_plan, Kind::Plan, "plan");
    apply_with_exporter(a, imp, ex, y, x, mode, alpha, beta);
  });
}

int32_t jp_apply_dot(jp_handle comm, jp_handle Y, jp_handle A, jp_handle X, jp_handle plan, int32_t mode, double alpha, double beta,
                     jp_handle W, double* out_k) {
  return guarded([&] {
    require_init();
    Comm* c = get<Comm>(comm, Kind::Comm, "comm");
    Csr* a = get<Csr>(A, Kind::Csr, "csr matrix");
    MultiVector* y = get<MultiVector>(Y, Kind::MultiVector, "multivector");
    MultiVector* x = get<MultiVector>(X, Kind::MultiVector, "multivector");
    MultiVector* w = get<MultiVector>(W, Kind::MultiVector, "multivector");
    Plan* p = plan ? get<Plan>(plan, Kind::Plan, "plan") : nullptr;
    if (y->k == 0) return;
    JP_REQUIRE(out_k != nullptr, "jp_apply_dot: null output");
    double* dev = apply_dot_local(a, p, y, x, w, mode, alpha, beta);
    mv_finish_reduction(*c, dev, y->k, out_k);
  });
}

int32_t jp_apply_host(jp_handle comm, jp_handle A, jp_handle plan, const double* x_host, double* y_host, int32_t k, int32_t mode,
                      double alpha, double beta) {
  return guarded([&] {
    require_init();
    (void)get<Comm>(comm, Kind::Comm, "comm");
    Csr* a = get<Csr>(A, Kind::Csr, "csr matrix");
    Plan* p = plan ? get<Plan>(plan, Kind::Plan, "plan") : nullptr;
    JP_REQUIRE(k >= 0 && x_host && y_host, "jp_apply_host: bad arguments");
    const bool trans = mode != JP_NO_TRANS;
    const int32_t nx = trans ? a->n_rows : (p ? p->n_src : a->n_cols);
    const int32_t ny = trans ? a->n_cols : a->n_rows;
    MultiVector* x = cached_mv(g_host_cache.x, nx, k);
    MultiVector* y = cached_mv(g_host_cache.y, ny, k);
    Context& c = ctx();
    const size_t xb = (size_t)nx * k * 8, yb = (size_t)ny * k * 8;
    const int nch = (int)a->chunk_maxcol.size();
    static const bool pipeline_enabled = []() {
      const char* v = std::getenv("JP_HOST_PIPELINE");
      return !(v && *v == '0');
    }();
    // below this many bytes of X the chunked pipeline is not worth its events (the tests lower it to cover the pipeline)
    static const size_t pipeline_min_bytes = []() {
      const char* v = std::getenv("JP_HOST_PIPELINE_MIN_BYTES");
      return (v && *v) ? (size_t)std::atoll(v) : ((size_t)4 << 20);
    }();
    // The run-staged SpMM kernel works on its own row blocks, not on chunks of the all-blocks list; when a matrix takes
    // it, a k > 1 host apply is one launch (like jp_apply), so both entry points give bit-identical results.
    if (k > 1 && !trans) csr_prepare_spmm(*a);
    const bool pipelined = pipeline_enabled && !trans && p == nullptr && beta == 0.0 && alpha != 0.0 && nch > 1 && xb >= pipeline_min_bytes &&
                           nx == a->n_cols && ny == a->n_rows && (k == 1 || a->runs_state != 1);
    // With an importer whose halo lands in place (the multi-GPU bench path): the same three-stream pipeline.  X goes up
    // in chunks on the comm stream, which then carries the halo exchange (pack + peer stores / NCCL) and the boundary
    // rows; the interior rows run chunk by chunk on the compute stream as their columns arrive; every chunk of Y goes
    // down on the copy stream as soon as the kernels that write it are done, while later pieces of X are still coming up.
    const int nic = (int)a->ichunk_maxcol.size();
    const bool plan_pipelined = pipeline_enabled && !trans && p != nullptr && beta == 0.0 && alpha != 0.0 && nic > 1 &&
                                xb >= pipeline_min_bytes && ny == a->n_rows && nx == p->n_src && p->n_tgt == a->n_cols && halo_in_place(a, p, x) &&
                                (k == 1 || a->runs_state != 1);
    if (plan_pipelined) {
      while ((int)g_host_events.size() < 2 * std::max(nch, nic) + 1) {
        cudaEvent_t e;
        JP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        g_host_events.push_back(e);
      }
      cudaStream_t h2d = c.comm, comp = c.compute, d2h = c.copy;
      JP_CUDA(cudaEventRecord(c.ev_tmp, comp));  // order after whatever the compute stream was doing
      JP_CUDA(cudaStreamWaitEvent(h2d, c.ev_tmp, 0));
      std::vector<int32_t> xr(nic + 1);
      for (int i = 0; i <= nic; ++i) xr[i] = (int32_t)((int64_t)nx * i / nic);
      for (int i = 0; i < nic; ++i) {
        const size_t w = (size_t)(xr[i + 1] - xr[i]) * 8;
        if (w)
          JP_CUDA(cudaMemcpy2DAsync(x->d() + xr[i], (size_t)nx * 8, x_host + xr[i], (size_t)nx * 8, w, (size_t)k, cudaMemcpyHostToDevice, h2d));
        JP_CUDA(cudaEventRecord(g_host_events[i], h2d));
      }
      // comm stream, behind the last piece of X: halo exchange, then the boundary rows
      plan_ensure_buffers(*p, k);
      const double* halo = plan_post_p2p(*p, *x);  // nullptr -> NCCL send/recv
      const bool p2p = halo != nullptr;
      if (!p2p) {
        plan_post(*p, *x, nullptr, /*halo_only=*/true, JP_INSERT, false);
        halo = p->recvbuf.as<double>();
      }
      launch_spmv(*a, a->blocks_boundary, a->n_blocks_boundary, true, x->d(), nx, halo, k, y->d(), ny, k, alpha, beta, c.comm);
      if (p2p) plan_p2p_ack(*p);
      cudaEvent_t ev_boundary = g_host_events[2 * nic];
      JP_CUDA(cudaEventRecord(ev_boundary, c.comm));
      const char* iblocks = static_cast<const char*>(a->blocks_interior.p);
      int uploaded_chunk = -1;
      for (int ch = 0; ch < nic; ++ch) {
        const int32_t need = a->ichunk_maxcol[ch];
        int want = 0;
        while (want < nic - 1 && xr[want + 1] <= need) ++want;
        if (want > uploaded_chunk) {
          JP_CUDA(cudaStreamWaitEvent(comp, g_host_events[want], 0));
          uploaded_chunk = want;
        }
        const int32_t b0 = a->ichunk_block[ch], b1 = a->ichunk_block[ch + 1];
        launch_spmv_blocks(*a, iblocks + (size_t)b0 * kBlockDescBytes, b1 - b0, false, x->d(), nx, nullptr, 1, y->d(), ny, k, alpha, beta, comp);
        JP_CUDA(cudaEventRecord(g_host_events[nic + ch], comp));
      }
      // the chunks without boundary rows go down first (the boundary rows finish last: they need the whole of X and the halo)
      for (int pass = 0; pass < 2; ++pass)
        for (int ch = 0; ch < nic; ++ch) {
          if ((a->ichunk_has_boundary[ch] != 0) != (pass == 1)) continue;
          JP_CUDA(cudaStreamWaitEvent(d2h, g_host_events[nic + ch], 0));
          if (pass == 1) JP_CUDA(cudaStreamWaitEvent(d2h, ev_boundary, 0));
          const int32_t r0 = a->ichunk_row[ch], r1 = a->ichunk_row[ch + 1];
          if (r1 > r0)
            JP_CUDA(cudaMemcpy2DAsync(y_host + r0, (size_t)ny * 8, y->d() + r0, (size_t)ny * 8, (size_t)(r1 - r0) * 8, (size_t)k,
                                      cudaMemcpyDeviceToHost, d2h));
        }
      JP_CUDA(cudaStreamSynchronize(d2h));
      JP_CUDA(cudaStreamSynchronize(comp));
      JP_CUDA(cudaStreamSynchronize(h2d));
      JP_CUDA(cudaEventRecord(c.ev_tmp, c.comm));  // later compute-stream work is ordered after the halo traffic of this call
      JP_CUDA(cudaStreamWaitEvent(comp, c.ev_tmp, 0));
      check_p2p_health();
      return;
    }
    if (!pipelined) {
      if (xb) JP_CUDA(cudaMemcpyAsync(x->d(), x_host, xb, cudaMemcpyHostToDevice, c.compute));
      if (beta != 0.0 && yb) JP_CUDA(cudaMemcpyAsync(y->d(), y_host, yb, cudaMemcpyHostToDevice, c.compute));
      apply_impl(a, p, y, x, mode, alpha, beta);
      if (yb) JP_CUDA(cudaMemcpyAsync(y_host, y->d(), yb, cudaMemcpyDeviceToHost, c.compute));
      JP_CUDA(cudaStreamSynchronize(c.compute));
      check_p2p_health();
      return;
    }
    // Pipelined single-GPU path: PCIe is full duplex, so X goes up in row chunks on one stream, each chunk of row
    // blocks runs as soon as the columns it reads have arrived (prefix maximum of the block's column indices), and
    // its slice of Y goes down on a third stream while the next pieces of X are still uploading.
    while ((int)g_host_events.size() < 2 * nch) {
      cudaEvent_t e;
      JP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      g_host_events.push_back(e);
    }
    cudaStream_t h2d = c.comm, comp = c.compute, d2h = c.copy;
    JP_CUDA(cudaEventRecord(c.ev_tmp, comp));  // order after whatever the compute stream was doing
    JP_CUDA(cudaStreamWaitEvent(h2d, c.ev_tmp, 0));
    std::vector<int32_t> xr(nch + 1);
    for (int i = 0; i <= nch; ++i) xr[i] = (int32_t)((int64_t)nx * i / nch);
    for (int i = 0; i < nch; ++i) {
      const size_t w = (size_t)(xr[i + 1] - xr[i]) * 8;
      if (w)
        JP_CUDA(cudaMemcpy2DAsync(x->d() + xr[i], (size_t)nx * 8, x_host + xr[i], (size_t)nx * 8, w, (size_t)k, cudaMemcpyHostToDevice, h2d));
      JP_CUDA(cudaEventRecord(g_host_events[i], h2d));
    }
    const char* blocks = static_cast<const char*>(a->blocks_all.p);
    int uploaded_chunk = -1;  // highest X chunk the compute stream already waits for
    for (int ch = 0; ch < nch; ++ch) {
      const int32_t need = a->chunk_maxcol[ch];
      int want = 0;
      while (want < nch - 1 && xr[want + 1] <= need) ++want;
      if (want > uploaded_chunk) {
        JP_CUDA(cudaStreamWaitEvent(comp, g_host_events[want], 0));
        uploaded_chunk = want;
      }
      const int32_t b0 = a->chunk_block[ch], b1 = a->chunk_block[ch + 1];
      launch_spmv_blocks(*a, blocks + (size_t)b0 * kBlockDescBytes, b1 - b0, false, x->d(), nx, nullptr, 1, y->d(), ny, k, alpha, beta, comp);
      JP_CUDA(cudaEventRecord(g_host_events[nch + ch], comp));
      JP_CUDA(cudaStreamWaitEvent(d2h, g_host_events[nch + ch], 0));
      const int32_t r0 = a->chunk_row[ch], r1 = a->chunk_row[ch + 1];
      if (r1 > r0)
        JP_CUDA(cudaMemcpy2DAsync(y_host + r0, (size_t)ny * 8, y->d() + r0, (size_t)ny * 8, (size_t)(r1 - r0) * 8, (size_t)k,
                                  cudaMemcpyDeviceToHost, d2h));
    }
    JP_CUDA(cudaStreamSynchronize(d2h));
    JP_CUDA(cudaStreamSynchronize(comp));
    JP_CUDA(cudaStreamSynchronize(h2d));
  });
}

}  // extern "C"
