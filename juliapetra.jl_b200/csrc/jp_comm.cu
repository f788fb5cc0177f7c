// jp_comm.cu -- Comm backends behind the reference's abstract Comm (src/Comm.jl:12-57):
//   serial  = SerialComm  (src/SerialComm.jl:90-131), identities
//   nccl    = the replacement of MPIComm (src/MPIComm.jl:42-91): one rank per GPU, collectives over NVLink.
// The host-array collectives here serve the one-time setup (maps, directories, plan construction); the hot
// reductions (dot/norm) run on device buffers through comm_allreduce_device().
#include "jp_common.h"
#include "jp_p2p.cuh"

namespace jp {

Comm::~Comm() {
  for (int q = 0; q < (int)ar_peer_host.size(); ++q)
    if (ar_peer_host[q] && q != rank0) cudaIpcCloseMemHandle(ar_peer_host[q]);
  retire_arena(ar_arena);  // a peer may still have it mapped
  if (nccl) ncclCommDestroy(nccl);
}

static ncclDataType_t nccl_type(int32_t dtype) {
  JP_REQUIRE(dtype == JP_F64 || dtype == JP_I64, "dtype must be JP_F64 or JP_I64");
  return dtype == JP_F64 ? ncclDouble : ncclInt64;
}

static ncclRedOp_t nccl_op(int32_t op) {
  JP_REQUIRE(op == JP_SUM || op == JP_MAX || op == JP_MIN, "op must be JP_SUM, JP_MAX or JP_MIN");
  return op == JP_SUM ? ncclSum : op == JP_MAX ? ncclMax : ncclMin;
}

// all-gather of `bytes` host bytes per rank (setup only)
void allgather_host_bytes(Comm& c, const void* in, size_t bytes, void* out) {
  Context& x = ctx();
  char* buf = (char*)scratch(bytes * (c.nproc + 1));
  JP_CUDA(cudaMemcpyAsync(buf, in, bytes, cudaMemcpyHostToDevice, x.comm));
  JP_NCCL(ncclAllGather(buf, buf + bytes, bytes, ncclChar, c.nccl, x.comm));
  JP_CUDA(cudaMemcpyAsync(out, buf + bytes, bytes * c.nproc, cudaMemcpyDeviceToHost, x.comm));
  JP_CUDA(cudaStreamSynchronize(x.comm));
}

// min over the ranks of a host integer (setup only): "everybody or nobody" decisions
int64_t allreduce_min_host(Comm& c, int64_t v) {
  Context& x = ctx();
  int64_t out = 0;
  int64_t* buf = (int64_t*)scratch(16);
  JP_CUDA(cudaMemcpyAsync(buf, &v, 8, cudaMemcpyHostToDevice, x.comm));
  JP_NCCL(ncclAllReduce(buf, buf, 1, ncclInt64, ncclMin, c.nccl, x.comm));
  JP_CUDA(cudaMemcpyAsync(&out, buf, 8, cudaMemcpyDeviceToHost, x.comm));
  JP_CUDA(cudaStreamSynchronize(x.comm));
  return out;
}


// ---------------------------------------------------------------------------------------------------------------
// Small allreduce over peer memory (the dot / norm scalars of src/MultiVector.jl:107,132: k doubles per call).
// Arena of rank r:  flags[q] (uint64, q < nproc) = last epoch rank q has stored into r's slots
//                   slots[parity][q][kSmallAllreduceMax] doubles
// Call number e (identical on every rank: collectives are ordered) uses parity e & 1.  A rank that is in call e + 2 has
// seen every peer's flag of call e + 1, and a peer stores its values of call e + 1 only after it has finished reading
// call e -- so the slots of parity e & 1 are free again: two parities suffice, no ack is needed.
// The sum is taken in rank order on every rank: all ranks get bit-identical results (like MPI_Allreduce on one node).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kSmallAllreduceMax = 64;
constexpr int kSmallAllreduceRanks = 64;
constexpr size_t kArSlotsOffset = 4096;

__global__ void __launch_bounds__(256) allreduce_p2p_kernel(double* __restrict__ inout, int k, int P, int me, char* const* __restrict__ peers,
                                                            uint64_t epoch, HostSink sink, unsigned long long* timeout_flag) {
  const int tid = threadIdx.x;
  const size_t par_off = kArSlotsOffset + (size_t)(epoch & 1) * P * kSmallAllreduceMax * 8;
  for (int idx = tid; idx < P * k; idx += blockDim.x) {
    const int q = idx / k, j = idx - q * k;
    reinterpret_cast<double*>(peers[q] + par_off)[(size_t)me * kSmallAllreduceMax + j] = inout[j];
  }
  __threadfence_system();
  __syncthreads();
  if (tid < P) {
    st_release_sys(reinterpret_cast<uint64_t*>(peers[tid]) + me, epoch);
    wait_flag_ge(reinterpret_cast<const uint64_t*>(peers[me]) + tid, epoch, timeout_flag);
  }
  __syncthreads();
  if (tid < k) {
    const volatile double* slots = reinterpret_cast<const volatile double*>(peers[me] + par_off);
    double s = 0.0;
    for (int q = 0; q < P; ++q) s += slots[(size_t)q * kSmallAllreduceMax + tid];
    inout[tid] = s;
    if (sink.out) *reinterpret_cast<volatile double*>(sink.out + tid) = s;
  }
  if (sink.out) {
    __threadfence_system();
    __syncthreads();
    if (tid == 0) *reinterpret_cast<volatile unsigned long long*>(sink.flag) = sink.seq;
  }
}

// collective, once per comm: arenas, IPC handles, mappings.  False on every rank when any rank cannot take part.
static bool small_allreduce_setup(Comm& c) {
  Context& x = ctx();
  static const bool disabled = []() {
    const char* v = std::getenv("JP_ALLREDUCE");
    return v && std::string(v) == "nccl";
  }();
  if (disabled) return false;
  const int P = c.nproc, me = c.rank0;
  JP_CUDA(cudaStreamSynchronize(x.comm));
  JP_CUDA(cudaStreamSynchronize(x.compute));
  int ok = P <= kSmallAllreduceRanks ? 1 : 0;
  const size_t bytes = kArSlotsOffset + 2 * (size_t)std::min(P, kSmallAllreduceRanks) * kSmallAllreduceMax * 8;
  c.ar_arena.alloc(bytes);
  JP_CUDA(cudaMemset(c.ar_arena.p, 0, bytes));
  cudaIpcMemHandle_t mine;
  std::memset(&mine, 0, sizeof mine);
  if (cudaIpcGetMemHandle(&mine, c.ar_arena.p) != cudaSuccess) {
    cudaGetLastError();
    ok = 0;
  }
  std::vector<cudaIpcMemHandle_t> all(P);
  allgather_host_bytes(c, &mine, sizeof mine, all.data());
  c.ar_peer_host.assign(P, nullptr);
  c.ar_peer_host[me] = c.ar_arena.p;
  for (int q = 0; q < P && ok; ++q) {
    if (q == me) continue;
    void* ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, all[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      ok = 0;
      break;
    }
    c.ar_peer_host[q] = ptr;
  }
  if (!allreduce_min_host(c, ok)) {
    for (int q = 0; q < P; ++q)
      if (c.ar_peer_host[q] && q != me) cudaIpcCloseMemHandle(c.ar_peer_host[q]);
    c.ar_peer_host.clear();
    allreduce_min_host(c, 1);  // every mapping is closed before any arena is released
    c.ar_arena.release();
    return false;
  }
  c.ar_peers.alloc((size_t)P * sizeof(void*));
  JP_CUDA(cudaMemcpy(c.ar_peers.p, c.ar_peer_host.data(), (size_t)P * sizeof(void*), cudaMemcpyHostToDevice));
  c.ar_epoch = 0;
  allreduce_min_host(c, 1);  // nobody stores into an arena that is not mapped everywhere yet
  return true;
}

// sumAll of a device buffer in place (MPI.allreduce(+), src/MPIComm.jl:57-59)
bool comm_allreduce_device(Comm& c, double* buf, int64_t n, cudaStream_t stream, const HostSink* sink) {
  if (c.serial || n == 0) return false;  // src/SerialComm.jl:105-107
  if (n <= kSmallAllreduceMax) {
    if (c.ar_state == 0) c.ar_state = small_allreduce_setup(c) ? 1 : -1;
    if (c.ar_state == 1) {
      allreduce_p2p_kernel<<<1, 256, 0, stream>>>(buf, (int)n, c.nproc, c.rank0, c.ar_peers.as<char*>(), ++c.ar_epoch,
                                                   sink ? *sink : HostSink(), p2p_timeout_flag_device());
      JP_LAUNCH_CHECK();
      return sink != nullptr;
    }
  }
  JP_NCCL(ncclAllReduce(buf, buf, (size_t)n, ncclDouble, ncclSum, c.nccl, stream));
  return false;
}

}  // namespace jp

using namespace jp;

extern "C" {

int32_t jp_comm_unique_id(uint8_t* id128) {
  return guarded([&] {
    JP_REQUIRE(id128 != nullptr, "jp_comm_unique_id: null buffer");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    ncclUniqueId id;
    JP_NCCL(ncclGetUniqueId(&id));
    std::memcpy(id128, &id, 128);
  });
}

int32_t jp_comm_create_serial(jp_handle* comm) {
  return guarded([&] {
    require_init();
    JP_REQUIRE(comm != nullptr, "jp_comm_create_serial: null output");
    auto c = std::make_unique<Comm>();
    *comm = register_object(std::move(c));
  });
}

int32_t jp_comm_create_nccl(const uint8_t* id128, int32_t nranks, int32_t rank0, jp_handle* comm) {
  return guarded([&] {
    require_init();
    JP_REQUIRE(comm != nullptr && id128 != nullptr, "jp_comm_create_nccl: null argument");
    JP_REQUIRE(nranks >= 1 && rank0 >= 0 && rank0 < nranks, "jp_comm_create_nccl: rank out of range");
    auto c = std::make_unique<Comm>();
    c->serial = false;
    c->nproc = nranks;
    c->rank0 = rank0;
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    JP_NCCL(ncclCommInitRank(&c->nccl, nranks, id, rank0));
    *comm = register_object(std::move(c));
  });
}

int32_t jp_comm_destroy(jp_handle comm) {
  return guarded([&] { destroy_object(comm, Kind::Comm, "comm"); });
}

int32_t jp_comm_rank(jp_handle comm, int32_t* pid1, int32_t* nproc) {
  return guarded([&] {
    Comm* c = get<Comm>(comm, Kind::Comm, "comm");
    if (pid1) *pid1 = c->rank0 + 1;
    if (nproc) *nproc = c->nproc;
  });
}

int32_t jp_comm_barrier(jp_handle comm) {
  return guarded([&] {
    require_init();
    Comm* c = get<Comm>(comm, Kind::Comm, "comm");
    Context& x = ctx();
    JP_CUDA(cudaStreamSynchronize(x.compute));
    if (!c->serial) {
      double* buf = (double*)scratch(8);
      JP_CUDA(cudaMemsetAsync(buf, 0, 8, x.comm));
      JP_NCCL(ncclAllReduce(buf, buf, 1, ncclDouble, ncclSum, c->nccl, x.comm));
    }
    JP_CUDA(cudaStreamSynchronize(x.comm));
  });
}

int32_t jp_comm_allreduce(jp_handle comm, int32_t op, int32_t dtype, const void* in, void* out, int64_t n) {
  return guarded([&] {
    require_init();
    Comm* c = get<Comm>(comm, Kind::Comm, "comm");
    JP_REQUIRE(n >= 0 && (n == 0 || (in && out)), "jp_comm_allreduce: bad arguments");
    ncclDataType_t t = nccl_type(dtype);
    ncclRedOp_t o = nccl_op(op);
    if (n == 0) return;
    if (c->serial) {
      if (out != in) std::memmove(out, in, (size_t)n * 8);
      return;
    }
    Context& x = ctx();
    void* buf = scratch((size_t)n * 8);
    JP_CUDA(cudaMemcpyAsync(buf, in, (size_t)n * 8, cudaMemcpyHostToDevice, x.comm));
    JP_NCCL(ncclAllReduce(buf, buf, (size_t)n, t, o, c->nccl, x.comm));
    JP_CUDA(cudaMemcpyAsync(out, buf, (size_t)n * 8, cudaMemcpyDeviceToHost, x.comm));
    JP_CUDA(cudaStreamSynchronize(x.comm));
  });
}

int32_t jp_comm_scan_sum(jp_handle comm, int32_t dtype, const void* in, void* out, int64_t n) {
  return guarded([&] {
    require_init();
    Comm* c = get<Comm>(comm, Kind::Comm, "comm");
    JP_REQUIRE(n >= 0 && (n == 0 || (in && out)), "jp_comm_scan_sum: bad arguments");
    ncclDataType_t t = nccl_type(dtype);
    if (n == 0) return;
    if (c->serial) {
      if (out != in) std::memmove(out, in, (size_t)n * 8);
      return;
    }
    // MPI.Scan(SUM) (src/MPIComm.jl:77-79): all-gather (P <= 8 ranks) + prefix on the host, ascending rank order
    Context& x = ctx();
    size_t bytes = (size_t)n * 8;
    char* buf = (char*)scratch(bytes * (c->nproc + 1));
    JP_CUDA(cudaMemcpyAsync(buf, in, bytes, cudaMemcpyHostToDevice, x.comm));
    JP_NCCL(ncclAllGather(buf, buf + bytes, (size_t)n, t, c->nccl, x.comm));
    std::vector<char> all(bytes * c->nproc);
    JP_CUDA(cudaMemcpyAsync(all.data(), buf + bytes, bytes * c->nproc, cudaMemcpyDeviceToHost, x.comm));
    JP_CUDA(cudaStreamSynchronize(x.comm));
    if (dtype == JP_F64) {
      double* o = (double*)out;
      const double* a = (const double*)all.data();
      for (int64_t i = 0; i < n; ++i) {
        double s = a[i];
        for (int p = 1; p <= c->rank0; ++p) s = s + a[(size_t)p * n + i];
        o[i] = s;
      }
    } else {
      int64_t* o = (int64_t*)out;
      const int64_t* a = (const int64_t*)all.data();
      for (int64_t i = 0; i < n; ++i) {
        int64_t s = a[i];
        for (int p = 1; p <= c->rank0; ++p) s += a[(size_t)p * n + i];
        o[i] = s;
      }
    }
  });
}

int32_t jp_comm_gather_all(jp_handle comm, int32_t dtype, const void* in, int64_t n, const int64_t* counts, void* out) {
  return guarded([&] {
    require_init();
    Comm* c = get<Comm>(comm, Kind::Comm, "comm");
    ncclDataType_t t = nccl_type(dtype);
    JP_REQUIRE(n >= 0 && counts != nullptr, "jp_comm_gather_all: bad arguments");
    JP_REQUIRE(counts[c->rank0] == n, "jp_comm_gather_all: counts[myPid] must equal n");
    if (c->serial) {
      if (n > 0 && out != in) std::memmove(out, in, (size_t)n * 8);
      return;
    }
    // MPI.Allgatherv (src/MPIComm.jl:52-55): pad every rank's block to the longest, all-gather, compact
    int64_t maxc = 0, total = 0;
    for (int p = 0; p < c->nproc; ++p) {
      JP_REQUIRE(counts[p] >= 0, "jp_comm_gather_all: negative count");
      maxc = counts[p] > maxc ? counts[p] : maxc;
      total += counts[p];
    }
    if (maxc == 0) return;
    Context& x = ctx();
    size_t blk = (size_t)maxc * 8;
    char* buf = (char*)scratch(blk * (c->nproc + 1));
    if (n > 0) JP_CUDA(cudaMemcpyAsync(buf, in, (size_t)n * 8, cudaMemcpyHostToDevice, x.comm));
    JP_NCCL(ncclAllGather(buf, buf + blk, (size_t)maxc, t, c->nccl, x.comm));
    char* o = (char*)out;
    size_t pos = 0;
    for (int p = 0; p < c->nproc; ++p) {
      if (counts[p] > 0)
        JP_CUDA(cudaMemcpyAsync(o + pos, buf + blk * (p + 1), (size_t)counts[p] * 8, cudaMemcpyDeviceToHost, x.comm));
      pos += (size_t)counts[p] * 8;
    }
    JP_CUDA(cudaStreamSynchronize(x.comm));
    (void)total;
  });
}

int32_t jp_comm_broadcast(jp_handle comm, int32_t dtype, void* inout, int64_t n, int32_t root1) {
  return guarded([&] {
    require_init();
    Comm* c = get<Comm>(comm, Kind::Comm, "comm");
    ncclDataType_t t = nccl_type(dtype);
    JP_REQUIRE(root1 >= 1 && root1 <= c->nproc,
               c->serial ? "SerialComm can only accept PID of 1" : "jp_comm_broadcast: root PID out of range");
    JP_REQUIRE(n >= 0 && (n == 0 || inout), "jp_comm_broadcast: bad arguments");
    if (c->serial || n == 0) return;
    Context& x = ctx();
    void* buf = scratch((size_t)n * 8);
    if (c->rank0 == root1 - 1) JP_CUDA(cudaMemcpyAsync(buf, inout, (size_t)n * 8, cudaMemcpyHostToDevice, x.comm));
    JP_NCCL(ncclBroadcast(buf, buf, (size_t)n, t, root1 - 1, c->nccl, x.comm));
    JP_CUDA(cudaMemcpyAsync(inout, buf, (size_t)n * 8, cudaMemcpyDeviceToHost, x.comm));
    JP_CUDA(cudaStreamSynchronize(x.comm));
  });
}

int32_t jp_comm_alltoallv(jp_handle comm, int32_t elem_bytes, const void* send, const int64_t* sendcounts, void* recv,
                          const int64_t* recvcounts) {
  return guarded([&] {
    require_init();
    Comm* c = get<Comm>(comm, Kind::Comm, "comm");
    JP_REQUIRE(elem_bytes > 0 && sendcounts && recvcounts, "jp_comm_alltoallv: bad arguments");
    int P = c->nproc;
    std::vector<size_t> soff(P + 1, 0), roff(P + 1, 0);
    for (int p = 0; p < P; ++p) {
      JP_REQUIRE(sendcounts[p] >= 0 && recvcounts[p] >= 0, "jp_comm_alltoallv: negative count");
      soff[p + 1] = soff[p] + (size_t)sendcounts[p] * elem_bytes;
      roff[p + 1] = roff[p] + (size_t)recvcounts[p] * elem_bytes;
    }
    if (c->serial) {
      JP_REQUIRE(sendcounts[0] == recvcounts[0], "jp_comm_alltoallv: self counts differ");
      if (soff[1] > 0) std::memmove(recv, send, soff[1]);
      return;
    }
    Context& x = ctx();
    size_t sa = (soff[P] + 255) & ~(size_t)255;
    char* sbuf = (char*)scratch(sa + roff[P] + 256);
    char* rbuf = sbuf + sa;
    if (soff[P] > 0) JP_CUDA(cudaMemcpyAsync(sbuf, send, soff[P], cudaMemcpyHostToDevice, x.comm));
    JP_NCCL(ncclGroupStart());
    for (int p = 0; p < P; ++p) {
      if (p == c->rank0) continue;
      if (recvcounts[p] > 0) JP_NCCL(ncclRecv(rbuf + roff[p], roff[p + 1] - roff[p], ncclChar, p, c->nccl, x.comm));
      if (sendcounts[p] > 0) JP_NCCL(ncclSend(sbuf + soff[p], soff[p + 1] - soff[p], ncclChar, p, c->nccl, x.comm));
    }
    JP_NCCL(ncclGroupEnd());
    int me = c->rank0;
    JP_REQUIRE(sendcounts[me] == recvcounts[me], "jp_comm_alltoallv: self counts differ");
    if (sendcounts[me] > 0)
      JP_CUDA(cudaMemcpyAsync(rbuf + roff[me], sbuf + soff[me], soff[me + 1] - soff[me], cudaMemcpyDeviceToDevice, x.comm));
    if (roff[P] > 0) JP_CUDA(cudaMemcpyAsync(recv, rbuf, roff[P], cudaMemcpyDeviceToHost, x.comm));
    JP_CUDA(cudaStreamSynchronize(x.comm));
  });
}

}  // extern "C"
