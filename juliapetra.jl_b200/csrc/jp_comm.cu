// jp_comm.cu -- Comm backends behind the reference's abstract Comm (src/Comm.jl:12-57):
//   serial  = SerialComm  (src/SerialComm.jl:90-131), identities
//   nccl    = the replacement of MPIComm (src/MPIComm.jl:42-91): one rank per GPU, collectives over NVLink.
// The host-array collectives here serve the one-time setup (maps, directories, plan construction); the hot
// reductions (dot/norm) run on device buffers through comm_allreduce_device().
#include "jp_common.h"

namespace jp {

Comm::~Comm() {
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

// sumAll of a device buffer in place (MPI.allreduce(+), src/MPIComm.jl:57-59)
void comm_allreduce_device(Comm& c, double* buf, int64_t n, cudaStream_t stream) {
  if (c.serial || n == 0) return;  // src/SerialComm.jl:105-107
  JP_NCCL(ncclAllReduce(buf, buf, (size_t)n, ncclDouble, ncclSum, c.nccl, stream));
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
