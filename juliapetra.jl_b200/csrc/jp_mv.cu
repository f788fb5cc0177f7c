// jp_mv.cu -- DenseMultiVector storage (src/DenseMultiVector.jl:6-60) and the MultiVector kernels
// (src/MultiVector.jl:52-161): fill!, scale!, axpby (broadcast copyto!), dot, norm, commReduce.
//
// All of these are HBM-bound streaming kernels: coalesced 8-byte loads with 4-way unrolled independent
// accumulators, one launch per call.  dot/norm are fused block reductions: every block writes one partial per
// vector, the last block to finish (ticket counter) adds the partials in block order, so the result is
// deterministic for a given launch geometry; the cross-rank sum is one ncclAllReduce of k doubles.
// Elementwise arithmetic uses explicit non-fused __dmul_rn/__dadd_rn so results are bit-identical to the
// reference's scalar loops (Julia does not contract a*x+b*y into an FMA).
#include "jp_common.h"
#include "jp_mv_kernels.cuh"

namespace jp {

namespace {
using namespace mvk;

constexpr int kMaxReduceBlocks = 1024;  // partials per vector

int grid_for(int64_t n, int per_thread) {
  int64_t want = (n + (int64_t)kThreads * per_thread - 1) / ((int64_t)kThreads * per_thread);
  int64_t cap = (int64_t)ctx().sm_count * 8;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}

struct ReduceWorkspace {
  double* partials;      // [k][blocks]
  unsigned int* tickets; // [k]
  double* out;           // [k]
};

// carve the reduction workspace out of a dedicated device buffer (tickets must stay zero between calls)
DeviceBuffer g_reduce_buf;
int32_t g_reduce_k = 0;

ReduceWorkspace reduce_workspace(int32_t k) {
  if (k > g_reduce_k) {
    JP_CUDA(cudaStreamSynchronize(ctx().compute));
    int32_t kk = k < 32 ? 32 : k;
    size_t bytes = (size_t)kk * kMaxReduceBlocks * 8 + (size_t)kk * 8 + (size_t)kk * 8;
    g_reduce_buf.alloc(bytes);
    JP_CUDA(cudaMemsetAsync(g_reduce_buf.p, 0, bytes, ctx().compute));
    g_reduce_k = kk;
  }
  ReduceWorkspace w;
  char* base = (char*)g_reduce_buf.p;
  w.partials = (double*)base;
  w.out = (double*)(base + (size_t)g_reduce_k * kMaxReduceBlocks * 8);
  w.tickets = (unsigned int*)(base + (size_t)g_reduce_k * kMaxReduceBlocks * 8 + (size_t)g_reduce_k * 8);
  return w;
}

// local sums for every vector -> w.out (device; or `out` when given), stream-ordered on compute
template <int MODE>
double* reduce_local(const MultiVector& x, const MultiVector* y, double p, double* out = nullptr) {
  ReduceWorkspace w = reduce_workspace(x.k);
  if (out) w.out = out;
  int blocks = grid_for(x.n, 8);
  if (blocks > kMaxReduceBlocks) blocks = kMaxReduceBlocks;
  dim3 grid(blocks, x.k);
  reduce_kernel<MODE><<<grid, kThreads, 0, ctx().compute>>>(x.d(), y ? y->d() : nullptr, x.n, p, w.partials, w.tickets, w.out);
  JP_LAUNCH_CHECK();
  return w.out;
}

// sumAll of the k local sums + hand-over to the host (the only synchronising step).  The results travel through
// host-mapped pinned memory: publish_results_kernel stores them and then a sequence number, the host spins on the
// sequence number (checking the stream every few thousand polls so that a failed kernel surfaces as an error).
constexpr int kMappedResults = 1024;  // doubles in the mapped result block; wider multivectors take memcpy + synchronise

struct MappedResults {
  double* host = nullptr;
  double* dev = nullptr;
  unsigned long long* flag_host = nullptr;
  unsigned long long* flag_dev = nullptr;
  unsigned long long seq = 0;
};
MappedResults g_mapped;

MappedResults& mapped_results() {
  if (!g_mapped.host) {
    void* h = nullptr;
    JP_CUDA(cudaHostAlloc(&h, (size_t)kMappedResults * 8 + 64, cudaHostAllocMapped));
    std::memset(h, 0, (size_t)kMappedResults * 8 + 64);
    g_mapped.host = static_cast<double*>(h);
    g_mapped.flag_host = reinterpret_cast<unsigned long long*>(g_mapped.host + kMappedResults);
    void* d = nullptr;
    JP_CUDA(cudaHostGetDevicePointer(&d, h, 0));
    g_mapped.dev = static_cast<double*>(d);
    g_mapped.flag_dev = reinterpret_cast<unsigned long long*>(g_mapped.dev + kMappedResults);
  }
  return g_mapped;
}

void finish_reduction(Comm& c, double* dev_out, int32_t k, double* host_out) {
  Context& x = ctx();
  if (k > kMappedResults) {
    comm_allreduce_device(c, dev_out, k, x.compute);
    double* pinned = (double*)host_scratch((size_t)k * 8);
    JP_CUDA(cudaMemcpyAsync(pinned, dev_out, (size_t)k * 8, cudaMemcpyDeviceToHost, x.compute));
    JP_CUDA(cudaStreamSynchronize(x.compute));
    check_p2p_health();
    std::memcpy(host_out, pinned, (size_t)k * 8);
    return;
  }
  MappedResults& m = mapped_results();
  const unsigned long long seq = ++m.seq;
  HostSink sink;
  sink.out = m.dev;
  sink.flag = m.flag_dev;
  sink.seq = seq;
  if (!comm_allreduce_device(c, dev_out, k, x.compute, &sink)) {  // the peer-memory allreduce kernel hands over by itself
    publish_results_kernel<<<1, 32, 0, x.compute>>>(dev_out, k, m.dev, m.flag_dev, seq);
    JP_LAUNCH_CHECK();
  }
  volatile unsigned long long* flag = m.flag_host;
  for (unsigned spins = 1; *flag != seq; ++spins) {
    if ((spins & 0xfff) == 0) {
      const cudaError_t e = cudaStreamQuery(x.compute);
      if (e != cudaSuccess && e != cudaErrorNotReady) JP_CUDA(e);
      if (e == cudaSuccess && *flag != seq) {  // the stream has drained: the store is at most a fence away
        JP_CUDA(cudaStreamSynchronize(x.compute));
        if (*flag != seq) fail(JP_CUDA_ERROR, "reduction results did not reach the host");
      }
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  check_p2p_health();
  for (int32_t v = 0; v < k; ++v) host_out[v] = reinterpret_cast<volatile double*>(m.host)[v];
}

}  // namespace

double* mv_dot_local(const MultiVector& x, const MultiVector& y, double* out) { return reduce_local<0>(x, &y, 0.0, out); }
void mv_finish_reduction(Comm& c, double* dev_out, int32_t k, double* host_out) { finish_reduction(c, dev_out, k, host_out); }

void release_reduce_workspace() {
  g_reduce_buf.release();
  g_reduce_k = 0;
  if (g_mapped.host) cudaFreeHost(g_mapped.host);
  g_mapped = MappedResults();
}

void launch_scale_fill(double* y, int64_t n, double beta, cudaStream_t stream) {
  if (n == 0) return;
  if (beta == 0.0) {
    JP_CUDA(cudaMemsetAsync(y, 0, (size_t)n * 8, stream));
  } else if (beta != 1.0) {
    scale_kernel<<<grid_for(n, 4), kThreads, 0, stream>>>(y, n, beta);
    JP_LAUNCH_CHECK();
  }
}

}  // namespace jp

using namespace jp;

extern "C" {

int32_t jp_mv_create(int32_t n_local, int32_t k, jp_handle* out) {
  return guarded([&] {
    require_init();
    JP_REQUIRE(out != nullptr, "jp_mv_create: null output");
    JP_REQUIRE(n_local >= 0 && k >= 0, "jp_mv_create: negative shape");
    auto mv = std::make_unique<MultiVector>();
    mv->n = n_local;
    mv->k = k;
    size_t bytes = (size_t)n_local * k * 8;
    mv->data.alloc(bytes + 64);  // slack: vectorised tails may over-read, never over-write
    JP_CUDA(cudaMemsetAsync(mv->data.p, 0, bytes + 64, ctx().compute));
    *out = register_object(std::move(mv));
  });
}

int32_t jp_mv_destroy(jp_handle mv) {
  return guarded([&] { destroy_object(mv, Kind::MultiVector, "multivector"); });
}

int32_t jp_mv_shape(jp_handle mv, int32_t* n_local, int32_t* k) {
  return guarded([&] {
    MultiVector* m = get<MultiVector>(mv, Kind::MultiVector, "multivector");
    if (n_local) *n_local = m->n;
    if (k) *k = m->k;
  });
}

int32_t jp_mv_upload(jp_handle mv, const double* host) {
  return guarded([&] {
    require_init();
    MultiVector* m = get<MultiVector>(mv, Kind::MultiVector, "multivector");
    size_t bytes = (size_t)m->n * m->k * 8;
    if (bytes == 0) return;
    JP_REQUIRE(host != nullptr, "jp_mv_upload: null host pointer");
    JP_CUDA(cudaMemcpyAsync(m->d(), host, bytes, cudaMemcpyHostToDevice, ctx().compute));
    JP_CUDA(cudaStreamSynchronize(ctx().compute));  // the host array is not retained
  });
}

int32_t jp_mv_download(jp_handle mv, double* host) {
  return guarded([&] {
    require_init();
    MultiVector* m = get<MultiVector>(mv, Kind::MultiVector, "multivector");
    size_t bytes = (size_t)m->n * m->k * 8;
    if (bytes == 0) return;
    JP_REQUIRE(host != nullptr, "jp_mv_download: null host pointer");
    JP_CUDA(cudaStreamSynchronize(ctx().comm));
    JP_CUDA(cudaMemcpyAsync(host, m->d(), bytes, cudaMemcpyDeviceToHost, ctx().compute));
    JP_CUDA(cudaStreamSynchronize(ctx().compute));
    check_p2p_health();
  });
}

int32_t jp_mv_copy(jp_handle dst, jp_handle src) {
  return guarded([&] {
    require_init();
    MultiVector* d = get<MultiVector>(dst, Kind::MultiVector, "multivector");
    MultiVector* s = get<MultiVector>(src, Kind::MultiVector, "multivector");
    JP_REQUIRE(d->n == s->n && d->k == s->k, "jp_mv_copy: shapes differ");
    size_t bytes = (size_t)d->n * d->k * 8;
    if (bytes && d != s) JP_CUDA(cudaMemcpyAsync(d->d(), s->d(), bytes, cudaMemcpyDeviceToDevice, ctx().compute));
  });
}

int32_t jp_mv_fill(jp_handle mv, double value) {
  return guarded([&] {
    require_init();
    MultiVector* m = get<MultiVector>(mv, Kind::MultiVector, "multivector");
    int64_t n = (int64_t)m->n * m->k;
    if (n == 0) return;
    if (value == 0.0 && !std::signbit(value)) {
      JP_CUDA(cudaMemsetAsync(m->d(), 0, (size_t)n * 8, ctx().compute));
    } else {
      fill_kernel<<<grid_for(n, 4), kThreads, 0, ctx().compute>>>(m->d(), n, value);
      JP_LAUNCH_CHECK();
    }
  });
}

int32_t jp_mv_scale(jp_handle mv, double alpha) {
  return guarded([&] {
    require_init();
    MultiVector* m = get<MultiVector>(mv, Kind::MultiVector, "multivector");
    int64_t n = (int64_t)m->n * m->k;
    if (n == 0) return;
    scale_kernel<<<grid_for(n, 4), kThreads, 0, ctx().compute>>>(m->d(), n, alpha);
    JP_LAUNCH_CHECK();
  });
}

int32_t jp_mv_scale_vec(jp_handle mv, const double* alpha_k) {
  return guarded([&] {
    require_init();
    MultiVector* m = get<MultiVector>(mv, Kind::MultiVector, "multivector");
    if (m->n == 0 || m->k == 0) return;
    JP_REQUIRE(alpha_k != nullptr, "jp_mv_scale_vec: null alpha");
    double* pinned = (double*)host_scratch((size_t)m->k * 8);
    std::memcpy(pinned, alpha_k, (size_t)m->k * 8);
    double* dev = (double*)scratch((size_t)m->k * 8);
    JP_CUDA(cudaMemcpyAsync(dev, pinned, (size_t)m->k * 8, cudaMemcpyHostToDevice, ctx().compute));
    dim3 grid(grid_for(m->n, 4), m->k);
    scale_vec_kernel<<<grid, kThreads, 0, ctx().compute>>>(m->d(), m->n, dev);
    JP_LAUNCH_CHECK();
    JP_CUDA(cudaStreamSynchronize(ctx().compute));  // pinned/scratch staging is shared
  });
}

int32_t jp_mv_axpby(jp_handle y, double a, jp_handle x, double b) {
  return guarded([&] {
    require_init();
    MultiVector* my = get<MultiVector>(y, Kind::MultiVector, "multivector");
    MultiVector* mx = get<MultiVector>(x, Kind::MultiVector, "multivector");
    JP_REQUIRE(my->n == mx->n && my->k == mx->k, "jp_mv_axpby: shapes differ");
    int64_t n = (int64_t)my->n * my->k;
    if (n == 0) return;
    axpby_kernel<<<grid_for(n, 4), kThreads, 0, ctx().compute>>>(my->d(), mx->d(), n, a, b);
    JP_LAUNCH_CHECK();
  });
}

int32_t jp_mv_axpby_norm2(jp_handle comm, jp_handle y, double a, jp_handle x, double b, double* out_k) {
  return guarded([&] {
    require_init();
    Comm* c = get<Comm>(comm, Kind::Comm, "comm");
    MultiVector* my = get<MultiVector>(y, Kind::MultiVector, "multivector");
    MultiVector* mx = get<MultiVector>(x, Kind::MultiVector, "multivector");
    JP_REQUIRE(my->n == mx->n && my->k == mx->k, "jp_mv_axpby_norm2: shapes differ");
    JP_REQUIRE(my != mx, "jp_mv_axpby_norm2: x and y must be different multivectors");
    if (my->k == 0) return;
    JP_REQUIRE(out_k != nullptr, "jp_mv_axpby_norm2: null output");
    ReduceWorkspace w = reduce_workspace(my->k);
    int blocks = grid_for(my->n, 8);
    if (blocks > kMaxReduceBlocks) blocks = kMaxReduceBlocks;
    dim3 grid(blocks, my->k);
    axpby_norm2_kernel<<<grid, kThreads, 0, ctx().compute>>>(my->d(), mx->d(), my->n, a, b, w.partials, w.tickets, w.out);
    JP_LAUNCH_CHECK();
    finish_reduction(*c, w.out, my->k, out_k);
    for (int v = 0; v < my->k; ++v) out_k[v] = std::sqrt(out_k[v]);  // sqrt after the allreduce, src/MultiVector.jl:132-133
  });
}

int32_t jp_mv_dot(jp_handle comm, jp_handle x, jp_handle y, double* out_k) {
  return guarded([&] {
    require_init();
    Comm* c = get<Comm>(comm, Kind::Comm, "comm");
    MultiVector* mx = get<MultiVector>(x, Kind::MultiVector, "multivector");
    MultiVector* my = get<MultiVector>(y, Kind::MultiVector, "multivector");
    // src/MultiVector.jl:88-93
    JP_REQUIRE(mx->k == my->k, "MultiVectors must have the same number of vectors to take the dot product of them");
    JP_REQUIRE(mx->n == my->n, "Vectors must have the same length to take the dot product of them");
    if (mx->k == 0) return;
    JP_REQUIRE(out_k != nullptr, "jp_mv_dot: null output");
    double* dev = reduce_local<0>(*mx, my, 0.0);
    finish_reduction(*c, dev, mx->k, out_k);
  });
}

int32_t jp_mv_norm2(jp_handle comm, jp_handle x, double* out_k) {
  return guarded([&] {
    require_init();
    Comm* c = get<Comm>(comm, Kind::Comm, "comm");
    MultiVector* mx = get<MultiVector>(x, Kind::MultiVector, "multivector");
    if (mx->k == 0) return;
    JP_REQUIRE(out_k != nullptr, "jp_mv_norm2: null output");
    double* dev = reduce_local<1>(*mx, nullptr, 2.0);
    finish_reduction(*c, dev, mx->k, out_k);
    for (int v = 0; v < mx->k; ++v) out_k[v] = std::sqrt(out_k[v]);  // sqrt after the allreduce, src/MultiVector.jl:132-133
  });
}

int32_t jp_mv_normp(jp_handle comm, jp_handle x, double p, double* out_k) {
  return guarded([&] {
    require_init();
    if (p == 2.0) {
      int32_t rc = jp_mv_norm2(comm, x, out_k);
      if (rc != JP_OK) {
        char buf[512];
        jp_last_error(buf, sizeof buf);
        fail(rc, buf);
      }
      return;
    }
    Comm* c = get<Comm>(comm, Kind::Comm, "comm");
    MultiVector* mx = get<MultiVector>(x, Kind::MultiVector, "multivector");
    if (mx->k == 0) return;
    JP_REQUIRE(out_k != nullptr, "jp_mv_normp: null output");
    double* dev = reduce_local<2>(*mx, nullptr, p);
    finish_reduction(*c, dev, mx->k, out_k);
    for (int v = 0; v < mx->k; ++v) out_k[v] = std::pow(out_k[v], 1.0 / p);  // src/MultiVector.jl:143-144
  });
}

int32_t jp_mv_comm_reduce(jp_handle comm, jp_handle mv) {
  return guarded([&] {
    require_init();
    Comm* c = get<Comm>(comm, Kind::Comm, "comm");
    MultiVector* m = get<MultiVector>(mv, Kind::MultiVector, "multivector");
    // view .= sumAll(comm, view), src/MultiVector.jl:159-160 (identity on SerialComm)
    comm_allreduce_device(*c, m->d(), (int64_t)m->n * m->k, ctx().compute);
  });
}

}  // extern "C"
