// jp_solver.cu -- device-resident scalars for solver loops (BASELINE configs[4]).
//
// The reference's API returns every dot / norm to the host (src/MultiVector.jl:84-146), so a CG iteration written
// against it synchronises the host with the GPU twice, ~26 us each on one GPU plus an allreduce latency on several --
// a third of a 128^3 iteration, ~90 us of the 531 us iteration measured on 512^3 over 8 GPUs.  Here reductions leave
// their result in a small device array (a `scalars` object: jp_*_dev), the allreduce runs in place on the stream, and
// the update kernels take their coefficients as  scale * s[num] / s[den]  evaluated on the device.  An iteration is
// then pure stream work (and CUDA-graph capturable); the host looks at the residual every few iterations
// (jp_scalars_download).  The arithmetic is the one of the host-scalar calls: same kernels for the reductions, the
// same unfused update, the quotient formed once per thread instead of once on the host.
#include "jp_common.h"
#include "jp_solver_kernels.cuh"

namespace jp {

namespace {
using namespace solk;

struct Scalars : Object {
  Scalars() : Object(Kind::Scalars) {}
  int32_t n = 0;
  DeviceBuffer data;  // n doubles
  DeviceBuffer ws;    // axpby_norm2_dev workspace: [partials (1024) | ticket]
  double* d() const { return data.as<double>(); }
};

constexpr int kMaxBlocks = 1024;

int grid_for(int64_t n, int per_thread) {
  int64_t want = (n + (int64_t)kThreads * per_thread - 1) / ((int64_t)kThreads * per_thread);
  int64_t cap = (int64_t)ctx().sm_count * 8;
  if (want > cap) want = cap;
  if (want > kMaxBlocks) want = kMaxBlocks;
  if (want < 1) want = 1;
  return (int)want;
}

Scalars* scalars(jp_handle h) { return get<Scalars>(h, Kind::Scalars, "scalars"); }

void check_slot(const Scalars& s, int32_t slot, const char* what) {
  JP_REQUIRE(slot >= 0 && slot < s.n, std::string(what) + ": slot outside the scalars object");
}

Coeff make_coeff(const Scalars& s, double scale, int32_t num, int32_t den, const char* what) {
  JP_REQUIRE(num < s.n && den < s.n, std::string(what) + ": coefficient slot outside the scalars object");
  Coeff c;
  c.scale = scale;
  c.num = num;
  c.den = den;
  return c;
}

}  // namespace
}  // namespace jp

using namespace jp;

extern "C" {

int32_t jp_scalars_create(int32_t n, jp_handle* out) {
  return guarded([&] {
    require_init();
    JP_REQUIRE(out != nullptr && n >= 1 && n <= 4096, "jp_scalars_create: need 1 <= n <= 4096 and an output");
    auto s = std::make_unique<Scalars>();
    s->n = n;
    s->data.alloc((size_t)n * 8);
    s->ws.alloc((size_t)(kMaxBlocks + 2) * 8);
    JP_CUDA(cudaMemsetAsync(s->data.p, 0, (size_t)n * 8, ctx().compute));
    JP_CUDA(cudaMemsetAsync(s->ws.p, 0, (size_t)(kMaxBlocks + 2) * 8, ctx().compute));
    *out = register_object(std::move(s));
  });
}

int32_t jp_scalars_destroy(jp_handle s) {
  return guarded([&] { destroy_object(s, Kind::Scalars, "scalars"); });
}

int32_t jp_scalars_upload(jp_handle s, const double* host) {
  return guarded([&] {
    require_init();
    Scalars* sc = scalars(s);
    JP_REQUIRE(host != nullptr, "jp_scalars_upload: null host pointer");
    JP_CUDA(cudaMemcpyAsync(sc->d(), host, (size_t)sc->n * 8, cudaMemcpyHostToDevice, ctx().compute));
    JP_CUDA(cudaStreamSynchronize(ctx().compute));
  });
}

int32_t jp_scalars_download(jp_handle s, double* host) {
  return guarded([&] {
    require_init();
    Scalars* sc = scalars(s);
    JP_REQUIRE(host != nullptr, "jp_scalars_download: null host pointer");
    JP_CUDA(cudaMemcpyAsync(host, sc->d(), (size_t)sc->n * 8, cudaMemcpyDeviceToHost, ctx().compute));
    JP_CUDA(cudaStreamSynchronize(ctx().compute));
    check_p2p_health();
  });
}

int32_t jp_scalars_copy(jp_handle s, int32_t dst, int32_t src) {
  return guarded([&] {
    require_init();
    Scalars* sc = scalars(s);
    check_slot(*sc, dst, "jp_scalars_copy");
    check_slot(*sc, src, "jp_scalars_copy");
    if (dst != src) JP_CUDA(cudaMemcpyAsync(sc->d() + dst, sc->d() + src, 8, cudaMemcpyDeviceToDevice, ctx().compute));
  });
}

int32_t jp_mv_dot_dev(jp_handle comm, jp_handle x, jp_handle y, jp_handle s, int32_t slot) {
  return guarded([&] {
    require_init();
    Comm* c = get<Comm>(comm, Kind::Comm, "comm");
    MultiVector* mx = get<MultiVector>(x, Kind::MultiVector, "multivector");
    MultiVector* my = get<MultiVector>(y, Kind::MultiVector, "multivector");
    Scalars* sc = scalars(s);
    check_slot(*sc, slot, "jp_mv_dot_dev");
    JP_REQUIRE(mx->k == 1 && my->k == 1 && mx->n == my->n, "jp_mv_dot_dev: two single vectors of the same length");
    mv_dot_local(*mx, *my, sc->d() + slot);  // the host-scalar path's kernel, its result written straight into the slot
    comm_allreduce_device(*c, sc->d() + slot, 1, ctx().compute);
  });
}

int32_t jp_apply_dot_dev(jp_handle comm, jp_handle Y, jp_handle A, jp_handle X, jp_handle plan, double alpha, double beta, jp_handle W,
                         jp_handle s, int32_t slot) {
  return guarded([&] {
    require_init();
    Comm* c = get<Comm>(comm, Kind::Comm, "comm");
    Csr* a = get<Csr>(A, Kind::Csr, "csr matrix");
    MultiVector* y = get<MultiVector>(Y, Kind::MultiVector, "multivector");
    MultiVector* x = get<MultiVector>(X, Kind::MultiVector, "multivector");
    MultiVector* w = get<MultiVector>(W, Kind::MultiVector, "multivector");
    Scalars* sc = scalars(s);
    check_slot(*sc, slot, "jp_apply_dot_dev");
    JP_REQUIRE(x != y && x->k == 1 && y->k == 1 && w->k == 1 && w->n == y->n, "jp_apply_dot_dev: single vectors, X != Y, W shaped like Y");
    Plan* p = plan ? get<Plan>(plan, Kind::Plan, "plan") : nullptr;
    double* dev = apply_dot_local(a, p, y, x, w, JP_NO_TRANS, alpha, beta);  // one kernel with or without a peer-memory halo
    JP_CUDA(cudaMemcpyAsync(sc->d() + slot, dev, 8, cudaMemcpyDeviceToDevice, ctx().compute));
    comm_allreduce_device(*c, sc->d() + slot, 1, ctx().compute);
  });
}

int32_t jp_mv_axpby_dev(jp_handle y, jp_handle x, jp_handle s, double a_scale, int32_t a_num, int32_t a_den, double b_scale, int32_t b_num,
                        int32_t b_den) {
  return guarded([&] {
    require_init();
    MultiVector* my = get<MultiVector>(y, Kind::MultiVector, "multivector");
    MultiVector* mx = get<MultiVector>(x, Kind::MultiVector, "multivector");
    Scalars* sc = scalars(s);
    JP_REQUIRE(my->n == mx->n && my->k == mx->k && my != mx, "jp_mv_axpby_dev: two different multivectors of the same shape");
    const Coeff ca = make_coeff(*sc, a_scale, a_num, a_den, "jp_mv_axpby_dev"), cb = make_coeff(*sc, b_scale, b_num, b_den, "jp_mv_axpby_dev");
    const int64_t n = (int64_t)my->n * my->k;
    if (n == 0) return;
    axpby_dev_kernel<<<grid_for(n, 4), kThreads, 0, ctx().compute>>>(my->d(), mx->d(), n, ca, cb, sc->d());
    JP_LAUNCH_CHECK();
  });
}

int32_t jp_mv_axpby_norm2sq_dev(jp_handle comm, jp_handle y, jp_handle x, jp_handle s, double a_scale, int32_t a_num, int32_t a_den,
                                double b_scale, int32_t b_num, int32_t b_den, int32_t out_slot) {
  return guarded([&] {
    require_init();
    Comm* c = get<Comm>(comm, Kind::Comm, "comm");
    MultiVector* my = get<MultiVector>(y, Kind::MultiVector, "multivector");
    MultiVector* mx = get<MultiVector>(x, Kind::MultiVector, "multivector");
    Scalars* sc = scalars(s);
    check_slot(*sc, out_slot, "jp_mv_axpby_norm2sq_dev");
    JP_REQUIRE(my->n == mx->n && my->k == 1 && mx->k == 1 && my != mx, "jp_mv_axpby_norm2sq_dev: two different single vectors of the same length");
    JP_REQUIRE(out_slot != a_num && out_slot != a_den && out_slot != b_num && out_slot != b_den,
               "jp_mv_axpby_norm2sq_dev: the result slot must not be one of the coefficient slots");
    const Coeff ca = make_coeff(*sc, a_scale, a_num, a_den, "jp_mv_axpby_norm2sq_dev"),
                cb = make_coeff(*sc, b_scale, b_num, b_den, "jp_mv_axpby_norm2sq_dev");
    double* partials = sc->ws.as<double>();
    unsigned int* ticket = reinterpret_cast<unsigned int*>(partials + kMaxBlocks);
    if (my->n == 0) {
      JP_CUDA(cudaMemsetAsync(sc->d() + out_slot, 0, 8, ctx().compute));
    } else {
      axpby_norm2_dev_kernel<<<grid_for(my->n, 8), kThreads, 0, ctx().compute>>>(my->d(), mx->d(), my->n, ca, cb, sc->d(), partials, ticket,
                                                                                 sc->d() + out_slot);
      JP_LAUNCH_CHECK();
    }
    comm_allreduce_device(*c, sc->d() + out_slot, 1, ctx().compute);
  });
}

}  // extern "C"
