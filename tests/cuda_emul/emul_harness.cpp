// emul_harness.cpp -- C entry points (ctypes) that run the product's kernel headers and host pipelines on the CPU
// execution model of cuda_emul.h.  TEST INFRASTRUCTURE ONLY: built by tests/cuda_emul/Makefile, loaded by
// tests/test_emul_*.py, never by the product.
#include "emul_runtime.h"

#include "../../juliapetra.jl_b200/csrc/jp_csr_blocks.h"
#include "../../juliapetra.jl_b200/csrc/jp_dia_plan.h"
#include "../../juliapetra.jl_b200/csrc/jp_cb_kernels.cuh"
#include "../../juliapetra.jl_b200/csrc/jp_fill_pipeline.h"
#include "../../juliapetra.jl_b200/csrc/jp_mv_kernels.cuh"
#include "../../juliapetra.jl_b200/csrc/jp_plan_kernels.cuh"
#include "../../juliapetra.jl_b200/csrc/jp_solver_kernels.cuh"

namespace {
std::string g_error;

template <class F>
int guarded(F&& f) {
  try {
    f();
    return 0;
  } catch (const jp::Error& e) {
    g_error = e.msg;
    return e.code;
  }
}
}  // namespace

extern "C" {

const char* emul_last_error() { return g_error.c_str(); }
uint64_t emul_launch_count() { return emul::state().launches; }
void emul_set_sm_count(int n) { jp::ctx().sm_count = n; }
// 0 forward, 1 reverse, 2 random (seeded): the order in which a block's runnable threads get their turn
void emul_set_schedule(int mode, uint64_t seed) {
  emul::schedule_mode() = mode;
  emul::schedule_rng() = 0x9E3779B97F4A7C15ull ^ (seed * 0xBF58476D1CE4E5B9ull);
}

// self-test of the model: out[t] = value written by thread t+1 into (dynamic, garbage-initialised) shared memory, with or
// without the barrier that makes it well defined
static void neighbour_kernel(int* out, int with_barrier) {
  JP_DYN_SMEM(smem_raw);
  int* s_v = reinterpret_cast<int*>(smem_raw);
  const int t = threadIdx.x, n = blockDim.x;
  s_v[t] = 1000 + t;
  if (with_barrier) __syncthreads();
  out[t] = s_v[(t + 1) % n];
}
int emul_selftest_neighbour(int* out, int nthreads, int with_barrier) {
  int* p = out;
  emul::launch(dim3(1), dim3(nthreads), (size_t)nthreads * 4, [&] { neighbour_kernel(p, with_barrier); });
  return 0;
}

// ---- fillComplete ----
int emul_fill_complete(int32_t n_rows, const int64_t* ptr, const int64_t* gids, const double* vals, int64_t gmin, int64_t gmax,
                       int64_t dom_min, int32_t dom_n, int64_t* info, void** handle) {
  return guarded([&] {
    auto* out = new jp::FillOutput();
    try {
      jp::fill_complete_device(n_rows, ptr, gids, vals, gmin, gmax, dom_min, dom_n, *out);
    } catch (...) {
      delete out;
      throw;
    }
    info[0] = out->n_cols;
    info[1] = out->n_local_used;
    info[2] = out->n_same;
    info[3] = out->nnz;
    *handle = out;
  });
}

void emul_fill_get(void* handle, int32_t* rowptr, int32_t* colind, double* vals, int64_t* colmap, int32_t* rowmax) {
  auto* out = static_cast<jp::FillOutput*>(handle);
  std::memcpy(rowptr, out->rowptr.p, ((size_t)out->n_rows + 1) * 4);
  std::memcpy(colind, out->colind.p, (size_t)out->nnz * 4);
  std::memcpy(vals, out->vals.p, (size_t)out->nnz * 8);
  std::memcpy(colmap, out->colmap.p, (size_t)out->n_cols * 8);
  std::memcpy(rowmax, out->rowmax_host.data(), (size_t)out->n_rows * 4);
}

void emul_fill_free(void* handle) { delete static_cast<jp::FillOutput*>(handle); }


// ---- localApply (csrc/jp_csr_kernels.cuh) ----
// One rank's y = beta*y + alpha*A*x through the product's kernels.  rowptr/colind are 0-based; x holds the first
// x_rows rows of the column space (ld = x_rows) and halo the remaining columns LID-major, vector-minor (halo == null:
// x is the whole column-map multivector and the HALO=false kernels run).  variant: 0 gather kernels (spmv_kernel for
// k == 1, spmm_kernel otherwise), 6 spmm_runs_kernel.  stats[4] = blocks, DIRECT, SUBWARP,
// LONGROW.  Returns 0, or -1 when the tile tables are not usable for this matrix.
int emul_local_apply(int32_t n_rows, int32_t n_cols, int32_t n_local_cols, const int32_t* rowptr, const int32_t* colind,
                     const double* vals, const double* x, int32_t x_rows, const double* halo, double* y, int32_t k, double alpha,
                     double beta, int32_t variant, int32_t cap, int32_t max_rows, int32_t T, int32_t tma, int32_t allow_direct,
                     int64_t* stats) {
  using namespace jp;
  using namespace jp::csrk;
  const int64_t nnz = rowptr[n_rows];
  std::vector<int32_t> rp(rowptr, rowptr + n_rows + 1);
  std::vector<uint8_t> cls((size_t)n_rows, 0);
  for (int32_t r = 0; r < n_rows; ++r) {
    int32_t mx = -1;
    for (int32_t i = rp[r]; i < rp[r + 1]; ++i) mx = std::max(mx, colind[i]);
    cls[r] = mx >= n_local_cols;
  }
  std::vector<BlockDesc> bi, bb, ba;
  DirectRule rule;  // same knobs as the library (csr_finalize_structure)
  if (const char* v = std::getenv("JP_SPMV_DIRECT_MAXLEN")) rule.maxlen = std::atoi(v);
  if (const char* v = std::getenv("JP_SPMV_DIRECT_RATIO")) rule.ratio = std::atof(v);
  if (const char* v = std::getenv("JP_SPMV_SUBWARP_DIV")) rule.subwarp_div = std::atof(v);
  build_blocks(rp, cls, cap, max_rows, allow_direct != 0, bi, bb, ba, rule);
  if (stats) {
    stats[0] = (int64_t)ba.size();
    stats[1] = stats[2] = stats[3] = 0;
    for (const BlockDesc& b : ba) {
      const int kind = (b.nrows_kind >> 16) & 0xf;
      ++stats[kind == KIND_DIRECT ? 1 : kind == KIND_LONGROW ? 3 : 2];
    }
  }
  // device arrays with the slack jp_csr_create gives them
  DeviceBuffer d_rp, d_ci, d_va, d_blocks, d_x, d_halo, d_y, d_ucols, d_lidx;
  d_rp.alloc(((size_t)n_rows + 1 + 16) * 4);
  std::memset(d_rp.p, 0, d_rp.bytes);
  std::memcpy(d_rp.p, rp.data(), rp.size() * 4);
  d_ci.alloc(((size_t)nnz + 64) * 4);
  d_va.alloc(((size_t)nnz + 64) * 8);
  std::memset(d_ci.p, 0, d_ci.bytes);
  std::memset(d_va.p, 0, d_va.bytes);
  std::memcpy(d_ci.p, colind, (size_t)nnz * 4);
  std::memcpy(d_va.p, vals, (size_t)nnz * 8);
  const bool use_halo = halo != nullptr;
  const int32_t n_remote = n_cols - x_rows;
  d_x.alloc((size_t)x_rows * k * 8 + 64);
  std::memcpy(d_x.p, x, (size_t)x_rows * k * 8);
  d_halo.alloc((size_t)std::max(n_remote, 0) * k * 8 + 64);
  if (use_halo) std::memcpy(d_halo.p, halo, (size_t)n_remote * k * 8);
  d_y.alloc((size_t)n_rows * k * 8 + 64);
  std::memcpy(d_y.p, y, (size_t)n_rows * k * 8);
  const int32_t nloc = use_halo ? n_local_cols : n_cols;
  const size_t smem = stage_smem_bytes(cap);
  const unsigned nb = (unsigned)ba.size();
  constexpr int KT = 8;
  int rc = 0;
  if (nb == 0 || k == 0) {
  } else if (variant == 0) {
    d_blocks.alloc(ba.size() * sizeof(BlockDesc) + 16);
    std::memcpy(d_blocks.p, ba.data(), ba.size() * sizeof(BlockDesc));
    const BlockDesc* b = d_blocks.as<BlockDesc>();
    auto launch = [&](auto kernel_spmv, auto kernel_spmm) {
      if (k == 1)
        JP_KLAUNCH(kernel_spmv, nb, kSpmvThreads, smem, nullptr, b, d_rp.as<int32_t>(), d_ci.as<int32_t>(), d_va.as<double>(),
                   d_x.as<double>(), d_halo.as<double>(), nloc, d_y.as<double>(), cap, alpha, beta);
      else
        JP_KLAUNCH(kernel_spmm, nb, kSpmvThreads, smem, nullptr, b, d_rp.as<int32_t>(), d_ci.as<int32_t>(), d_va.as<double>(),
                   d_x.as<double>(), (int64_t)x_rows, d_halo.as<double>(), k, nloc, d_y.as<double>(), (int64_t)n_rows, k, cap, T, alpha,
                   beta);
    };
    if (use_halo && tma) launch(spmv_kernel<true, 1>, spmm_kernel<KT, true, 1>);
    else if (use_halo) launch(spmv_kernel<true, 0>, spmm_kernel<KT, true, 0>);
    else if (tma) launch(spmv_kernel<false, 1>, spmm_kernel<KT, false, 1>);
    else launch(spmv_kernel<false, 0>, spmm_kernel<KT, false, 0>);
  } else if (variant == 6) {
    // spmm_runs_kernel in the library's configuration: KT = 4, 128 rows per block, 128 threads (one lane per row)
    if (use_halo || (x_rows & 1)) return -3;
    std::vector<BlockDesc> ri, rb, ra;
    build_blocks(rp, cls, cap, kRunsMaxRows, allow_direct != 0, ri, rb, ra);
    const int kt = 4;
    const int max_tile_cols = (int)((110 * 1024 - (int64_t)runs_smem_bytes(cap, 0, kt)) / (8 * kt)) & ~1;
    std::vector<int32_t> ci(colind, colind + nnz);
    RunTables t;
    if (!build_run_tables(ci, ra, max_tile_cols, t)) {
      rc = -1;
    } else {
      if (stats) stats[0] = (int64_t)ra.size();
      d_blocks.alloc(ra.size() * sizeof(BlockDesc) + 16);
      std::memcpy(d_blocks.p, ra.data(), ra.size() * sizeof(BlockDesc));
      d_ucols.alloc(t.runs.size() * 4);
      d_lidx.alloc(t.lidx.size() * 2);
      std::memcpy(d_ucols.p, t.runs.data(), t.runs.size() * 4);
      std::memcpy(d_lidx.p, t.lidx.data(), t.lidx.size() * 2);
      const int wpad = (t.wmax + 1) & ~1;
      const size_t sm = runs_smem_bytes(cap, wpad, kt);
      JP_KLAUNCH((spmm_runs_kernel<4, 128>), (unsigned)ra.size(), 128, sm, nullptr, d_blocks.as<BlockDesc>(), d_rp.as<int32_t>(),
                 d_va.as<double>(), d_lidx.as<uint16_t>(), d_ucols.as<int2>(), d_x.as<double>(), (int64_t)x_rows, d_y.as<double>(),
                 (int64_t)n_rows, k, cap, wpad, alpha, beta);
    }
  } else {
    return -5;  // unknown variant
  }
  std::memcpy(y, d_y.p, (size_t)n_rows * k * 8);
  return rc;
}


// ---- fused CG building blocks ----
// y = beta*y + alpha*A*x and out = sum_r w[r]*y[r] (spmv_dot_kernel), `repeats` times with the same workspace so the
// workspace is reused; y is reset to y0 before every repeat.  x is the whole column-map vector.
int emul_spmv_dot(int32_t n_rows, int32_t n_cols, const int32_t* rowptr, const int32_t* colind, const double* vals, const double* x,
                  const double* w, double* y, double alpha, double beta, int32_t cap, int32_t max_rows, int32_t tma, int32_t repeats,
                  double* out) {
  using namespace jp;
  using namespace jp::csrk;
  const int64_t nnz = rowptr[n_rows];
  std::vector<int32_t> rp(rowptr, rowptr + n_rows + 1);
  std::vector<uint8_t> cls((size_t)n_rows, 0);
  std::vector<BlockDesc> bi, bb, ba;
  build_blocks(rp, cls, cap, max_rows, true, bi, bb, ba);
  const unsigned nb = (unsigned)ba.size();
  DeviceBuffer d_rp, d_ci, d_va, d_blocks, d_x, d_w, d_y, ws;
  d_rp.alloc(((size_t)n_rows + 1 + 16) * 4);
  std::memset(d_rp.p, 0, d_rp.bytes);
  std::memcpy(d_rp.p, rp.data(), rp.size() * 4);
  d_ci.alloc(((size_t)nnz + 64) * 4);
  d_va.alloc(((size_t)nnz + 64) * 8);
  std::memset(d_ci.p, 0, d_ci.bytes);
  std::memset(d_va.p, 0, d_va.bytes);
  std::memcpy(d_ci.p, colind, (size_t)nnz * 4);
  std::memcpy(d_va.p, vals, (size_t)nnz * 8);
  d_blocks.alloc(ba.size() * sizeof(BlockDesc) + 16);
  std::memcpy(d_blocks.p, ba.data(), ba.size() * sizeof(BlockDesc));
  d_x.alloc((size_t)n_cols * 8 + 64);
  std::memcpy(d_x.p, x, (size_t)n_cols * 8);
  d_w.alloc((size_t)n_rows * 8 + 64);
  std::memcpy(d_w.p, w, (size_t)n_rows * 8);
  d_y.alloc((size_t)n_rows * 8 + 64);
  const unsigned np = (unsigned)nb * kDotPartialsPerBlock;  // one partial per warp
  const unsigned g2 = sum_partials_grid(np);                // workspace layout of the library (jp_csr.cu dot_workspace)
  ws.alloc(((size_t)np + g2 + 4) * 8);
  std::memset(ws.p, 0, ws.bytes);
  double* partials = ws.as<double>();
  double* level2 = partials + np;
  double* d_out = level2 + g2;
  unsigned int* ticket = reinterpret_cast<unsigned int*>(d_out + 1);
  const size_t smem = stage_smem_bytes(cap);
  for (int rep = 0; rep < repeats; ++rep) {
    std::memcpy(d_y.p, y, (size_t)n_rows * 8);
    *d_out = -12345.0;
    if (nb == 0) {
      *d_out = 0.0;
    } else if (tma) {
      JP_KLAUNCH((spmv_dot_kernel<false, 1>), nb, kSpmvThreads, smem, nullptr, d_blocks.as<BlockDesc>(), d_rp.as<int32_t>(),
                 d_ci.as<int32_t>(), d_va.as<double>(), d_x.as<double>(), (const double*)nullptr, n_cols, d_y.as<double>(), cap, alpha, beta,
                 d_w.as<double>(), partials);
    } else {
      JP_KLAUNCH((spmv_dot_kernel<false, 0>), nb, kSpmvThreads, smem, nullptr, d_blocks.as<BlockDesc>(), d_rp.as<int32_t>(),
                 d_ci.as<int32_t>(), d_va.as<double>(), d_x.as<double>(), (const double*)nullptr, n_cols, d_y.as<double>(), cap, alpha, beta,
                 d_w.as<double>(), partials);
    }
    if (nb) JP_KLAUNCH(sum_partials_kernel, g2, kSpmvThreads, 0, nullptr, partials, np, d_out, level2, ticket);
    out[rep] = *d_out;
  }
  std::memcpy(y, d_y.p, (size_t)n_rows * 8);
  return 0;
}

// y = a*x + b*y (n x k, column-major) and out[v] = sum y_new[:, v]^2, `repeats` times on the same workspace
int emul_axpby_norm2(double* y, const double* x, int64_t n, int32_t k, double a, double b, int32_t blocks, int32_t repeats, double* out) {
  using namespace jp;
  DeviceBuffer d_y, d_x, ws;
  d_y.alloc((size_t)n * k * 8 + 64);
  d_x.alloc((size_t)n * k * 8 + 64);
  std::memcpy(d_y.p, y, (size_t)n * k * 8);
  std::memcpy(d_x.p, x, (size_t)n * k * 8);
  ws.alloc(((size_t)k * blocks + 2 * (size_t)k) * 8);
  std::memset(ws.p, 0, ws.bytes);
  double* partials = ws.as<double>();
  double* d_out = partials + (size_t)k * blocks;
  unsigned int* tickets = reinterpret_cast<unsigned int*>(d_out + k);
  for (int rep = 0; rep < repeats; ++rep) {
    JP_KLAUNCH(mvk::axpby_norm2_kernel, dim3(blocks, k), mvk::kThreads, 0, nullptr, d_y.as<double>(), d_x.as<double>(), n, a, b, partials,
               tickets, d_out);
    for (int v = 0; v < k; ++v) out[(size_t)rep * k + v] = d_out[v];
  }
  std::memcpy(y, d_y.p, (size_t)n * k * 8);
  return 0;
}


// ---- MultiVector kernels (csrc/jp_mv_kernels.cuh) ----
// op: 0 dot(x, y), 1 sum x^2, 2 sum x^p -> out[k] (the local sums before the allreduce), `repeats` launches on one
// workspace (the ticket must re-arm); x, y: n x k column-major
int emul_mv_reduce(int32_t op, const double* x, const double* y, int64_t n, int32_t k, double p, int32_t blocks, int32_t repeats, double* out) {
  using namespace jp;
  using namespace jp::mvk;
  DeviceBuffer d_x, d_y, ws;
  d_x.alloc((size_t)n * k * 8 + 64);
  d_y.alloc((size_t)n * k * 8 + 64);
  std::memcpy(d_x.p, x, (size_t)n * k * 8);
  if (y) std::memcpy(d_y.p, y, (size_t)n * k * 8);
  ws.alloc(((size_t)k * blocks + 2 * (size_t)k) * 8);
  std::memset(ws.p, 0, ws.bytes);
  double* partials = ws.as<double>();
  double* d_out = partials + (size_t)k * blocks;
  unsigned int* tickets = reinterpret_cast<unsigned int*>(d_out + k);
  for (int rep = 0; rep < repeats; ++rep) {
    if (op == 0)
      JP_KLAUNCH(reduce_kernel<0>, dim3(blocks, k), kThreads, 0, nullptr, d_x.as<double>(), d_y.as<double>(), n, p, partials, tickets, d_out);
    else if (op == 1)
      JP_KLAUNCH(reduce_kernel<1>, dim3(blocks, k), kThreads, 0, nullptr, d_x.as<double>(), (const double*)nullptr, n, p, partials, tickets, d_out);
    else
      JP_KLAUNCH(reduce_kernel<2>, dim3(blocks, k), kThreads, 0, nullptr, d_x.as<double>(), (const double*)nullptr, n, p, partials, tickets, d_out);
    for (int v = 0; v < k; ++v) out[(size_t)rep * k + v] = d_out[v];
  }
  return 0;
}

// op: 0 fill(y, a), 1 scale(y, a), 2 scale_vec(y, alpha_k), 3 axpby(y, a, x, b); y is n x k column-major, updated in place
int emul_mv_elementwise(int32_t op, double* y, const double* x, const double* alpha_k, int64_t n, int32_t k, double a, double b, int32_t blocks) {
  using namespace jp;
  using namespace jp::mvk;
  DeviceBuffer d_y, d_x, d_a;
  d_y.alloc((size_t)n * k * 8 + 64);
  d_x.alloc((size_t)n * k * 8 + 64);
  d_a.alloc((size_t)k * 8 + 64);
  std::memcpy(d_y.p, y, (size_t)n * k * 8);
  if (x) std::memcpy(d_x.p, x, (size_t)n * k * 8);
  if (alpha_k) std::memcpy(d_a.p, alpha_k, (size_t)k * 8);
  const int64_t total = n * k;
  if (op == 0) JP_KLAUNCH(fill_kernel, (unsigned)blocks, kThreads, 0, nullptr, d_y.as<double>(), total, a);
  else if (op == 1) JP_KLAUNCH(scale_kernel, (unsigned)blocks, kThreads, 0, nullptr, d_y.as<double>(), total, a);
  else if (op == 2) JP_KLAUNCH(scale_vec_kernel, dim3(blocks, k), kThreads, 0, nullptr, d_y.as<double>(), n, d_a.as<double>());
  else JP_KLAUNCH(axpby_kernel, (unsigned)blocks, kThreads, 0, nullptr, d_y.as<double>(), d_x.as<double>(), total, a, b);
  std::memcpy(y, d_y.p, (size_t)n * k * 8);
  return 0;
}

// ---- transfer kernels (csrc/jp_plan_kernels.cuh); LIDs are 0-based ----
// op: 0 pack (buf <- src[lids]), 1 unpack (tgt[lids] <- buf), 2 scatter_add (tgt[lids] += buf), 3 permute (tgt[to] <- src[from]),
//     4 add_permute (tgt[to] += src[from]), 5 add_rows (tgt[0:n_ids] += src[0:n_ids]).  mv: n_mv x k (tgt for 1, 2; src for 0);
//     other: the second multivector of ops 3-5 (n_other x k) or the LID-major buffer (n_ids x k, vector-minor) of ops 0-2.
int emul_plan_op(int32_t op, double* mv, int64_t n_mv, double* other, int64_t n_other, const int32_t* lids_a, const int32_t* lids_b,
                 int32_t n_ids, int32_t k, int32_t blocks) {
  using namespace jp;
  using namespace jp::plank;
  DeviceBuffer d_mv, d_other, d_a, d_b;
  const size_t other_elems = op <= 2 ? (size_t)n_ids * k : (size_t)n_other * k;
  d_mv.alloc((size_t)n_mv * k * 8 + 64);
  d_other.alloc(other_elems * 8 + 64);
  d_a.alloc((size_t)n_ids * 4 + 64);
  d_b.alloc((size_t)n_ids * 4 + 64);
  std::memcpy(d_mv.p, mv, (size_t)n_mv * k * 8);
  std::memcpy(d_other.p, other, other_elems * 8);
  if (lids_a) std::memcpy(d_a.p, lids_a, (size_t)n_ids * 4);
  if (lids_b) std::memcpy(d_b.p, lids_b, (size_t)n_ids * 4);
  const unsigned g = (unsigned)blocks;
  if (op == 0) JP_KLAUNCH(pack_kernel, g, kThreads, 0, nullptr, d_mv.as<double>(), n_mv, d_a.as<int32_t>(), n_ids, k, d_other.as<double>());
  else if (op == 1) JP_KLAUNCH(unpack_kernel, g, kThreads, 0, nullptr, d_mv.as<double>(), n_mv, d_a.as<int32_t>(), n_ids, k, d_other.as<double>());
  else if (op == 2) JP_KLAUNCH(scatter_add_kernel, g, kThreads, 0, nullptr, d_mv.as<double>(), n_mv, d_a.as<int32_t>(), n_ids, k, d_other.as<double>());
  else if (op == 3) JP_KLAUNCH(permute_kernel, g, kThreads, 0, nullptr, d_mv.as<double>(), n_mv, d_other.as<double>(), n_other, d_a.as<int32_t>(), d_b.as<int32_t>(), n_ids, k);
  else if (op == 4) JP_KLAUNCH(add_permute_kernel, g, kThreads, 0, nullptr, d_mv.as<double>(), n_mv, d_other.as<double>(), n_other, d_a.as<int32_t>(), d_b.as<int32_t>(), n_ids, k);
  else JP_KLAUNCH(add_rows_kernel, g, kThreads, 0, nullptr, d_mv.as<double>(), n_mv, d_other.as<double>(), n_other, n_ids, k);
  std::memcpy(mv, d_mv.p, (size_t)n_mv * k * 8);
  std::memcpy(other, d_other.p, other_elems * 8);
  return 0;
}

// ---- spmm_dia_kernel (diagonal-structured matrices): the library's plan (jp_dia_plan.h) + a host copy of what
//      dia_mark_kernel / dia_fill_kernel produce.  Rows [range[0], range[1]) of y are computed; rc -1: not eligible ----
int emul_spmm_dia(int32_t n_rows, int32_t n_cols, const int32_t* rowptr, const int32_t* colind, const double* vals, const double* x, double* y,
                  int32_t k, double alpha, double beta, int32_t parity, int32_t* range) {
  using namespace jp;
  using namespace jp::csrk;
  const int64_t nnz = rowptr[n_rows];
  // offsets through the bitmap kernel
  const size_t nwords = ((size_t)n_rows + n_cols + 31) / 32;
  DeviceBuffer d_rp, d_ci, d_va, bitmap;
  d_rp.alloc(((size_t)n_rows + 1) * 4);
  d_ci.alloc((size_t)nnz * 4 + 16);
  d_va.alloc((size_t)nnz * 8 + 16);
  bitmap.alloc(nwords * 4);
  std::memcpy(d_rp.p, rowptr, ((size_t)n_rows + 1) * 4);
  std::memcpy(d_ci.p, colind, (size_t)nnz * 4);
  std::memcpy(d_va.p, vals, (size_t)nnz * 8);
  std::memset(bitmap.p, 0, nwords * 4);
  JP_KLAUNCH(dia_mark_kernel, 3u, 256, 0, nullptr, d_rp.as<int32_t>(), d_ci.as<int32_t>(), n_rows, bitmap.as<uint32_t>());
  std::vector<int32_t> offsets;
  for (size_t w = 0; w < nwords; ++w)
    for (int b = 0; b < 32; ++b)
      if (bitmap.as<uint32_t>()[w] >> b & 1u) offsets.push_back((int32_t)((int64_t)w * 32 + b - (n_rows - 1)));
  DiaParams P;
  if (offsets.size() > (size_t)kDiaMaxDiags || !dia_plan(offsets, P, parity)) return -1;
  int32_t row_lo, row_hi;
  dia_row_range(P, n_rows, n_cols, row_lo, row_hi, parity);
  range[0] = row_lo;
  range[1] = row_hi;
  if (row_hi <= row_lo) return -1;
  const int64_t npad = (((int64_t)n_rows + kDiaThreads - 1) / kDiaThreads + 1) * kDiaThreads;
  DeviceBuffer d_dia, d_off, d_missing, d_x, d_y;
  d_dia.alloc((size_t)P.D * npad * 8);
  std::memset(d_dia.p, 0, d_dia.bytes);
  d_off.alloc(offsets.size() * 4);
  std::memcpy(d_off.p, offsets.data(), offsets.size() * 4);
  d_missing.alloc(8);
  std::memset(d_missing.p, 0, 8);
  JP_KLAUNCH(dia_fill_kernel, 2u, 256, 0, nullptr, d_rp.as<int32_t>(), d_ci.as<int32_t>(), d_va.as<double>(), n_rows, P, d_off.as<int32_t>(), npad,
             d_dia.as<double>(), d_missing.as<unsigned long long>());
  if (*d_missing.as<unsigned long long>() != 0) return -2;
  d_x.alloc((size_t)n_cols * k * 8);  // exactly n_cols rows: a window that leaves X faults in the emulator's guarded buffers
  std::memcpy(d_x.p, x, (size_t)n_cols * k * 8);
  d_y.alloc((size_t)n_rows * k * 8 + 64);
  std::memcpy(d_y.p, y, (size_t)n_rows * k * 8);
  const unsigned grid = (unsigned)((row_hi - row_lo + kDiaThreads - 1) / kDiaThreads);
  constexpr int KT = 4;
  const size_t smem = dia_smem_bytes(P, KT);
  if (P.D <= 8)
    JP_KLAUNCH((spmm_dia_kernel<KT, 8>), grid, kDiaThreads, smem, nullptr, d_dia.as<double>(), npad, d_x.as<double>(), (int64_t)n_cols,
               d_y.as<double>(), (int64_t)n_rows, k, row_lo, row_hi, alpha, beta, P);
  else if (P.D <= 28)
    JP_KLAUNCH((spmm_dia_kernel<KT, 28>), grid, kDiaThreads, smem, nullptr, d_dia.as<double>(), npad, d_x.as<double>(), (int64_t)n_cols,
               d_y.as<double>(), (int64_t)n_rows, k, row_lo, row_hi, alpha, beta, P);
  else
    JP_KLAUNCH((spmm_dia_kernel<KT, 32>), grid, kDiaThreads, smem, nullptr, d_dia.as<double>(), npad, d_x.as<double>(), (int64_t)n_cols,
               d_y.as<double>(), (int64_t)n_rows, k, row_lo, row_hi, alpha, beta, P);
  std::memcpy(y, d_y.p, (size_t)n_rows * k * 8);
  return 0;
}

// ---- spmv_cb_kernel / spmv_cb_carry_kernel (column-slab SpMV of large irregular matrices, csrc/jp_cb_kernels.cuh) with the
//      device builders (cb_keys / cb_expand_rows / cb_gather / cb_slab_ptr) and a host stable sort in place of the radix
//      sort.  slab_cols: columns per slab.  y = beta*y + alpha*A*x, k = 1 ----
int emul_spmv_cb(int32_t n_rows, int32_t n_cols, const int32_t* rowptr, const int32_t* colind, const double* vals, const double* x, double* y,
                 double alpha, double beta, int32_t slab_cols) {
  using namespace jp;
  using namespace jp::cbk;
  const int64_t nnz = rowptr[n_rows];
  const int32_t n_slabs = (n_cols + slab_cols - 1) / slab_cols;
  DeviceBuffer d_rp, d_ci, d_va, key, order, key_s, order_s, row_of, row, col, val, slab_ptr, carry, d_x, d_y;
  d_rp.alloc(((size_t)n_rows + 1) * 4);
  d_ci.alloc((size_t)nnz * 4 + 16);
  d_va.alloc((size_t)nnz * 8 + 16);
  std::memcpy(d_rp.p, rowptr, ((size_t)n_rows + 1) * 4);
  std::memcpy(d_ci.p, colind, (size_t)nnz * 4);
  std::memcpy(d_va.p, vals, (size_t)nnz * 8);
  key.alloc((size_t)nnz * 4 + 16);
  order.alloc((size_t)nnz * 4 + 16);
  key_s.alloc((size_t)nnz * 4 + 16);
  order_s.alloc((size_t)nnz * 4 + 16);
  row_of.alloc((size_t)nnz * 4 + 16);
  JP_KLAUNCH(cb_keys_kernel, 3u, kCbThreads, 0, nullptr, d_ci.as<int32_t>(), nnz, slab_cols, key.as<uint32_t>(), order.as<uint32_t>());
  JP_KLAUNCH(cb_expand_rows_kernel, 2u, kCbThreads, 0, nullptr, d_rp.as<int32_t>(), n_rows, row_of.as<int32_t>());
  {  // stable sort of the positions by slab (the library uses cub::DeviceRadixSort, which is stable)
    std::vector<uint32_t> idx(order.as<uint32_t>(), order.as<uint32_t>() + nnz);
    const uint32_t* k = key.as<uint32_t>();
    std::stable_sort(idx.begin(), idx.end(), [k](uint32_t a, uint32_t b) { return k[a] < k[b]; });
    for (int64_t i = 0; i < nnz; ++i) {
      order_s.as<uint32_t>()[i] = idx[i];
      key_s.as<uint32_t>()[i] = k[idx[i]];
    }
  }
  row.alloc((size_t)nnz * 4);  // exact sizes: an over-read faults in the guarded buffers
  col.alloc((size_t)nnz * 4);
  val.alloc((size_t)nnz * 8);
  JP_KLAUNCH(cb_gather_kernel, 3u, kCbThreads, 0, nullptr, order_s.as<uint32_t>(), row_of.as<int32_t>(), d_ci.as<int32_t>(), d_va.as<double>(), nnz,
             row.as<int32_t>(), col.as<int32_t>(), val.as<double>());
  slab_ptr.alloc(((size_t)n_slabs + 1) * 8);
  JP_KLAUNCH(cb_slab_ptr_kernel, 2u, kCbThreads, 0, nullptr, key_s.as<uint32_t>(), nnz, n_slabs, slab_ptr.as<long long>());
  d_x.alloc((size_t)n_cols * 8);
  std::memcpy(d_x.p, x, (size_t)n_cols * 8);
  d_y.alloc((size_t)n_rows * 8);
  for (int32_t r = 0; r < n_rows; ++r) d_y.as<double>()[r] = beta == 0.0 ? 0.0 : y[r] * beta;  // launch_scale_fill
  int64_t max_slab = 0;
  for (int s = 0; s < n_slabs; ++s) max_slab = std::max<int64_t>(max_slab, slab_ptr.as<long long>()[s + 1] - slab_ptr.as<long long>()[s]);
  carry.alloc((size_t)(2 * ((max_slab + 31) / 32) + 2) * sizeof(CbCarry));
  for (int s = 0; s < n_slabs; ++s) {
    const int64_t e0 = slab_ptr.as<long long>()[s], e1 = slab_ptr.as<long long>()[s + 1];
    if (e1 <= e0) continue;
    const int64_t passes = (e1 - e0 + 31) / 32;
    const unsigned grid = (unsigned)((e1 - e0 + kCbPerBlock - 1) / kCbPerBlock);
    JP_KLAUNCH(spmv_cb_kernel, grid, kCbThreads, 0, nullptr, row.as<int32_t>(), col.as<int32_t>(), val.as<double>(), e0, e1, d_x.as<double>(),
               d_y.as<double>(), alpha, carry.as<CbCarry>());
    JP_KLAUNCH(spmv_cb_carry_kernel, 2u, kCbThreads, 0, nullptr, carry.as<CbCarry>(), 2 * passes, d_y.as<double>(), alpha);
  }
  std::memcpy(y, d_y.p, (size_t)n_rows * 8);
  return 0;
}

// ---- update kernels with device-resident coefficients (csrc/jp_solver_kernels.cuh) ----
// y = a*x + b*y with a = a_scale * s[a_num] / s[a_den] (b likewise); with_norm: also s[out_slot] = sum of squares of the
// result.  `repeats` launches on one workspace.  y (n doubles) and s (ns doubles) are updated in place.
int emul_axpby_dev(double* y, const double* x, int64_t n, double* s, int32_t ns, double a_scale, int32_t a_num, int32_t a_den, double b_scale,
                   int32_t b_num, int32_t b_den, int32_t with_norm, int32_t out_slot, int32_t blocks, int32_t repeats) {
  using namespace jp;
  using namespace jp::solk;
  DeviceBuffer d_y, d_x, d_s, ws;
  d_y.alloc((size_t)n * 8 + 64);
  d_x.alloc((size_t)n * 8 + 64);
  d_s.alloc((size_t)ns * 8);
  ws.alloc((size_t)(blocks + 2) * 8);
  std::memcpy(d_y.p, y, (size_t)n * 8);
  std::memcpy(d_x.p, x, (size_t)n * 8);
  std::memcpy(d_s.p, s, (size_t)ns * 8);
  std::memset(ws.p, 0, ws.bytes);
  Coeff ca{a_scale, a_num, a_den}, cb{b_scale, b_num, b_den};
  double* partials = ws.as<double>();
  unsigned int* ticket = reinterpret_cast<unsigned int*>(partials + blocks);
  for (int rep = 0; rep < repeats; ++rep) {
    if (with_norm)
      JP_KLAUNCH(axpby_norm2_dev_kernel, (unsigned)blocks, kThreads, 0, nullptr, d_y.as<double>(), d_x.as<double>(), n, ca, cb, d_s.as<double>(),
                 partials, ticket, d_s.as<double>() + out_slot);
    else
      JP_KLAUNCH(axpby_dev_kernel, (unsigned)blocks, kThreads, 0, nullptr, d_y.as<double>(), d_x.as<double>(), n, ca, cb, d_s.as<double>());
  }
  std::memcpy(y, d_y.p, (size_t)n * 8);
  std::memcpy(s, d_s.p, (size_t)ns * 8);
  return 0;
}

}  // extern "C"
