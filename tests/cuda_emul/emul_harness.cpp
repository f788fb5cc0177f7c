// emul_harness.cpp -- C entry points (ctypes) that run the product's kernel headers and host pipelines on the CPU
// execution model of cuda_emul.h.  TEST INFRASTRUCTURE ONLY: built by tests/cuda_emul/Makefile, loaded by
// tests/test_emul_*.py, never by the product.
#include "emul_runtime.h"

#include "../../juliapetra.jl_b200/csrc/jp_fill_pipeline.h"

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

}  // extern "C"
