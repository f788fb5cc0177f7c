/* c_abi_smoke.c -- the drop-in boundary used from plain C99, the way a foreign-function interface (Julia's ccall, cgo,
 * JNI ...) sees it: include/jpetra_b200.h compiles as C, the library is opened at run time and every call goes
 * through a pointer whose type is taken from the header's own declaration.
 *
 *   gcc -std=c99 -Iinclude examples/c_abi_smoke.c -ldl -o c_abi_smoke && ./c_abi_smoke juliapetra.jl_b200/libjpetra_b200.so
 *
 * On a machine with a B200 it runs y = A*x for the 2x2 matrix of the reference's test/CSRMatrixTests.jl:120-174 and
 * prints "apply ok"; without a GPU jp_init fails with JP_CUDA_ERROR and the message of jp_last_error -- there is no CPU
 * fallback -- which is reported as "no device" (exit code 0 in both cases, 1 on any other outcome). */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include "jpetra_b200.h"

#define LOAD(name)                                              \
  __typeof__(name)* p_##name = (__typeof__(name)*)dlsym(lib, #name); \
  if (!p_##name) {                                              \
    fprintf(stderr, "missing symbol %s\n", #name);              \
    return 1;                                                   \
  }

int main(int argc, char** argv) {
  void* lib = dlopen(argc > 1 ? argv[1] : "libjpetra_b200.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) {
    fprintf(stderr, "dlopen failed: %s\n", dlerror());
    return 1;
  }
  LOAD(jp_version) LOAD(jp_init) LOAD(jp_last_error) LOAD(jp_comm_create_serial) LOAD(jp_csr_create) LOAD(jp_mv_create)
  LOAD(jp_mv_upload) LOAD(jp_mv_download) LOAD(jp_apply) LOAD(jp_finalize)
  if (p_jp_version() != 1) return 1;
  char msg[512];
  int32_t rc = p_jp_init(0);
  if (rc != JP_OK) {
    p_jp_last_error(msg, sizeof msg);
    if (rc == JP_CUDA_ERROR && strlen(msg) > 0) {
      printf("no device: jp_init -> JP_CUDA_ERROR (%s)\n", msg);
      return 0;
    }
    fprintf(stderr, "unexpected status %d: %s\n", (int)rc, msg);
    return 1;
  }
  /* A = [[2,3],[5,7]] as 1-based CSR, X = 2s: A*X = [[10,10],[24,24]] */
  const int32_t rowptr[3] = {1, 3, 5}, colind[4] = {1, 2, 1, 2};
  const double vals[4] = {2, 3, 5, 7}, x[4] = {2, 2, 2, 2};
  double y[4] = {0, 0, 0, 0};
  jp_handle comm, A, X, Y;
  if (p_jp_comm_create_serial(&comm) || p_jp_csr_create(2, 2, 2, 4, rowptr, colind, vals, 1, &A) || p_jp_mv_create(2, 2, &X) ||
      p_jp_mv_create(2, 2, &Y) || p_jp_mv_upload(X, x) || p_jp_apply(comm, Y, A, X, 0, JP_NO_TRANS, 1.0, 0.0) || p_jp_mv_download(Y, y)) {
    p_jp_last_error(msg, sizeof msg);
    fprintf(stderr, "call failed: %s\n", msg);
    return 1;
  }
  if (y[0] != 10 || y[1] != 24 || y[2] != 10 || y[3] != 24) {
    fprintf(stderr, "wrong result %g %g %g %g\n", y[0], y[1], y[2], y[3]);
    return 1;
  }
  p_jp_finalize();
  printf("apply ok\n");
  return 0;
}
