// jp_fill.cu -- CSRMatrix fillComplete on the device: the C ABI around jp_fill_pipeline.h.
// The reference does this on the host in O(nnz) scalar work with a hash set and per-row sorts
// (src/CSRMatrix.jl:706-747, src/CSRGraphInternalMethods.jl:852-982, 636-712); here the entries are uploaded once and
// the column map, the GID->LID remap, the per-row stable sort + duplicate merge and the CSR packing all run as kernels.
// Scope: a contiguous domain map (BlockMap(N, comm) / BlockMap(N, n, comm)); general GID-list domain maps stay on
// the host path of the mirror.
#include "jp_common.h"
#include "jp_fill_pipeline.h"

using namespace jp;

extern "C" {

int32_t jp_csr_fill_complete(int32_t n_rows, const int64_t* entry_ptr, const int64_t* col_gids, const double* vals, int64_t min_all_gid,
                             int64_t max_all_gid, int64_t dom_min_gid, int32_t dom_num_my, jp_handle* out, int64_t* info) {
  return guarded([&] {
    require_init();
    JP_REQUIRE(out != nullptr, "jp_csr_fill_complete: null output");
    FillOutput fo;
    fill_complete_device(n_rows, entry_ptr, col_gids, vals, min_all_gid, max_all_gid, dom_min_gid, dom_num_my, fo);
    auto A = std::make_unique<Csr>();
    A->n_rows = n_rows;
    A->n_cols = fo.n_cols;
    A->n_local_cols = fo.n_same;  // the numSameIDs prefix of the column map aliases X's own rows (jp_csr_create's contract)
    A->nnz = fo.nnz;
    A->rowptr.swap(fo.rowptr);
    A->colind.swap(fo.colind);
    A->vals.swap(fo.vals);
    A->colmap_gids.swap(fo.colmap);
    csr_finalize_structure(*A, fo.rp_host, fo.rowmax_host);
    if (info) {
      info[0] = fo.n_cols;
      info[1] = fo.n_local_used;
      info[2] = fo.n_same;
      info[3] = fo.nnz;
    }
    *out = register_object(std::move(A));
  });
}

int32_t jp_csr_col_map(jp_handle A, int64_t* gids) {
  return guarded([&] {
    require_init();
    Csr* a = get<Csr>(A, Kind::Csr, "csr matrix");
    JP_REQUIRE_STATE(a->colmap_gids.p != nullptr, "jp_csr_col_map: the matrix was not fill-completed on the device");
    if (a->n_cols == 0) return;
    JP_REQUIRE(gids != nullptr, "jp_csr_col_map: null output");
    JP_CUDA(cudaStreamSynchronize(ctx().compute));
    JP_CUDA(cudaMemcpy(gids, a->colmap_gids.p, (size_t)a->n_cols * 8, cudaMemcpyDeviceToHost));
  });
}

}  // extern "C"
