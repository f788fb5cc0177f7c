/* jpetra_b200.h -- C ABI of the B200-native backend for JuliaPetra's sparse
 * matrix-vector hot path (CSRMatrix apply!, Import/Export halo exchange,
 * MultiVector dot/norm).  This is the drop-in boundary: plain pointers, sizes and
 * opaque int64 handles; no CUDA, NCCL or torch types.  A Julia host binds these
 * with `ccall` (see INTEGRATION.md and juliapetra.jl_b200/julia/JuliaPetraB200.jl).
 *
 * Citations `src/...` are files of the reference (Collegeville/JuliaPetra.jl):
 * each entry point names the reference interface it replaces.
 *
 * Conventions
 *  - every function returns an int32 status: JP_OK or an error class; the message
 *    is read with jp_last_error().  JP_INVALID_ARGUMENT / JP_INVALID_STATE map to
 *    the reference's InvalidArgumentError / InvalidStateError (src/Utils.jl:37-51).
 *    Nothing unwinds across the boundary.
 *  - host arrays are copied during the call and never retained.
 *  - index arrays carry an explicit `index_base` (1 from Julia, 0 from C/Python).
 *  - PIDs are 1-based everywhere, like the reference (src/MPIComm.jl:81-83).
 *  - MultiVector data is column-major n_local x k (src/DenseMultiVector.jl:6-11).
 *  - one host thread per rank; all ranks call collectives in the same order.  Work
 *    is stream-ordered on the library's compute/comm streams; only calls that
 *    return host data synchronise.
 *  - there is no CPU fallback: every compute call fails with JP_CUDA_ERROR when no
 *    sm_100 device is present.
 */
#ifndef JPETRA_B200_H
#define JPETRA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t jp_handle; /* opaque object id; 0 is never valid */

enum { JP_OK = 0, JP_INVALID_ARGUMENT = 1, JP_INVALID_STATE = 2, JP_CUDA_ERROR = 3, JP_NCCL_ERROR = 4 };
/* TransposeMode, src/Operator.jl:10 */
enum { JP_NO_TRANS = 1, JP_TRANS = 2, JP_CONJ_TRANS = 3 };
/* CombineMode, src/DistObject.jl:21 */
enum { JP_ADD = 1, JP_INSERT = 2, JP_REPLACE = 3, JP_ABSMAX = 4, JP_ZERO = 5 };
/* reduction ops / element types of the host-array collectives */
enum { JP_SUM = 0, JP_MAX = 1, JP_MIN = 2 };
enum { JP_F64 = 0, JP_I64 = 1 };

/* ---- runtime ---------------------------------------------------------------- */
int32_t jp_init(int32_t device);            /* bind this process to a CUDA device, create streams (idempotent) */
int32_t jp_finalize(void);                  /* destroy every live object and the streams */
int32_t jp_last_error(char* buf, int64_t len);
int32_t jp_device_info(int32_t* sm_count, int64_t* hbm_bytes, int32_t* cc_major, int32_t* cc_minor);
int32_t jp_synchronize(void);               /* wait for the compute and comm streams */
int32_t jp_version(void);                   /* ABI version, currently 1 */
/* pinned host staging for callers that want full-speed H2D/D2H (optional) */
int32_t jp_host_alloc(int64_t bytes, void** ptr);
int32_t jp_host_free(void* ptr);
/* CUDA events on the compute stream (device timing without torch) */
int32_t jp_event_create(jp_handle* ev);
int32_t jp_event_record(jp_handle ev);
int32_t jp_event_elapsed_ms(jp_handle start, jp_handle stop, double* ms); /* synchronises on `stop` */
int32_t jp_event_destroy(jp_handle ev);
/* number of kernels this library has launched since jp_init (bench.py: gpu_launches) */
int32_t jp_launch_count(int64_t* n);
/* write `bytes` of scratch device memory (L2 flush between timed iterations) */
int32_t jp_flush_l2(int64_t bytes);

/* ---- Comm: replaces MPIComm (src/MPIComm.jl:42-91) / SerialComm
 *      (src/SerialComm.jl:90-131) behind the abstract Comm (src/Comm.jl:12-57) -- */
int32_t jp_comm_unique_id(uint8_t* id128);  /* rank 0 creates, the host bootstrap distributes 128 bytes */
int32_t jp_comm_create_serial(jp_handle* comm);
int32_t jp_comm_create_nccl(const uint8_t* id128, int32_t nranks, int32_t rank0, jp_handle* comm);
int32_t jp_comm_destroy(jp_handle comm);
int32_t jp_comm_rank(jp_handle comm, int32_t* pid1, int32_t* nproc);        /* myPid / numProc */
int32_t jp_comm_barrier(jp_handle comm);                                     /* barrier */
/* sumAll / maxAll / minAll on host arrays (element-wise over ranks) */
int32_t jp_comm_allreduce(jp_handle comm, int32_t op, int32_t dtype, const void* in, void* out, int64_t n);
/* scanSum: inclusive prefix over ranks */
int32_t jp_comm_scan_sum(jp_handle comm, int32_t dtype, const void* in, void* out, int64_t n);
/* gatherAll: counts[nproc] are the per-rank lengths (gather them first with n=1); out has sum(counts) */
int32_t jp_comm_gather_all(jp_handle comm, int32_t dtype, const void* in, int64_t n, const int64_t* counts, void* out);
/* broadcastAll from 1-based root, in place */
int32_t jp_comm_broadcast(jp_handle comm, int32_t dtype, void* inout, int64_t n, int32_t root1);
/* the exchange Distributor.resolve performs at setup time on host records of elem_bytes each
 * (src/MPIDistributor.jl:451-590): counts are per destination / per source PID (length nproc) */
int32_t jp_comm_alltoallv(jp_handle comm, int32_t elem_bytes, const void* send, const int64_t* sendcounts, void* recv,
                          const int64_t* recvcounts);

/* ---- LocalCSRMatrix: device-resident CSR with column indices already local to the
 *      column map (src/LocalCSRMatrix.jl:3-8, src/LocalCSRGraph.jl:11-14; produced by
 *      fillLocalGraphAndMatrix, src/CSRMatrix.jl:403-484).  Columns < n_local_cols are
 *      the numSameIDs prefix of the column map (they alias X's own rows); the rest are
 *      remote columns in column-map order. ---------------------------------------- */
int32_t jp_csr_create(int32_t n_rows, int32_t n_cols, int32_t n_local_cols, int64_t nnz, const int32_t* rowptr,
                      const int32_t* colind, const double* vals, int32_t index_base, jp_handle* out);
int32_t jp_csr_destroy(jp_handle A);
/* info[8] = n_rows, n_cols, n_local_cols, nnz, interior row blocks, boundary row blocks, max row length,
 *           threads per row chosen from the row-length statistics */
int32_t jp_csr_info(jp_handle A, int64_t* info);
/* which derived layouts the matrix has (each decided lazily at the first apply that could use it; 0 not decided, 1 active,
 * -1 not used): info[10] = reserved (-1), run-staged SpMM tables, row blocks of the run-staged path, widest X tile in
 * columns, diagonal-major copy for SpMM, its number of diagonals, rows served by the diagonal kernel, width of its X tile
 * in columns, column-slab copy for SpMV on large irregular matrices, its number of slabs */
int32_t jp_csr_spmm_info(jp_handle A, int64_t* info);
/* read the device copy back (0-based) -- parity tests */
int32_t jp_csr_download(jp_handle A, int32_t* rowptr, int32_t* colind, double* vals);

/* fillComplete on the device for a CONTIGUOUS domain map (src/CSRMatrix.jl:706-747): __makeColMap
 * (src/CSRGraphInternalMethods.jl:852-982; remote GIDs ordered by owner PID, then GID), makeIndicesLocal (:636-712),
 * sortAndMergeIndicesAndValues (src/CSRMatrix.jl:549-606; duplicates summed in insertion order) and
 * fillLocalGraphAndMatrix (:403-484).  entry_ptr[n_rows+1] are 0-based offsets of every row's entries in insertion
 * order, col_gids / vals the entries; [min_all_gid, max_all_gid] is the domain map's global id range and the calling
 * rank owns dom_num_my ids from dom_min_gid.  info[4] = columns of the column map, used local columns, numSameIDs of
 * Import(domainMap, colMap), nnz after the merge.  The result is an ordinary CSR handle (n_local_cols = numSameIDs). */
int32_t jp_csr_fill_complete(int32_t n_rows, const int64_t* entry_ptr, const int64_t* col_gids, const double* vals,
                             int64_t min_all_gid, int64_t max_all_gid, int64_t dom_min_gid, int32_t dom_num_my,
                             jp_handle* out, int64_t* info);
/* GIDs of the column map in LID order (n_cols of them) of a matrix made by jp_csr_fill_complete */
int32_t jp_csr_col_map(jp_handle A, int64_t* gids);

/* ---- DenseMultiVector storage (src/DenseMultiVector.jl:6-60) and the MultiVector
 *      kernels (src/MultiVector.jl:52-161) -------------------------------------------- */
int32_t jp_mv_create(int32_t n_local, int32_t k, jp_handle* out);           /* zero-filled */
int32_t jp_mv_destroy(jp_handle mv);
int32_t jp_mv_shape(jp_handle mv, int32_t* n_local, int32_t* k);
int32_t jp_mv_upload(jp_handle mv, const double* host);                     /* n_local*k doubles, column-major */
int32_t jp_mv_download(jp_handle mv, double* host);
int32_t jp_mv_copy(jp_handle dst, jp_handle src);                           /* copyto!, src/DenseMultiVector.jl:54-60 */
int32_t jp_mv_fill(jp_handle mv, double value);                             /* fill!, src/MultiVector.jl:52-55 */
int32_t jp_mv_scale(jp_handle mv, double alpha);                            /* scale!, src/MultiVector.jl:62-65 */
int32_t jp_mv_scale_vec(jp_handle mv, const double* alpha_k);               /* scale! per vector, :72-77 */
int32_t jp_mv_axpby(jp_handle y, double a, jp_handle x, double b);          /* y = a*x + b*y; broadcast copyto!, :497-502 */
/* y = a*x + b*y and norm(y, 2) of the result in one pass (the residual update of a CG iteration + its norm):
 * broadcast copyto! (:497-502) fused with norm (:122-133); out_k as jp_mv_norm2 */
int32_t jp_mv_axpby_norm2(jp_handle comm, jp_handle y, double a, jp_handle x, double b, double* out_k);
int32_t jp_mv_dot(jp_handle comm, jp_handle x, jp_handle y, double* out_k);    /* dot, :84-108 (allreduce included) */
int32_t jp_mv_norm2(jp_handle comm, jp_handle x, double* out_k);               /* norm(.,2), :122-133 */
int32_t jp_mv_normp(jp_handle comm, jp_handle x, double p, double* out_k);     /* norm(.,p), :135-144 */
int32_t jp_mv_comm_reduce(jp_handle comm, jp_handle mv);                       /* commReduce, :154-161 */

/* ---- device-resident scalars for solver loops (no reference counterpart: the reference returns every dot / norm to
 *      the host, src/MultiVector.jl:84-146, which costs two host synchronisations per CG iteration).  A `scalars` object
 *      is a small device array; the *_dev calls leave their (all-reduced) result in one of its slots and take update
 *      coefficients as  scale * s[num] / s[den]  (num / den = -1: factor 1), evaluated on the device.  Single vectors
 *      (k = 1).  Everything is stream-ordered; only jp_scalars_download synchronises. ------------------------------- */
int32_t jp_scalars_create(int32_t n, jp_handle* out);                       /* n doubles, zeroed */
int32_t jp_scalars_destroy(jp_handle s);
int32_t jp_scalars_upload(jp_handle s, const double* host);
int32_t jp_scalars_download(jp_handle s, double* host);
int32_t jp_scalars_copy(jp_handle s, int32_t dst, int32_t src);            /* s[dst] = s[src] */
int32_t jp_mv_dot_dev(jp_handle comm, jp_handle x, jp_handle y, jp_handle s, int32_t slot);   /* s[slot] = dot(x, y) */
/* y = a*x + b*y with a = a_scale * s[a_num] / s[a_den], b likewise (the arithmetic of jp_mv_axpby) */
int32_t jp_mv_axpby_dev(jp_handle y, jp_handle x, jp_handle s, double a_scale, int32_t a_num, int32_t a_den, double b_scale,
                        int32_t b_num, int32_t b_den);
/* the same update and s[out_slot] = sum of squares of the result (norm(y, 2)^2, all-reduced) in one pass */
int32_t jp_mv_axpby_norm2sq_dev(jp_handle comm, jp_handle y, jp_handle x, jp_handle s, double a_scale, int32_t a_num, int32_t a_den,
                                double b_scale, int32_t b_num, int32_t b_den, int32_t out_slot);
/* apply!(Y, A, X, NO_TRANS, alpha, beta) and s[slot] = dot(W, Y) (jp_apply_dot without the host round trip) */
int32_t jp_apply_dot_dev(jp_handle comm, jp_handle Y, jp_handle A, jp_handle X, jp_handle plan, double alpha, double beta,
                         jp_handle W, jp_handle s, int32_t slot);

/* ---- Import / Export plan + its Distributor plan on the device
 *      (src/ImportExportData.jl:1-16; src/MPIDistributor.jl:12-65).  export_lids must be
 *      in send order (grouped by procs_to; compose with indices_to on the host when the
 *      distributor is not blocked).  procs_*, lengths_* include the self message. ------- */
int32_t jp_plan_create(jp_handle comm, int32_t n_src, int32_t n_tgt, int32_t num_same, const int32_t* permute_to,
                       const int32_t* permute_from, int32_t n_permute, const int32_t* remote_lids, int32_t n_remote,
                       const int32_t* export_lids, int32_t n_export, const int32_t* procs_to, const int64_t* lengths_to,
                       int32_t n_sends, const int32_t* procs_from, const int64_t* lengths_from, int32_t n_recvs,
                       int32_t index_base, jp_handle* out);
int32_t jp_plan_destroy(jp_handle plan);
/* doImport / doExport -> doTransfer (src/DistObject.jl:86-164, 200-258): copyAndPermute, pack,
 * exchange, unpack.  reverse != 0 runs the distributor in reverse with the same lists, as the
 * reference does. */
int32_t jp_transfer(jp_handle plan, jp_handle src, jp_handle tgt, int32_t combine_mode, int32_t reverse);
/* resolvePosts / resolveWaits split (src/Distributor.jl:24-29): post = copyAndPermute + pack + sends/recvs on
 * the comm stream; wait = unpack ordered after them.  Waiting without a post is JP_INVALID_STATE. */
int32_t jp_transfer_post(jp_handle plan, jp_handle src, jp_handle tgt, int32_t combine_mode, int32_t reverse);
int32_t jp_transfer_wait(jp_handle plan);

/* ---- Operator (src/Operator.jl:35-88) -------------------------------------------------- */
/* localApply(Y, A, X, mode, alpha, beta), src/CSRMatrix.jl:1123-1162.  X is a column-map MultiVector
 * (n_cols rows) for NO_TRANS, a row-map one for TRANS. */
int32_t jp_local_apply(jp_handle Y, jp_handle A, jp_handle X, int32_t mode, double alpha, double beta);
/* apply!(Y, A, X, mode, alpha, beta), src/RowMatrix.jl:261-357, with the Import halo (plan, or 0 when the
 * importer is `nothing`) inside or overlapped with the interior rows.  NO_TRANS: X on the domain map, Y on the row
 * map.  TRANS / CONJ_TRANS: X on the row map, Y on the domain map; with a plan the exchange runs backwards and ADDs
 * (the semantics of src/RowMatrix.jl:324-352, i.e. Tpetra's applyTranspose).  Collective over the plan's comm. */
int32_t jp_apply(jp_handle comm, jp_handle Y, jp_handle A, jp_handle X, jp_handle plan, int32_t mode, double alpha,
                 double beta);
/* apply! of a matrix whose range map differs from its row map (the exporter branch, src/RowMatrix.jl:298-308, 324-330,
 * with the semantics of Tpetra::CrsMatrix::apply that those lines transcribe; the reference's own lines cannot run).
 * export_plan is the plan of Export(rowMap, rangeMap).  NO_TRANS: the product is formed on the row map and ADDED into
 * Y (range map) -- rows held by several ranks are summed; TRANS: X (range map) is first brought onto the row map.
 * import_plan as in jp_apply (0 when the importer is `nothing`).  Collective over the plans' comm. */
int32_t jp_apply_export(jp_handle comm, jp_handle Y, jp_handle A, jp_handle X, jp_handle import_plan, jp_handle export_plan,
                        int32_t mode, double alpha, double beta);
/* apply! followed by dot(W, Y) (src/MultiVector.jl:84-108) without a host round trip in between; for k = 1, NO_TRANS
 * and no importer the dot product is accumulated by the SpMV kernel itself (p'Ap of a CG iteration: W = X). */
int32_t jp_apply_dot(jp_handle comm, jp_handle Y, jp_handle A, jp_handle X, jp_handle plan, int32_t mode, double alpha,
                     double beta, jp_handle W, double* out_k);
/* the same call with HOST buffers (what a Julia caller holding Array{Float64,2} makes): uploads X, applies,
 * downloads Y.  y_host is also read when beta != 0. */
int32_t jp_apply_host(jp_handle comm, jp_handle A, jp_handle plan, const double* x_host, double* y_host, int32_t k,
                      int32_t mode, double alpha, double beta);

#ifdef __cplusplus
}
#endif
#endif /* JPETRA_B200_H */
