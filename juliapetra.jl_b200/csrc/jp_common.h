// jp_common.h -- internal runtime shared by the .cu files of libjpetra_b200.so
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/jpetra_b200.h"
#include "jp_blockdesc.h"

namespace jp {

struct Error {
  int32_t code;
  std::string msg;
};

void set_last_error(const std::string& m);

[[noreturn]] inline void fail(int32_t code, const std::string& m) { throw Error{code, m}; }

#define JP_CUDA(call)                                                                                         \
  do {                                                                                                        \
    cudaError_t e__ = (call);                                                                                 \
    if (e__ != cudaSuccess)                                                                                   \
      ::jp::fail(JP_CUDA_ERROR, std::string(#call) + " failed: " + cudaGetErrorString(e__) + " (" __FILE__ ":" + \
                                    std::to_string(__LINE__) + ")");                                          \
  } while (0)

#define JP_NCCL(call)                                                                                         \
  do {                                                                                                        \
    ncclResult_t r__ = (call);                                                                                \
    if (r__ != ncclSuccess)                                                                                   \
      ::jp::fail(JP_NCCL_ERROR, std::string(#call) + " failed: " + ncclGetErrorString(r__) + " (" __FILE__ ":" + \
                                    std::to_string(__LINE__) + ")");                                          \
  } while (0)

#define JP_REQUIRE(cond, msg)                                  \
  do {                                                         \
    if (!(cond)) ::jp::fail(JP_INVALID_ARGUMENT, (msg));       \
  } while (0)

#define JP_REQUIRE_STATE(cond, msg)                         \
  do {                                                      \
    if (!(cond)) ::jp::fail(JP_INVALID_STATE, (msg));       \
  } while (0)

// ---- runtime context: one per process (one process per GPU) -------------------------------
struct Context {
  bool initialised = false;
  int device = -1;
  int sm_count = 0;
  cudaStream_t compute = nullptr;  // kernels of the hot path
  cudaStream_t comm = nullptr;     // pack + NCCL send/recv, overlapped with interior rows
  cudaStream_t copy = nullptr;     // D2H of results in the host-buffer path
  cudaEvent_t ev_x_ready = nullptr, ev_halo_ready = nullptr, ev_tmp = nullptr;
  std::atomic<int64_t> launches{0};
  void* scratch = nullptr;  // device staging for host-array collectives / reductions
  size_t scratch_bytes = 0;
  void* host_scratch = nullptr;  // pinned
  size_t host_scratch_bytes = 0;
  void* flush_buf = nullptr;
  size_t flush_bytes = 0;
  // host-mapped word the peer-to-peer halo kernels increment when a flag wait times out
  unsigned long long* p2p_timeout_host = nullptr;
  unsigned long long* p2p_timeout_dev = nullptr;
  std::vector<void*> retired_arenas;  // IPC-exported halo arenas of destroyed / regrown plans (a peer may still map them)
};
Context& ctx();
void require_init();
void* scratch(size_t bytes);
void* host_scratch(size_t bytes);
unsigned long long* p2p_timeout_flag_device();  // allocates the mapped word on first use
void check_p2p_health();                        // throws JP_CUDA_ERROR when a peer-to-peer halo wait has timed out

inline void count_launch(int n = 1) { ctx().launches.fetch_add(n, std::memory_order_relaxed); }

#define JP_LAUNCH_CHECK()            \
  do {                               \
    ::jp::count_launch();            \
    JP_CUDA(cudaGetLastError());     \
  } while (0)

// Kernels that live in *_kernels.cuh headers are written so that tests/cuda_emul can execute the same source on the
// CPU; these two macros are the only launch-side syntax that differs between the two builds.
#define JP_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#define JP_KLAUNCH(kernel, grid, block, smem, stream, ...)         \
  do {                                                             \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);    \
    JP_LAUNCH_CHECK();                                             \
  } while (0)

// ---- handle registry ------------------------------------------------------------------------
enum class Kind : int { Comm = 1, Csr, MultiVector, Plan, Event, Scalars };

struct Object {
  Kind kind;
  explicit Object(Kind k) : kind(k) {}
  virtual ~Object() {}
};

jp_handle register_object(std::unique_ptr<Object> obj);
Object* lookup(jp_handle h, Kind kind, const char* what);
void destroy_object(jp_handle h, Kind kind, const char* what);
void destroy_all_objects();

template <class T>
T* get(jp_handle h, Kind kind, const char* what) {
  return static_cast<T*>(lookup(h, kind, what));
}

template <class F>
int32_t guarded(F&& f) {
  try {
    f();
    return JP_OK;
  } catch (const Error& e) {
    set_last_error(e.msg);
    return e.code;
  } catch (const std::exception& e) {
    set_last_error(std::string("internal error: ") + e.what());
    return JP_CUDA_ERROR;
  }
}

// ---- objects ----------------------------------------------------------------------------------
struct DeviceBuffer {
  void* p = nullptr;
  size_t bytes = 0;
  DeviceBuffer() {}
  DeviceBuffer(const DeviceBuffer&) = delete;
  DeviceBuffer& operator=(const DeviceBuffer&) = delete;
  ~DeviceBuffer() { release(); }
  void alloc(size_t n) {
    release();
    if (n == 0) n = 16;
    JP_CUDA(cudaMalloc(&p, n));
    bytes = n;
  }
  void ensure(size_t n) {
    if (n > bytes) alloc(n);
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  void swap(DeviceBuffer& o) {
    std::swap(p, o.p);
    std::swap(bytes, o.bytes);
  }
  template <class T>
  T* as() const { return static_cast<T*>(p); }
};

struct Comm : Object {
  Comm() : Object(Kind::Comm) {}
  ~Comm() override;
  bool serial = true;
  int nproc = 1;
  int rank0 = 0;  // 0-based
  ncclComm_t nccl = nullptr;
  // ---- small allreduces over peer memory (jp_comm.cu): every rank exposes [flags | 2 x nproc x kSmallAllreduceMax
  //      doubles] through CUDA IPC; one kernel stores this rank's values into every peer's slots, waits for everybody's
  //      flag and adds the slots up in rank order.  ~5 us instead of ~25 us for an NCCL allreduce of a few doubles.
  int ar_state = 0;  // 0 = not tried, 1 = active, -1 = unavailable (NCCL is used)
  DeviceBuffer ar_arena, ar_peers;  // my arena; device array of the nproc mapped arena bases
  std::vector<void*> ar_peer_host;
  uint64_t ar_epoch = 0;
};

struct MultiVector : Object {
  MultiVector() : Object(Kind::MultiVector) {}
  int32_t n = 0, k = 0;  // column-major, ld == n
  DeviceBuffer data;
  double* d() const { return data.as<double>(); }
};


struct DiaData;  // jp_dia.cu
struct DiaDeleter {
  void operator()(DiaData* p) const;
};

struct CbData;  // jp_colblock.cu
struct CbDeleter {
  void operator()(CbData* p) const;
};

struct Csr : Object {
  Csr() : Object(Kind::Csr) {}
  int32_t n_rows = 0, n_cols = 0, n_local_cols = 0;
  int64_t nnz = 0;
  int32_t max_row_len = 0;
  int32_t threads_per_row = 1;  // chosen from the row-length statistics at create time
  int32_t nnz_cap = 2048;       // nnz staged in shared memory per block
  DeviceBuffer rowptr, colind, vals;  // 0-based; colind/vals padded for 16-byte over-reads
  // row blocks (BlockDesc); interior = no remote column
  DeviceBuffer blocks_interior, blocks_boundary, blocks_all;
  int32_t n_blocks_interior = 0, n_blocks_boundary = 0, n_blocks_all = 0;
  std::vector<BlockDesc> h_blocks_interior, h_blocks_boundary, h_blocks_all;  // host copies (the tile builder patches them)
  // SpMM runs path (spmm_runs_kernel; built lazily at the first k > 1 apply): its own row blocks (<= 64 rows), the
  // aligned column runs of every block and the uint16 tile position of every nonzero
  int runs_state = 0;        // 0 = not built, 1 = active, -1 = not usable / disabled
  int32_t runs_wmax = 0, runs_cap = 0;
  DeviceBuffer runs, runs_lidx, runs_blocks_all, runs_blocks_interior;
  int32_t n_runs_blocks_all = 0, n_runs_blocks_interior = 0;
  // cache of the column-map multivector (A.importMV, src/CSRMatrix.jl:10,772-781); only used when the column
  // map is not [domain prefix | remote tail] and the fused apply path does not apply
  std::unique_ptr<MultiVector> import_mv;
  // SpMM on diagonal-structured matrices (spmm_dia_kernel; built lazily at the first k > 1 apply, jp_dia.cu)
  int dia_state = 0;  // 0 = not decided, 1 = active, -1 = not a diagonal-structured matrix / disabled
  std::unique_ptr<DiaData, DiaDeleter> dia;
  // SpMV (k = 1) of a large irregular matrix through column slabs (spmv_cb_kernel; built lazily, jp_colblock.cu)
  int cb_state = 0;  // 0 = not decided, 1 = active, -1 = not that kind of matrix / disabled
  std::unique_ptr<CbData, CbDeleter> cb;
  // A^T as a device CSR of its own, built at the first transposed apply (jp_transpose.cu)
  std::unique_ptr<Csr> transposed;
  // cache of the row-map multivector (A.exportMV, src/CSRMatrix.jl:11): the exporter branch of apply! (rangeMap != rowMap)
  std::unique_ptr<MultiVector> export_mv;
  // GIDs of the column map in LID order, kept when the matrix was fill-completed on the device (jp_csr_fill_complete)
  DeviceBuffer colmap_gids;
  // jp_apply_dot workspace: [block partials | result]
  DeviceBuffer dot_ws;
  // host-buffer path (jp_apply_host): the all-blocks list cut into chunks; chunk c covers blocks
  // [chunk_block[c], chunk_block[c+1]) = rows [chunk_row[c], chunk_row[c+1]) and reads columns <= chunk_maxcol[c]
  // (prefix maximum), so its kernel can start as soon as that much of X has been uploaded
  std::vector<int32_t> chunk_block, chunk_row, chunk_maxcol;
  // the same for the host-buffer path WITH an importer: chunks of the INTERIOR list; chunk c covers interior blocks
  // [ichunk_block[c], ichunk_block[c+1]) and owns the rows [ichunk_row[c], ichunk_row[c+1]) of Y -- the boundary rows that
  // lie in that range included (ichunk_has_boundary[c]); its interior rows read local columns <= ichunk_maxcol[c]
  std::vector<int32_t> ichunk_block, ichunk_row, ichunk_maxcol;
  std::vector<uint8_t> ichunk_has_boundary;
};

struct Plan : Object {
  Plan() : Object(Kind::Plan) {}
  Comm* comm = nullptr;
  int32_t n_src = 0, n_tgt = 0, num_same = 0, n_permute = 0, n_remote = 0, n_export = 0;
  DeviceBuffer permute_to, permute_from, remote_lids, export_lids;  // 0-based int32
  std::vector<int32_t> procs_to, procs_from;                        // 1-based PIDs, incl. self
  std::vector<int64_t> lengths_to, lengths_from, starts_to, starts_from;  // element offsets (0-based)
  bool remote_is_tail = false;     // remote_lids == num_same + (0..n_remote-1): receives land in place
  bool export_contiguous = false;  // every send block is a contiguous LID range (k==1 needs no pack)
  DeviceBuffer sendbuf, recvbuf;   // LID-major, vector-minor (src/MultiVector.jl:335-349)
  int32_t buf_k = 0;
  // post/wait state (src/Distributor.jl:24-29)
  bool posted = false;
  jp_handle posted_tgt = 0;
  int32_t posted_k = 0;
  cudaEvent_t ev_done = nullptr;
  // ---- peer-to-peer halo (protocol and arena layout: jp_p2p.cuh) ----
  int p2p_state = 0;            // 0 = not tried, 1 = active, -1 = unavailable (NCCL send/recv is used)
  int32_t p2p_kcap = 0;         // vectors one halo buffer has room for
  DeviceBuffer arena;
  size_t arena_buf_bytes = 0;   // bytes of ONE of my two halo buffers
  std::vector<void*> peer_arena;   // per rank: mapped base of its arena (own rank: arena.p; non-neighbours: nullptr)
  DeviceBuffer send_table;      // P2PSend per send block
  DeviceBuffer wait_table;      // per source: {const uint64* ready flag (mine)}, {uint64* ack flag (in the source's arena)}
  int32_t n_send_blocks = 0, n_sources = 0;
  DeviceBuffer p2p_tickets;     // uint32 counters: [0] pack items done, [1] boundary blocks done, [2] pack-item claims
  uint64_t epoch = 0;           // applies done through the peer-memory halo (identical on every rank)
  ~Plan() override;
};

struct Event : Object {
  Event() : Object(Kind::Event) {}
  cudaEvent_t ev = nullptr;
  ~Event() override {
    if (ev) cudaEventDestroy(ev);
  }
};

// ---- cross-file entry points (internal) ---------------------------------------------------------
// y[r0:r1) rows of A applied to x (+ optional halo) on `stream`
void launch_spmv_blocks(const Csr& A, const void* blocks_int4, int32_t n_blocks, bool use_halo, const double* x, int64_t ldx,
                        const double* halo, int32_t halo_k, double* y, int64_t ldy, int32_t k, double alpha, double beta,
                        cudaStream_t stream);
void launch_spmv(const Csr& A, const DeviceBuffer& blocks, int32_t n_blocks, bool use_halo, const double* x, int64_t ldx,
                 const double* halo, int32_t halo_k, double* y, int64_t ldy, int32_t k, double alpha, double beta,
                 cudaStream_t stream);
Csr& csr_transposed(Csr& A);  // builds A.transposed on first use
void launch_spmv_trans(Csr& A, const double* x, int64_t ldx, double* y, int64_t ldy, int32_t ny, int32_t k, double alpha,
                       double beta, cudaStream_t stream);
void csr_prepare_spmm(Csr& A);  // decides (once) the SpMM path of the matrix: diagonal copy, run tables, or the gather kernel
void csr_prepare_dia(Csr& A);   // decides (once) whether the matrix gets a diagonal-major copy for spmm_dia_kernel
void csr_prepare_colblock(Csr& A);  // decides (once) whether the matrix gets a column-slab copy for spmv_cb_kernel
void launch_spmv_colblock(const Csr& A, const double* x, double* y, double alpha, double beta, cudaStream_t stream);
int64_t csr_colblock_info(const Csr& A, int what);  // 0: slabs, 1: columns per slab
void launch_spmm_dia(const Csr& A, const double* x, int64_t ldx, double* y, int64_t ldy, int32_t k, double alpha, double beta,
                     cudaStream_t stream);
// the gather SpMM kernel on a contiguous piece of a block list (k > 1, no halo columns)
void launch_spmm_gather_blocks(const Csr& A, const void* blocks, int32_t n_blocks, const double* x, int64_t ldx, double* y, int64_t ldy,
                               int32_t k, double alpha, double beta, cudaStream_t stream);
int64_t csr_dia_info(const Csr& A, int what);  // 0: diagonals, 1: rows served by the diagonal kernel, 2: tile width
void csr_finalize_structure(Csr& A, const std::vector<int32_t>& rp, const std::vector<int32_t>& rowmax);
void launch_scale_fill(double* y, int64_t n, double beta, cudaStream_t stream);  // y = beta==0 ? 0 : beta*y
void plan_ensure_buffers(Plan& p, int32_t k);
void plan_post(Plan& p, const MultiVector& src, MultiVector* tgt, bool halo_only, int32_t combine_mode, bool reverse);
void plan_unpack(Plan& p, MultiVector& tgt, cudaStream_t stream);
struct FusedHalo;
void launch_spmv_fused(const Csr& A, const FusedHalo& h, const double* x, const double* halo, double* y, double alpha, double beta,
                       cudaStream_t stream);
double* launch_spmv_fused_dot(Csr& A, const FusedHalo& h, const double* x, const double* halo, double* y, double alpha, double beta,
                              const double* w, cudaStream_t stream);
const double* plan_post_p2p(Plan& p, const MultiVector& src);
void plan_p2p_ack(Plan& p);
bool plan_p2p_ensure(Plan& p, int32_t k);
void retire_arena(DeviceBuffer& arena);  // an IPC-exported arena is parked until jp_finalize instead of freed
double* apply_dot_local(Csr* a, Plan* p, MultiVector* y, MultiVector* x, MultiVector* w, int32_t mode, double alpha, double beta);
void apply_trans_distributed(Csr* a, Plan* p, MultiVector* y, MultiVector* x, double alpha, double beta);
// where a reduction's result goes besides the device buffer: host-mapped memory + a sequence number (jp_mv.cu)
struct HostSink {
  double* out = nullptr;
  unsigned long long* flag = nullptr;
  unsigned long long seq = 0;
};
// sumAll of a device buffer in place; returns true when the values were also handed to `sink` (by the same kernel)
bool comm_allreduce_device(Comm& c, double* buf, int64_t n, cudaStream_t stream, const HostSink* sink = nullptr);
void allgather_host_bytes(Comm& c, const void* in, size_t bytes, void* out);  // setup-time helpers (comm stream, synchronising)
int64_t allreduce_min_host(Comm& c, int64_t v);
void release_host_path_cache();
void release_reduce_workspace();
// k local dot products -> device workspace, or straight into `out` (k doubles on the device) when given (compute stream)
double* mv_dot_local(const MultiVector& x, const MultiVector& y, double* out = nullptr);
void mv_finish_reduction(Comm& c, double* dev_out, int32_t k, double* host_out);  // allreduce + D2H + synchronise
// y = beta*y + alpha*A*x over all row blocks (x = the whole column-map vector) and sum_r w[r]*y[r] -> returned device slot
double* launch_spmv_dot(Csr& A, const double* x, double* y, double alpha, double beta, const double* w, cudaStream_t stream);

}  // namespace jp
