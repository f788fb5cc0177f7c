// jp_runtime.cu -- process context (device, streams), handle registry, events, pinned staging
#include "jp_common.h"

namespace jp {

static thread_local std::string t_last_error;
static std::string g_last_error;  // last error of any thread (read by hosts that call from a different thread)
static std::mutex g_err_mu;

void set_last_error(const std::string& m) {
  t_last_error = m;
  std::lock_guard<std::mutex> lk(g_err_mu);
  g_last_error = m;
}

static Context g_ctx;
Context& ctx() { return g_ctx; }

void require_init() {
  if (!g_ctx.initialised) fail(JP_INVALID_STATE, "jp_init has not been called");
}

void* scratch(size_t bytes) {
  Context& c = ctx();
  if (bytes > c.scratch_bytes) {
    JP_CUDA(cudaStreamSynchronize(c.compute));
    JP_CUDA(cudaStreamSynchronize(c.comm));
    if (c.scratch) cudaFree(c.scratch);
    size_t want = bytes < (1u << 20) ? (1u << 20) : bytes;
    JP_CUDA(cudaMalloc(&c.scratch, want));
    c.scratch_bytes = want;
  }
  return c.scratch;
}

void* host_scratch(size_t bytes) {
  Context& c = ctx();
  if (bytes > c.host_scratch_bytes) {
    JP_CUDA(cudaStreamSynchronize(c.compute));
    JP_CUDA(cudaStreamSynchronize(c.comm));
    if (c.host_scratch) cudaFreeHost(c.host_scratch);
    size_t want = bytes < (1u << 16) ? (1u << 16) : bytes;
    JP_CUDA(cudaMallocHost(&c.host_scratch, want));
    c.host_scratch_bytes = want;
  }
  return c.host_scratch;
}

unsigned long long* p2p_timeout_flag_device() {
  Context& c = ctx();
  if (!c.p2p_timeout_host) {
    JP_CUDA(cudaHostAlloc((void**)&c.p2p_timeout_host, 64, cudaHostAllocMapped));
    *c.p2p_timeout_host = 0;
    JP_CUDA(cudaHostGetDevicePointer((void**)&c.p2p_timeout_dev, c.p2p_timeout_host, 0));
  }
  return c.p2p_timeout_dev;
}

void retire_arena(DeviceBuffer& arena) {
  if (!arena.p) return;
  ctx().retired_arenas.push_back(arena.p);
  arena.p = nullptr;
  arena.bytes = 0;
}

void check_p2p_health() {
  Context& c = ctx();
  if (c.p2p_timeout_host && *reinterpret_cast<volatile unsigned long long*>(c.p2p_timeout_host) != 0) {
    *c.p2p_timeout_host = 0;
    fail(JP_CUDA_ERROR, "peer-to-peer halo exchange timed out waiting for a neighbour rank (20 s): results of the last apply! are invalid");
  }
}

// ---- registry -------------------------------------------------------------------------------------
static std::mutex g_reg_mu;
static std::unordered_map<jp_handle, std::unique_ptr<Object>> g_objects;
static jp_handle g_next = 1000;

jp_handle register_object(std::unique_ptr<Object> obj) {
  std::lock_guard<std::mutex> lk(g_reg_mu);
  jp_handle h = g_next++;
  g_objects[h] = std::move(obj);
  return h;
}

Object* lookup(jp_handle h, Kind kind, const char* what) {
  std::lock_guard<std::mutex> lk(g_reg_mu);
  auto it = g_objects.find(h);
  if (it == g_objects.end()) fail(JP_INVALID_ARGUMENT, std::string("invalid ") + what + " handle " + std::to_string(h));
  if (it->second->kind != kind) fail(JP_INVALID_ARGUMENT, std::string("handle ") + std::to_string(h) + " is not a " + what);
  return it->second.get();
}

void destroy_object(jp_handle h, Kind kind, const char* what) {
  std::unique_ptr<Object> victim;
  {
    std::lock_guard<std::mutex> lk(g_reg_mu);
    auto it = g_objects.find(h);
    if (it == g_objects.end()) fail(JP_INVALID_ARGUMENT, std::string("invalid ") + what + " handle " + std::to_string(h));
    if (it->second->kind != kind) fail(JP_INVALID_ARGUMENT, std::string("handle ") + std::to_string(h) + " is not a " + what);
    victim = std::move(it->second);
    g_objects.erase(it);
  }
  // in-flight work may still reference the object's buffers
  if (ctx().initialised) {
    cudaStreamSynchronize(ctx().compute);
    cudaStreamSynchronize(ctx().comm);
    cudaStreamSynchronize(ctx().copy);
  }
  victim.reset();
}

void destroy_all_objects() {
  std::unordered_map<jp_handle, std::unique_ptr<Object>> all;
  {
    std::lock_guard<std::mutex> lk(g_reg_mu);
    all.swap(g_objects);
  }
  // comms last: plans hold raw Comm pointers
  for (auto it = all.begin(); it != all.end();) {
    if (it->second->kind != Kind::Comm)
      it = all.erase(it);
    else
      ++it;
  }
  all.clear();
}

}  // namespace jp

using namespace jp;

extern "C" {

int32_t jp_version(void) { return 1; }

int32_t jp_last_error(char* buf, int64_t len) {
  if (!buf || len <= 0) return JP_INVALID_ARGUMENT;
  std::string m = t_last_error;
  if (m.empty()) {
    std::lock_guard<std::mutex> lk(g_err_mu);
    m = g_last_error;
  }
  size_t n = (size_t)len - 1 < m.size() ? (size_t)len - 1 : m.size();
  std::memcpy(buf, m.data(), n);
  buf[n] = 0;
  return JP_OK;
}

int32_t jp_init(int32_t device) {
  return guarded([&] {
    Context& c = ctx();
    if (c.initialised) {
      JP_REQUIRE_STATE(c.device == device, "jp_init: already bound to device " + std::to_string(c.device));
      return;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
      fail(JP_CUDA_ERROR, std::string("no CUDA device available (") + (e == cudaSuccess ? "device count 0" : cudaGetErrorString(e)) +
                              "); this backend has no CPU fallback");
    JP_REQUIRE(device >= 0 && device < count, "jp_init: device index out of range");
    JP_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    JP_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
      fail(JP_CUDA_ERROR, std::string("device ") + prop.name + " is sm_" + std::to_string(prop.major * 10 + prop.minor) +
                              "; the kernels are built for sm_100a only");
    c.device = device;
    c.sm_count = prop.multiProcessorCount;
    JP_CUDA(cudaStreamCreateWithFlags(&c.compute, cudaStreamNonBlocking));
    int lo = 0, hi = 0;
    JP_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    JP_CUDA(cudaStreamCreateWithPriority(&c.comm, cudaStreamNonBlocking, hi));  // halo first: it gates the boundary rows
    JP_CUDA(cudaStreamCreateWithFlags(&c.copy, cudaStreamNonBlocking));
    JP_CUDA(cudaEventCreateWithFlags(&c.ev_x_ready, cudaEventDisableTiming));
    JP_CUDA(cudaEventCreateWithFlags(&c.ev_halo_ready, cudaEventDisableTiming));
    JP_CUDA(cudaEventCreateWithFlags(&c.ev_tmp, cudaEventDisableTiming));
    c.initialised = true;
  });
}

int32_t jp_finalize(void) {
  return guarded([&] {
    Context& c = ctx();
    if (!c.initialised) return;
    cudaStreamSynchronize(c.compute);
    cudaStreamSynchronize(c.comm);
    cudaStreamSynchronize(c.copy);
    destroy_all_objects();
    release_host_path_cache();
    release_reduce_workspace();
    if (c.scratch) cudaFree(c.scratch);
    if (c.host_scratch) cudaFreeHost(c.host_scratch);
    if (c.flush_buf) cudaFree(c.flush_buf);
    if (c.p2p_timeout_host) cudaFreeHost(c.p2p_timeout_host);
    c.p2p_timeout_host = c.p2p_timeout_dev = nullptr;
    for (void* a : c.retired_arenas) cudaFree(a);  // every plan is gone: nobody publishes into a peer's arena any more
    c.retired_arenas.clear();
    c.scratch = c.host_scratch = c.flush_buf = nullptr;
    c.scratch_bytes = c.host_scratch_bytes = c.flush_bytes = 0;
    cudaEventDestroy(c.ev_x_ready);
    cudaEventDestroy(c.ev_halo_ready);
    cudaEventDestroy(c.ev_tmp);
    cudaStreamDestroy(c.compute);
    cudaStreamDestroy(c.comm);
    cudaStreamDestroy(c.copy);
    c.initialised = false;
    c.device = -1;
  });
}

int32_t jp_device_info(int32_t* sm_count, int64_t* hbm_bytes, int32_t* cc_major, int32_t* cc_minor) {
  return guarded([&] {
    require_init();
    cudaDeviceProp prop;
    JP_CUDA(cudaGetDeviceProperties(&prop, ctx().device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (hbm_bytes) *hbm_bytes = (int64_t)prop.totalGlobalMem;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
  });
}

int32_t jp_synchronize(void) {
  return guarded([&] {
    require_init();
    JP_CUDA(cudaStreamSynchronize(ctx().comm));
    JP_CUDA(cudaStreamSynchronize(ctx().compute));
    JP_CUDA(cudaStreamSynchronize(ctx().copy));
    check_p2p_health();
  });
}

int32_t jp_host_alloc(int64_t bytes, void** ptr) {
  return guarded([&] {
    require_init();
    JP_REQUIRE(ptr != nullptr && bytes >= 0, "jp_host_alloc: bad arguments");
    JP_CUDA(cudaMallocHost(ptr, bytes > 0 ? (size_t)bytes : 16));
  });
}

int32_t jp_host_free(void* ptr) {
  return guarded([&] {
    require_init();
    if (ptr) JP_CUDA(cudaFreeHost(ptr));
  });
}

int32_t jp_event_create(jp_handle* ev) {
  return guarded([&] {
    require_init();
    JP_REQUIRE(ev != nullptr, "jp_event_create: null output");
    auto e = std::make_unique<Event>();
    JP_CUDA(cudaEventCreate(&e->ev));
    *ev = register_object(std::move(e));
  });
}

int32_t jp_event_record(jp_handle ev) {
  return guarded([&] {
    require_init();
    JP_CUDA(cudaEventRecord(get<Event>(ev, Kind::Event, "event")->ev, ctx().compute));
  });
}

int32_t jp_event_elapsed_ms(jp_handle start, jp_handle stop, double* ms) {
  return guarded([&] {
    require_init();
    JP_REQUIRE(ms != nullptr, "jp_event_elapsed_ms: null output");
    Event* a = get<Event>(start, Kind::Event, "event");
    Event* b = get<Event>(stop, Kind::Event, "event");
    JP_CUDA(cudaEventSynchronize(b->ev));
    float f = 0;
    JP_CUDA(cudaEventElapsedTime(&f, a->ev, b->ev));
    *ms = (double)f;
  });
}

int32_t jp_event_destroy(jp_handle ev) {
  return guarded([&] { destroy_object(ev, Kind::Event, "event"); });
}

int32_t jp_launch_count(int64_t* n) {
  return guarded([&] {
    JP_REQUIRE(n != nullptr, "jp_launch_count: null output");
    *n = ctx().launches.load();
  });
}

int32_t jp_flush_l2(int64_t bytes) {
  return guarded([&] {
    require_init();
    Context& c = ctx();
    size_t want = bytes > 0 ? (size_t)bytes : (size_t)256 << 20;
    if (want > c.flush_bytes) {
      if (c.flush_buf) cudaFree(c.flush_buf);
      JP_CUDA(cudaMalloc(&c.flush_buf, want));
      c.flush_bytes = want;
    }
    JP_CUDA(cudaMemsetAsync(c.flush_buf, 0, want, c.compute));
  });
}

}  // extern "C"
