// onk_runtime.cu -- context, lanes (streams), memory, parallel-operation bracket, generic launch, graphs.
// Replaces the host-side CUDA glue of the reference: src/cuda/cuda_context.cpp, src/cuda/allocator.cpp,
// the CUDA branch of src/parallel/parallel_execution_queue.cpp:316-358 and ONIKA_CU_LAUNCH_KERNEL.
#include "onk_internal.cuh"

#include <cuda.h>
#include <cstdarg>
#include <condition_variable>
#include <thread>

namespace onk
{
  static thread_local char g_err[512] = "";

  void set_error(const char* fmt, ...)
  {
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
  }

  int fail_cuda(cudaError_t e, const char* what, const char* file, int line)
  {
    set_error("CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    cudaGetLastError();   // reset the (non-sticky) last-error state so that the next call does not report it again
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return ONK_ERR_NO_DEVICE;
    if (e == cudaErrorMemoryAllocation) return ONK_ERR_NOMEM;
    return ONK_ERR_CUDA;
  }

  Runtime& rt() { static Runtime* r = new Runtime(); return *r; }

  int require_init()
  {
    if (!rt().initialised)
    {
      // lazy default: device 0, like CudaContext::default_cuda_ctx() (src/cuda/cuda_context.cpp:42-58)
      const int rc = onk_init(0);
      if (rc != ONK_OK) return rc;
    }
    return ONK_OK;
  }

  int lane_get(int lane, Lane** out)
  {
    if (int rc = require_init()) return rc;
    if (lane == ONK_LANE_DEFAULT) lane = 0;
    if (lane < 0 || lane >= ONK_MAX_LANES) { set_error("invalid execution lane %d", lane); return ONK_ERR_ARG; }
    Runtime& r = rt();
    std::lock_guard<std::mutex> lk(r.mutex);
    // the table is sized once (onk_init) and never reallocated: callers keep Lane* across further lane_get() calls
    if (r.lanes.size() < (size_t)ONK_MAX_LANES) r.lanes.resize(ONK_MAX_LANES);
    Lane& L = r.lanes[lane];
    if (!L.created)
    {
      ONK_CUDA(cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking));
      L.created = true; L.owned = true;
    }
    if (L.ws.partials == nullptr)
    {
      // all or nothing: a half-built workspace must never be seen by a later call
      LaneWorkspace w;
      cudaError_t e = cudaMalloc(&w.partials, sizeof(double) * MAX_PARTIALS);
      if (e == cudaSuccess) e = cudaMalloc(&w.ticket, 256);
      if (e == cudaSuccess) e = cudaMemset(w.ticket, 0, 256);
      if (e == cudaSuccess) e = cudaMalloc(&w.result, 256);
      if (e != cudaSuccess)
      {
        if (w.partials) cudaFree(w.partials);
        if (w.ticket) cudaFree(w.ticket);
        if (w.result) cudaFree(w.result);
        ONK_CUDA(e);
      }
      L.ws = w;
    }
    *out = &L;
    return ONK_OK;
  }
}

using namespace onk;

extern "C" {

int onk_abi_version(void) { return ONK_ABI_VERSION; }
const char* onk_last_error(void) { return g_err; }

int onk_device_count(int* count)
{
  ONK_REQUIRE(count != nullptr, "onk_device_count: null argument");
  *count = 0;
  ONK_CUDA(cudaGetDeviceCount(count));
  return ONK_OK;
}

int onk_init(int device)
{
  Runtime& r = rt();
  std::lock_guard<std::mutex> lk(r.mutex);
  if (r.initialised)
  {
    if (r.device == device) { ONK_CUDA(cudaSetDevice(device)); return ONK_OK; }
    set_error("onk_init: already initialised on device %d (one process drives one device)", r.device);
    return ONK_ERR_STATE;
  }
  int n = 0;
  ONK_CUDA(cudaGetDeviceCount(&n));
  if (n <= 0) { set_error("onk_init: no CUDA device (this library has no CPU fallback)"); return ONK_ERR_NO_DEVICE; }
  ONK_REQUIRE(device >= 0 && device < n, "onk_init: device %d out of range [0,%d)", device, n);
  ONK_CUDA(cudaSetDevice(device));
  ONK_CUDA(cudaGetDeviceProperties(&r.prop, device));
  r.device = device;
  r.sm_count = r.prop.multiProcessorCount;
  r.lanes.assign(ONK_MAX_LANES, Lane{});        // fixed size: Lane* handed out by lane_get() stay valid
  r.initialised = true;
  return ONK_OK;
}

int onk_finalize(void)
{
  Runtime& r = rt();
  std::lock_guard<std::mutex> lk(r.mutex);
  if (!r.initialised) return ONK_OK;
  cudaDeviceSynchronize();
  for (Lane& L : r.lanes)
  {
    if (L.ws.partials) cudaFree(L.ws.partials);
    if (L.ws.ticket) cudaFree(L.ws.ticket);
    if (L.ws.result) cudaFree(L.ws.result);
    if (L.edge) cudaEventDestroy(L.edge);
    if (L.created && L.owned) cudaStreamDestroy(L.stream);
    L = Lane{};
  }
  r.lanes.clear();
  for (int i = 0; i < 4; i++) if (r.stage_dev[i]) { cudaFree(r.stage_dev[i]); r.stage_dev[i] = nullptr; }
  r.stage_bytes = 0;
  r.initialised = false;
  r.device = -1;
  return ONK_OK;
}

int onk_device_info(int* device, int* sm_count, size_t* total_mem, char* name, size_t name_len)
{
  if (int rc = require_init()) return rc;
  Runtime& r = rt();
  if (device) *device = r.device;
  if (sm_count) *sm_count = r.sm_count;
  if (total_mem) *total_mem = r.prop.totalGlobalMem;
  if (name && name_len) { strncpy(name, r.prop.name, name_len - 1); name[name_len - 1] = '\0'; }
  return ONK_OK;
}

int onk_device_sync(void)
{
  if (int rc = require_init()) return rc;
  ONK_CUDA(cudaDeviceSynchronize());
  return ONK_OK;
}

/* ---------------------------------------------- lanes ---------------------------------------------- */

int onk_lane_stream(int lane, void** cuda_stream)
{
  ONK_REQUIRE(cuda_stream != nullptr, "onk_lane_stream: null argument");
  Lane* L; if (int rc = lane_get(lane, &L)) return rc;
  *cuda_stream = (void*)L->stream;
  return ONK_OK;
}

int onk_lane_bind_stream(int lane, void* cuda_stream)
{
  if (int rc = require_init()) return rc;
  if (lane == ONK_LANE_DEFAULT) lane = 0;
  ONK_REQUIRE(lane >= 0 && lane < ONK_MAX_LANES, "invalid execution lane %d", lane);
  {
    Runtime& r = rt();
    std::lock_guard<std::mutex> lk(r.mutex);
    if (r.lanes.size() < (size_t)ONK_MAX_LANES) r.lanes.resize(ONK_MAX_LANES);
    Lane& L = r.lanes[lane];
    if (L.created && L.owned) { cudaStreamSynchronize(L.stream); cudaStreamDestroy(L.stream); }
    L.stream = (cudaStream_t)cuda_stream;
    L.created = true; L.owned = false;
  }
  Lane* L; return lane_get(lane, &L);   // make sure the workspace exists
}

int onk_lane_sync(int lane)
{
  if (int rc = require_init()) return rc;
  if (lane == ONK_LANE_ALL)
  {
    std::vector<cudaStream_t> s;
    { Runtime& r = rt(); std::lock_guard<std::mutex> lk(r.mutex); for (Lane& L : r.lanes) if (L.created) s.push_back(L.stream); }
    for (cudaStream_t st : s) ONK_CUDA(cudaStreamSynchronize(st));
    return ONK_OK;
  }
  Lane* L; if (int rc = lane_get(lane, &L)) return rc;
  ONK_CUDA(cudaStreamSynchronize(L->stream));
  return ONK_OK;
}

int onk_lane_query(int lane, int* idle)
{
  ONK_REQUIRE(idle != nullptr, "onk_lane_query: null argument");
  if (int rc = require_init()) return rc;
  std::vector<cudaStream_t> s;
  if (lane == ONK_LANE_ALL)
  { Runtime& r = rt(); std::lock_guard<std::mutex> lk(r.mutex); for (Lane& L : r.lanes) if (L.created) s.push_back(L.stream); }
  else { Lane* L; if (int rc = lane_get(lane, &L)) return rc; s.push_back(L->stream); }
  *idle = 1;
  for (cudaStream_t st : s)
  {
    cudaError_t e = cudaStreamQuery(st);
    if (e == cudaErrorNotReady) { *idle = 0; continue; }
    ONK_CUDA(e);
  }
  return ONK_OK;
}

int onk_lane_wait_lane(int lane, int other_lane)
{
  Lane *A, *B;
  if (int rc = lane_get(lane, &A)) return rc;
  if (int rc = lane_get(other_lane, &B)) return rc;
  if (A->stream == B->stream) return ONK_OK;
  // one event per source lane, re-recorded at every edge: a wait already queued keeps referring to the record it saw
  if (B->edge == nullptr) ONK_CUDA(cudaEventCreateWithFlags(&B->edge, cudaEventDisableTiming));
  ONK_CUDA(cudaEventRecord(B->edge, B->stream));
  ONK_CUDA(cudaStreamWaitEvent(A->stream, B->edge, 0));
  return ONK_OK;
}

} // extern "C"

namespace onk
{
  // Host operations of a lane (user-declared host-only functors: OpenMP regions, arbitrary user code) do not run on the CUDA
  // runtime's callback thread: the stream's host node hands them to ONE long-lived worker thread of the library and blocks
  // until it is done -- the lane's ordering is unchanged, but the user's code (and its OpenMP team, created once) lives on
  // an ordinary thread instead of a driver thread whose identity may change from stream to stream.
  class HostWorker
  {
  public:
    static HostWorker& instance() { static HostWorker* w = new HostWorker(); return *w; }     // never destroyed: sleeps on a condition variable
    void run(void (*fn)(void*), void* user)
    {
      std::unique_lock<std::mutex> lk(m_mutex);
      m_slot_free.wait(lk, [&]{ return m_fn == nullptr; });          // several lanes may reach their host nodes at the same time
      m_fn = fn; m_user = user; const uint64_t ticket = ++m_posted;
      m_wake.notify_one();
      m_done.wait(lk, [&]{ return m_finished >= ticket; });
    }
  private:
    HostWorker() { std::thread([this]{ loop(); }).detach(); }
    void loop()
    {
      std::unique_lock<std::mutex> lk(m_mutex);
      for (;;)
      {
        m_wake.wait(lk, [&]{ return m_fn != nullptr; });
        void (*fn)(void*) = m_fn; void* user = m_user;
        lk.unlock();
        fn(user);
        lk.lock();
        m_fn = nullptr; m_finished = m_posted;
        m_done.notify_all(); m_slot_free.notify_one();
      }
    }
    std::mutex m_mutex;
    std::condition_variable m_wake, m_done, m_slot_free;
    void (*m_fn)(void*) = nullptr; void* m_user = nullptr;
    uint64_t m_posted = 0, m_finished = 0;
  };

  struct HostTask { void (*fn)(void*); void* user; };
  static void CUDART_CB host_task_trampoline(void* p)
  {
    HostTask* t = (HostTask*)p;
    HostWorker::instance().run(t->fn, t->user);
    delete t;
  }
}

extern "C" {

int onk_lane_host_func(int lane, void (*fn)(void*), void* user)
{
  ONK_REQUIRE(fn != nullptr, "onk_lane_host_func: null function");
  Lane* L; if (int rc = lane_get(lane, &L)) return rc;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  ONK_CUDA(cudaStreamIsCapturing(L->stream, &cap));
  if (cap != cudaStreamCaptureStatusNone)
  {
    // a captured host node is replayed any number of times: it keeps the caller's (fn, user) pair, which must outlive the graph
    ONK_CUDA(cudaLaunchHostFunc(L->stream, fn, user));
    return ONK_OK;
  }
  HostTask* t = new HostTask{ fn, user };
  cudaError_t e = cudaLaunchHostFunc(L->stream, host_task_trampoline, t);
  if (e != cudaSuccess) { delete t; ONK_CUDA(e); }
  return ONK_OK;
}

/* ---------------------------------------------- memory ---------------------------------------------- */

int onk_alloc(size_t bytes, size_t align, int kind, void** ptr)
{
  ONK_REQUIRE(ptr != nullptr, "onk_alloc: null argument");
  *ptr = nullptr;
  if (int rc = require_init()) return rc;
  if (bytes == 0) return ONK_OK;
  ONK_REQUIRE(align == 0 || (align & (align - 1)) == 0, "onk_alloc: alignment %zu is not a power of two", align);
  // cudaMalloc / cudaMallocManaged / cudaHostAlloc all return >= 256-byte aligned memory (MINIMUM_CUDA_ALIGNMENT,
  // include/onika/memory/allocator.h:61-65); larger alignments are checked, as the reference does (allocator.cpp:158-163)
  void* p = nullptr;
  switch (kind)
  {
    case ONK_MEM_DEVICE:  ONK_CUDA(cudaMalloc(&p, bytes)); break;
    case ONK_MEM_MANAGED: ONK_CUDA(cudaMallocManaged(&p, bytes)); break;
    case ONK_MEM_PINNED:  ONK_CUDA(cudaHostAlloc(&p, bytes, cudaHostAllocMapped | cudaHostAllocPortable)); break;
    default: set_error("onk_alloc: unknown memory kind %d", kind); return ONK_ERR_ARG;
  }
  if (align > 1 && ((uintptr_t)p % align) != 0)
  {
    if (kind == ONK_MEM_PINNED) cudaFreeHost(p); else cudaFree(p);
    set_error("onk_alloc: allocator returned a pointer that is not aligned on %zu bytes", align);
    return ONK_ERR_NOMEM;
  }
  { Runtime& r = rt(); std::lock_guard<std::mutex> lk(r.mutex); r.allocs[p] = { bytes, kind }; }
  *ptr = p;
  return ONK_OK;
}

int onk_free(void* ptr, size_t bytes)
{
  if (ptr == nullptr) return ONK_OK;
  if (int rc = require_init()) return rc;
  int kind;
  {
    Runtime& r = rt(); std::lock_guard<std::mutex> lk(r.mutex);
    auto it = r.allocs.find(ptr);
    if (it == r.allocs.end()) { set_error("onk_free: %p was not allocated by onk_alloc", ptr); return ONK_ERR_ARG; }
    if (bytes != 0 && it->second.first != bytes)
    { // the reference aborts on a size/trailer mismatch (src/cuda/allocator.cpp:73-96)
      set_error("onk_free: size mismatch for %p (%zu given, %zu allocated)", ptr, bytes, it->second.first);
      return ONK_ERR_ARG;
    }
    kind = it->second.second;
    r.allocs.erase(it);
  }
  if (kind == ONK_MEM_PINNED) ONK_CUDA(cudaFreeHost(ptr)); else ONK_CUDA(cudaFree(ptr));
  return ONK_OK;
}

int onk_is_gpu_addressable(const void* ptr, int* yes)
{
  ONK_REQUIRE(yes != nullptr, "onk_is_gpu_addressable: null argument");
  *yes = 0;
  if (ptr == nullptr) { *yes = 1; return ONK_OK; }   // reference: nullptr is addressable (allocator.cpp:100)
  if (int rc = require_init()) return rc;
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, ptr);
  if (e != cudaSuccess) { cudaGetLastError(); return ONK_OK; }
  *yes = (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged || a.type == cudaMemoryTypeHost) ? 1 : 0;
  return ONK_OK;
}

int onk_memcpy(void* dst, const void* src, size_t bytes, int lane)
{
  if (bytes == 0) return ONK_OK;
  ONK_REQUIRE(dst != nullptr && src != nullptr, "onk_memcpy: null pointer");
  Lane* L; if (int rc = lane_get(lane, &L)) return rc;
  ONK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, L->stream));
  return ONK_OK;
}

int onk_memset_bytes(void* dst_dev, int byte, size_t bytes, int lane)
{
  if (bytes == 0) return ONK_OK;
  ONK_REQUIRE(dst_dev != nullptr, "onk_memset_bytes: null pointer");
  Lane* L; if (int rc = lane_get(lane, &L)) return rc;
  ONK_CUDA(cudaMemsetAsync(dst_dev, byte, bytes, L->stream));
  return ONK_OK;
}

int onk_prefetch(const void* ptr, size_t bytes, int to_device, int lane)
{
  if (bytes == 0 || ptr == nullptr) return ONK_OK;
  Lane* L; if (int rc = lane_get(lane, &L)) return rc;
  ONK_CUDA(cudaMemPrefetchAsync(ptr, bytes, to_device ? rt().device : cudaCpuDeviceId, L->stream));
  return ONK_OK;
}

/* ------------------------------------- parallel-operation bracket ------------------------------------- */

struct onk_op
{
  void*        scratch = nullptr;   // GPUKernelExecutionScratch: 8 x u64 counters, then 960 B of return data
  cudaEvent_t  start = nullptr, stop = nullptr;
  cudaStream_t stream = nullptr;
  void       (*epilog)(void*) = nullptr;
  void*        user = nullptr;
  bool         began = false, ended = false;
  bool         captured = false;    // bracket recorded into a CUDA graph: no timing events
};

static void CUDART_CB onk_op_epilog_trampoline(void* p)
{
  onk_op* op = (onk_op*)p;
  if (op->epilog) op->epilog(op->user);
}

int onk_op_create(onk_op** op)
{
  ONK_REQUIRE(op != nullptr, "onk_op_create: null argument");
  if (int rc = require_init()) return rc;
  onk_op* o = new onk_op();
  cudaError_t e = cudaEventCreate(&o->start);
  if (e == cudaSuccess) e = cudaEventCreate(&o->stop);
  if (e != cudaSuccess) { delete o; ONK_CUDA(e); }
  *op = o;
  return ONK_OK;
}

int onk_op_destroy(onk_op* op)
{
  if (!op) return ONK_OK;
  if (op->scratch) cudaFree(op->scratch);
  if (op->start) cudaEventDestroy(op->start);
  if (op->stop) cudaEventDestroy(op->stop);
  delete op;
  return ONK_OK;
}

int onk_op_device_scratch(onk_op* op, void** scratch_dev)
{
  ONK_REQUIRE(op != nullptr && scratch_dev != nullptr, "onk_op_device_scratch: null argument");
  if (!op->scratch)
  {
    ONK_CUDA(cudaMalloc(&op->scratch, ONK_OP_SCRATCH_BYTES));
    ONK_CUDA(cudaMemset(op->scratch, 0, ONK_OP_SCRATCH_BYTES));
  }
  *scratch_dev = op->scratch;
  return ONK_OK;
}

int onk_op_begin(onk_op* op, int lane, const void* return_init_host, size_t return_bytes, int reset_counters)
{
  ONK_REQUIRE(op != nullptr, "onk_op_begin: null op");
  ONK_REQUIRE(return_bytes <= ONK_OP_MAX_RETURN_BYTES, "return data size too large (%zu > %d)", return_bytes, ONK_OP_MAX_RETURN_BYTES);
  Lane* L; if (int rc = lane_get(lane, &L)) return rc;
  void* scratch; if (int rc = onk_op_device_scratch(op, &scratch)) return rc;
  op->stream = L->stream;
  op->began = true; op->ended = false;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  ONK_CUDA(cudaStreamIsCapturing(op->stream, &cap));
  // timing events mean nothing inside a graph, and a caller may opt out of them (ONK_OP_NO_TIMING): elapsed = 0 then
  op->captured = (cap != cudaStreamCaptureStatusNone) || (reset_counters & ONK_OP_NO_TIMING) != 0;
  if (!op->captured) ONK_CUDA(cudaEventRecord(op->start, op->stream));
  if (return_init_host != nullptr && return_bytes > 0)
    ONK_CUDA(cudaMemcpyAsync((char*)scratch + 64, return_init_host, return_bytes, cudaMemcpyDefault, op->stream));
  if (reset_counters & ONK_OP_RESET_COUNTERS)
    ONK_CUDA(cudaMemsetAsync(scratch, 0, 64, op->stream));
  return ONK_OK;
}

int onk_op_end(onk_op* op, void* return_out_host, size_t return_bytes, void (*epilog)(void*), void* user)
{
  ONK_REQUIRE(op != nullptr && op->began, "onk_op_end: op was not begun");
  ONK_REQUIRE(return_bytes <= ONK_OP_MAX_RETURN_BYTES, "return data size too large (%zu > %d)", return_bytes, ONK_OP_MAX_RETURN_BYTES);
  if (epilog != nullptr)
  {
    op->epilog = epilog; op->user = user;
    ONK_CUDA(cudaLaunchHostFunc(op->stream, onk_op_epilog_trampoline, op));
  }
  if (return_out_host != nullptr && return_bytes > 0)
    ONK_CUDA(cudaMemcpyAsync(return_out_host, (char*)op->scratch + 64, return_bytes, cudaMemcpyDefault, op->stream));
  if (!op->captured) ONK_CUDA(cudaEventRecord(op->stop, op->stream));
  op->ended = true;
  return ONK_OK;
}

int onk_op_elapsed_ms(onk_op* op, float* ms)
{
  ONK_REQUIRE(op != nullptr && ms != nullptr && op->ended, "onk_op_elapsed_ms: op has not ended");
  if (op->captured) { *ms = 0.0f; return ONK_OK; }
  ONK_CUDA(cudaEventElapsedTime(ms, op->start, op->stop));
  return ONK_OK;
}

/* ---------------------------------------------- generic launch ---------------------------------------------- */

typedef CUresult (*cuLaunchKernelEx_t)(const CUlaunchConfig*, CUfunction, void**, void**);

int onk_launch_ex(void* cu_function, const unsigned grid[3], const unsigned block[3], size_t shared_bytes, int lane, void** args, int flags)
{
  ONK_REQUIRE(cu_function != nullptr && grid != nullptr && block != nullptr, "onk_launch: null argument");
  ONK_REQUIRE((flags & ~ONK_LAUNCH_SUB_BLOCKS) == 0, "onk_launch_ex: unknown flags %d", flags);
  Lane* L; if (int rc = lane_get(lane, &L)) return rc;
  if ((uint64_t)grid[0] * grid[1] * grid[2] == 0) return ONK_OK;   // empty execution space: nothing to run
  static std::atomic<cuLaunchKernelEx_t> s_launch{nullptr};
  cuLaunchKernelEx_t launch = s_launch.load(std::memory_order_acquire);
  if (!launch)
  {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    ONK_CUDA(cudaGetDriverEntryPoint("cuLaunchKernelEx", &fn, cudaEnableDefault, &q));
    if (!fn) { set_error("onk_launch: cuLaunchKernelEx entry point not found"); return ONK_ERR_CUDA; }
    launch = (cuLaunchKernelEx_t)fn;
    s_launch.store(launch, std::memory_order_release);
  }
  CUlaunchConfig cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDimX = grid[0]; cfg.gridDimY = grid[1]; cfg.gridDimZ = grid[2];
  cfg.blockDimX = block[0]; cfg.blockDimY = block[1]; cfg.blockDimZ = block[2];
  cfg.sharedMemBytes = (unsigned)shared_bytes;
  cfg.hStream = (CUstream)L->stream;
  CUlaunchAttribute attr[1];
  if (flags & ONK_LAUNCH_SUB_BLOCKS)
  {
    // explicit cluster of one CTA: %is_explicit_cluster is how the ONIKA_CU_* macros recognise a CTA of sub-blocks
    attr[0].id = CU_LAUNCH_ATTRIBUTE_CLUSTER_DIMENSION;
    attr[0].value.clusterDim.x = 1; attr[0].value.clusterDim.y = 1; attr[0].value.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
  }
  CUresult r = launch(&cfg, (CUfunction)cu_function, args, nullptr);
  if (r != CUDA_SUCCESS) { set_error("onk_launch: cuLaunchKernelEx failed with CUresult %d", (int)r); return ONK_ERR_CUDA; }
  count_launch();
  return ONK_OK;
}

int onk_launch(void* cu_function, const unsigned grid[3], const unsigned block[3], size_t shared_bytes, int lane, void** args)
{
  return onk_launch_ex(cu_function, grid, block, shared_bytes, lane, args, 0);
}

int onk_launch_count(uint64_t* n)
{
  ONK_REQUIRE(n != nullptr, "onk_launch_count: null argument");
  *n = rt().launches.load();
  return ONK_OK;
}

/* ---------------------------------------------- lane-sequence graphs ---------------------------------------------- */

struct onk_graph { cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr; uint64_t kernels = 0; };

static struct { bool active = false; cudaStream_t origin = nullptr; std::vector<cudaStream_t> others; uint64_t launches0 = 0; } g_capture;   // guarded by rt().mutex

// leaves every lane of a failed capture out of capture mode and forgets it
static void capture_abort(cudaStream_t origin)
{
  cudaGraph_t g = nullptr;
  cudaStreamEndCapture(origin, &g);             // invalidates the capture on all forked streams as well
  if (g) cudaGraphDestroy(g);
  cudaGetLastError();
  g_capture.active = false; g_capture.origin = nullptr; g_capture.others.clear();
}

int onk_graph_begin(int origin_lane, const int* other_lanes, int n_other)
{
  ONK_REQUIRE(n_other == 0 || other_lanes != nullptr, "onk_graph_begin: null lane list");
  Lane* O; if (int rc = lane_get(origin_lane, &O)) return rc;
  ONK_REQUIRE(O->stream != nullptr && O->stream != cudaStreamLegacy,
              "onk_graph_begin: lane %d runs on the legacy default stream, which cannot be captured", origin_lane);
  std::vector<cudaStream_t> others;
  for (int i = 0; i < n_other; i++) { Lane* L; if (int rc = lane_get(other_lanes[i], &L)) return rc; if (L->stream != O->stream) others.push_back(L->stream); }
  std::lock_guard<std::mutex> lk(rt().mutex);
  ONK_REQUIRE(!g_capture.active, "onk_graph_begin: a capture is already active");
  ONK_CUDA(cudaStreamBeginCapture(O->stream, cudaStreamCaptureModeThreadLocal));
  // fork: every other lane joins the capture by waiting on an event recorded in the origin lane
  cudaEvent_t fork = nullptr;
  cudaError_t e = cudaEventCreateWithFlags(&fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventRecord(fork, O->stream);
  for (size_t i = 0; e == cudaSuccess && i < others.size(); i++) e = cudaStreamWaitEvent(others[i], fork, 0);
  if (fork) cudaEventDestroy(fork);
  if (e != cudaSuccess) { capture_abort(O->stream); ONK_CUDA(e); }
  g_capture.active = true; g_capture.origin = O->stream; g_capture.others = others; g_capture.launches0 = rt().launches.load();
  return ONK_OK;
}

int onk_graph_end(onk_graph** graph)
{
  std::lock_guard<std::mutex> lk(rt().mutex);
  ONK_REQUIRE(graph != nullptr && g_capture.active, "onk_graph_end: no active capture");
  // join: the origin lane waits for every other lane
  cudaError_t e = cudaSuccess;
  for (size_t i = 0; e == cudaSuccess && i < g_capture.others.size(); i++)
  {
    cudaEvent_t j = nullptr;
    e = cudaEventCreateWithFlags(&j, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(j, g_capture.others[i]);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(g_capture.origin, j, 0);
    if (j) cudaEventDestroy(j);
  }
  if (e != cudaSuccess) { capture_abort(g_capture.origin); ONK_CUDA(e); }
  onk_graph* g = new onk_graph();
  e = cudaStreamEndCapture(g_capture.origin, &g->graph);
  g_capture.active = false; g_capture.origin = nullptr; g_capture.others.clear();
  if (e == cudaSuccess) e = cudaGraphInstantiate(&g->exec, g->graph, 0);
  if (e != cudaSuccess) { if (g->graph) cudaGraphDestroy(g->graph); delete g; ONK_CUDA(e); }
  g->kernels = rt().launches.load() - g_capture.launches0;
  // captured launches did not execute: they are counted when the graph is launched
  rt().launches.fetch_sub(g->kernels);
  *graph = g;
  return ONK_OK;
}

int onk_graph_launch(onk_graph* graph, int lane)
{
  ONK_REQUIRE(graph != nullptr && graph->exec != nullptr, "onk_graph_launch: null graph");
  Lane* L; if (int rc = lane_get(lane, &L)) return rc;
  ONK_CUDA(cudaGraphLaunch(graph->exec, L->stream));
  count_launch(graph->kernels);
  return ONK_OK;
}

int onk_graph_destroy(onk_graph* graph)
{
  if (!graph) return ONK_OK;
  if (graph->exec) cudaGraphExecDestroy(graph->exec);
  if (graph->graph) cudaGraphDestroy(graph->graph);
  delete graph;
  return ONK_OK;
}

} // extern "C"
