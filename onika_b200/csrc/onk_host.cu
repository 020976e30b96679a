// onk_host.cu -- host-buffer entry points: what a plugin that holds its data in HOST memory calls.
//
// Reference behaviour: the reference keeps everything in CUDA managed memory (src/cuda/allocator.cpp:157) and lets
// UVM page-fault the data across on first touch ("distributed communication backend").  Here the
// transfers are explicit: the input is cut in chunks, chunk k+1 crosses PCIe (copy engine, lane ONK_LANE_H2D) while
// the kernel consumes chunk k (lane ONK_LANE_COMPUTE) and, for in-place updates, chunk k-1 travels back (lane
// ONK_LANE_D2H); device staging is double-buffered and recycled through events.  Results are identical to the
// device-resident entry points because the very same kernels run on the chunks (reductions: per-chunk partials are
// combined in chunk order by onk_combine_f64).
//
// PAGEABLE input (ordinary malloc'ed memory, what a plugin usually holds): the copy engine cannot read it directly and the
// driver's own bounce path runs at ~11 GB/s on this class of host.  The library stages it itself: a small pool of host
// threads copies chunk k+1 into a ring of pinned buffers while chunk k crosses PCIe from the previous ring slot
// (measured on the B200 box: see DESIGN.md "e2e").  Pinned / registered / managed input skips the staging.
#include "onk_internal.cuh"

#include <condition_variable>
#include <thread>
#include <sched.h>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

namespace onk
{
  // copy into a pinned staging buffer that the CPU will not read again: non-temporal stores skip the read-for-ownership
  // of the destination lines, one pass less over host DRAM (the staging path is bound by host memory bandwidth)
  static void copy_streaming(char* dst, const char* src, size_t bytes)
  {
#if defined(__x86_64__)
    if (((uintptr_t)dst & 15) == 0)
    {
      const size_t nvec = bytes / 64;
      const __m128i* s = (const __m128i*)src;
      __m128i* d = (__m128i*)dst;
      for (size_t i = 0; i < nvec; i++)
      {
        const __m128i a = _mm_loadu_si128(s + 4 * i), b = _mm_loadu_si128(s + 4 * i + 1), c = _mm_loadu_si128(s + 4 * i + 2), e = _mm_loadu_si128(s + 4 * i + 3);
        _mm_stream_si128(d + 4 * i, a); _mm_stream_si128(d + 4 * i + 1, b); _mm_stream_si128(d + 4 * i + 2, c); _mm_stream_si128(d + 4 * i + 3, e);
      }
      _mm_sfence();
      if (bytes & 63) memcpy(dst + nvec * 64, src + nvec * 64, bytes & 63);
      return;
    }
#endif
    memcpy(dst, src, bytes);
  }

  // ---- host copy pool: copy of one large block split over a few persistent threads ----
  class CopyPool
  {
  public:
    static CopyPool& instance() { static CopyPool* p = new CopyPool(); return *p; }     // never destroyed: workers sleep on a condition variable
    int threads() const { return m_nthreads; }
    void copy(void* dst, const void* src, size_t bytes)
    {
      if (bytes < (1u << 20) || m_nthreads <= 1) { copy_streaming((char*)dst, (const char*)src, bytes); return; }
      {
        std::lock_guard<std::mutex> lk(m_mutex);
        m_dst = (char*)dst; m_src = (const char*)src; m_bytes = bytes; m_pending = m_nthreads - 1; ++m_generation;
      }
      m_wake.notify_all();
      slice(0, (char*)dst, (const char*)src, bytes);                                       // the caller is thread 0
      std::unique_lock<std::mutex> lk(m_mutex);
      m_done.wait(lk, [&]{ return m_pending == 0; });
    }
  private:
    CopyPool()
    {
      int n = 0;
      cpu_set_t set;
      if (sched_getaffinity(0, sizeof(set), &set) == 0) n = CPU_COUNT(&set);
      if (n <= 0) n = (int)std::thread::hardware_concurrency();
      if (const char* e = getenv("ONK_HOST_COPY_THREADS")) n = atoi(e);
      m_nthreads = n < 1 ? 1 : (n > 12 ? 12 : n);
      for (int t = 1; t < m_nthreads; t++) std::thread([this, t]{ worker(t); }).detach();
    }
    void slice(int t, char* dst, const char* src, size_t bytes) const
    {
      const size_t per = ((bytes / m_nthreads) + 4095) & ~(size_t)4095;
      const size_t b = (size_t)t * per;
      if (b >= bytes) return;
      const size_t e = (t == m_nthreads - 1 || b + per > bytes) ? bytes : b + per;
      copy_streaming(dst + b, src + b, e - b);
    }
    void worker(int t)
    {
      uint64_t seen = 0;
      for (;;)
      {
        char* d; const char* s; size_t n;
        {
          std::unique_lock<std::mutex> lk(m_mutex);
          m_wake.wait(lk, [&]{ return m_generation != seen; });
          seen = m_generation; d = m_dst; s = m_src; n = m_bytes;
        }
        slice(t, d, s, n);
        {
          std::lock_guard<std::mutex> lk(m_mutex);
          if (--m_pending == 0) m_done.notify_one();
        }
      }
    }
    std::mutex m_mutex;
    std::condition_variable m_wake, m_done;
    char* m_dst = nullptr; const char* m_src = nullptr; size_t m_bytes = 0;
    uint64_t m_generation = 0;
    int m_pending = 0, m_nthreads = 1;
  };

  // pinned rings for pageable buffers: slots [0,3) first input, [3,6) second input (axpy's x), [6,8) results on their way back
  constexpr int    PIN_SLOTS = 3, PIN_X0 = 3, PIN_OUT0 = 6, PIN_OUT_SLOTS = 2, PIN_TOTAL = 8;
  static void*     g_pin[PIN_TOTAL] = {};

  // can the copy engine read this host pointer directly ?  (pinned, registered or managed memory)
  static bool dma_readable(const void* p)
  {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type != cudaMemoryTypeUnregistered;
  }

  constexpr int LANE_H2D = ONK_MAX_LANES - 1, LANE_COMPUTE = ONK_MAX_LANES - 2, LANE_D2H = ONK_MAX_LANES - 3;
  constexpr size_t STAGE_BYTES = 64u << 20;       // per staging buffer
  constexpr int MAX_CHUNK_PARTIALS = MAX_PARTIALS;

  // the staging buffers and the three internal lanes are shared: one host-buffer call at a time
  static std::mutex g_host_path_mutex;

  static int ensure_staging(int nbuf)
  {
    Runtime& r = rt();
    for (int i = 0; i < nbuf; i++)
    {
      if (!r.stage_dev[i]) ONK_CUDA(cudaMalloc(&r.stage_dev[i], STAGE_BYTES));
    }
    r.stage_bytes = STAGE_BYTES;
    return ONK_OK;
  }
  static int ensure_pinned_ring(int first = 0, int count = PIN_SLOTS)
  {
    for (int i = first; i < first + count; i++)
      if (!g_pin[i]) ONK_CUDA(cudaHostAlloc(&g_pin[i], STAGE_BYTES, cudaHostAllocDefault));
    return ONK_OK;
  }

  struct Ev
  {
    cudaEvent_t e = nullptr;
    int init() { ONK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); return ONK_OK; }
    ~Ev() { if (e) cudaEventDestroy(e); }
  };
}

using namespace onk;

extern "C" {

int onk_host_reduce_f64(int op, const double* in_host, uint64_t n, double* out_host)
{
  ONK_REQUIRE(out_host != nullptr && (n == 0 || in_host != nullptr), "onk_host_reduce_f64: null pointer");
  ONK_REQUIRE(op == ONK_OP_MIN || op == ONK_OP_MAX || op == ONK_OP_SUM, "onk_host_reduce_f64: unknown op %d", op);
  std::lock_guard<std::mutex> guard(g_host_path_mutex);
  Lane *H, *K;
  if (int rc = lane_get(LANE_H2D, &H)) return rc;
  if (int rc = lane_get(LANE_COMPUTE, &K)) return rc;
  if (int rc = ensure_staging(2)) return rc;
  Runtime& r = rt();
  const uint64_t chunk = STAGE_BYTES / sizeof(double);
  const uint64_t nchunks = (n + chunk - 1) / chunk;
  ONK_REQUIRE(nchunks <= (uint64_t)MAX_CHUNK_PARTIALS, "onk_host_reduce_f64: input too large (%llu chunks)", (unsigned long long)nchunks);
  // per-chunk partials live in the H2D lane's partial array (that lane never runs a reduction itself)
  double* chunk_part = H->ws.partials;
  Ev ready[2], freed[2], sent[PIN_SLOTS];
  for (int i = 0; i < 2; i++) { if (int rc = ready[i].init()) return rc; if (int rc = freed[i].init()) return rc; }
  const bool staged = n > 0 && !dma_readable(in_host);        // pageable input: through the pinned ring
  if (staged)
  {
    if (int rc = ensure_pinned_ring()) return rc;
    for (int i = 0; i < PIN_SLOTS; i++) if (int rc = sent[i].init()) return rc;
  }
  for (uint64_t k = 0; k < nchunks; k++)
  {
    const int b = (int)(k & 1);
    const uint64_t cnt = (k + 1 == nchunks) ? n - k * chunk : chunk;
    const double* src = in_host + k * chunk;
    if (staged)
    {
      const int s = (int)(k % PIN_SLOTS);
      if (k >= (uint64_t)PIN_SLOTS) ONK_CUDA(cudaEventSynchronize(sent[s].e));    // slot s has crossed PCIe
      CopyPool::instance().copy(g_pin[s], src, cnt * sizeof(double));            // overlaps the transfers of the previous slots
      src = (const double*)g_pin[s];
    }
    if (k >= 2) ONK_CUDA(cudaStreamWaitEvent(H->stream, freed[b].e, 0));
    ONK_CUDA(cudaMemcpyAsync(r.stage_dev[b], src, cnt * sizeof(double), cudaMemcpyHostToDevice, H->stream));
    if (staged) ONK_CUDA(cudaEventRecord(sent[(int)(k % PIN_SLOTS)].e, H->stream));
    ONK_CUDA(cudaEventRecord(ready[b].e, H->stream));
    ONK_CUDA(cudaStreamWaitEvent(K->stream, ready[b].e, 0));
    if (int rc = onk_reduce_f64(op, (const double*)r.stage_dev[b], cnt, chunk_part + k, LANE_COMPUTE)) return rc;
    ONK_CUDA(cudaEventRecord(freed[b].e, K->stream));
  }
  if (int rc = onk_combine_f64(op, chunk_part, (uint32_t)nchunks, K->ws.result, LANE_COMPUTE)) return rc;
  ONK_CUDA(cudaMemcpyAsync(out_host, K->ws.result, sizeof(double), cudaMemcpyDeviceToHost, K->stream));
  ONK_CUDA(cudaStreamSynchronize(K->stream));
  return ONK_OK;
}

// shared driver of the in-place host updates: y (and optionally x) travel in, y travels back
static int host_update(double* y_host, const double* x_host, uint64_t n, double a, bool is_axpy)
{
  std::lock_guard<std::mutex> guard(g_host_path_mutex);
  Lane *H, *K, *D;
  if (int rc = lane_get(LANE_H2D, &H)) return rc;
  if (int rc = lane_get(LANE_COMPUTE, &K)) return rc;
  if (int rc = lane_get(LANE_D2H, &D)) return rc;
  if (int rc = ensure_staging(is_axpy ? 4 : 2)) return rc;
  Runtime& r = rt();
  const uint64_t chunk = STAGE_BYTES / sizeof(double);
  const uint64_t nchunks = (n + chunk - 1) / chunk;
  Ev ready[2], done[2], drained[2], sent_y[PIN_SLOTS], sent_x[PIN_SLOTS], landed[PIN_OUT_SLOTS];
  for (int i = 0; i < 2; i++) { if (int rc = ready[i].init()) return rc; if (int rc = done[i].init()) return rc; if (int rc = drained[i].init()) return rc; }
  // pageable buffers travel through pinned rings filled / drained by the host copy pool (see onk_host_reduce_f64)
  const bool stage_y = !dma_readable(y_host), stage_x = is_axpy && !dma_readable(x_host);
  if (stage_y)
  {
    if (int rc = ensure_pinned_ring(0, PIN_SLOTS)) return rc;
    if (int rc = ensure_pinned_ring(PIN_OUT0, PIN_OUT_SLOTS)) return rc;
    for (int i = 0; i < PIN_SLOTS; i++) if (int rc = sent_y[i].init()) return rc;
    for (int i = 0; i < PIN_OUT_SLOTS; i++) if (int rc = landed[i].init()) return rc;
  }
  if (stage_x)
  {
    if (int rc = ensure_pinned_ring(PIN_X0, PIN_SLOTS)) return rc;
    for (int i = 0; i < PIN_SLOTS; i++) if (int rc = sent_x[i].init()) return rc;
  }
  auto unstage = [&](uint64_t j) -> int                     // results of chunk j: pinned ring -> the caller's pageable memory
  {
    const int so = (int)(j % PIN_OUT_SLOTS);
    const uint64_t cj = (j + 1 == nchunks) ? n - j * chunk : chunk;
    ONK_CUDA(cudaEventSynchronize(landed[so].e));
    CopyPool::instance().copy(y_host + j * chunk, g_pin[PIN_OUT0 + so], cj * sizeof(double));
    return ONK_OK;
  };
  for (uint64_t k = 0; k < nchunks; k++)
  {
    const int b = (int)(k & 1);
    const uint64_t cnt = (k + 1 == nchunks) ? n - k * chunk : chunk;
    double* yd = (double*)r.stage_dev[b];
    double* xd = is_axpy ? (double*)r.stage_dev[2 + b] : nullptr;
    const double* ysrc = y_host + k * chunk;
    const double* xsrc = is_axpy ? x_host + k * chunk : nullptr;
    const int si = (int)(k % PIN_SLOTS);
    if (stage_y)
    {
      if (k >= (uint64_t)PIN_SLOTS) ONK_CUDA(cudaEventSynchronize(sent_y[si].e));
      CopyPool::instance().copy(g_pin[si], ysrc, cnt * sizeof(double));
      ysrc = (const double*)g_pin[si];
    }
    if (stage_x)
    {
      if (k >= (uint64_t)PIN_SLOTS) ONK_CUDA(cudaEventSynchronize(sent_x[si].e));
      CopyPool::instance().copy(g_pin[PIN_X0 + si], xsrc, cnt * sizeof(double));
      xsrc = (const double*)g_pin[PIN_X0 + si];
    }
    if (k >= 2) ONK_CUDA(cudaStreamWaitEvent(H->stream, drained[b].e, 0));      // buffer b has been copied back
    ONK_CUDA(cudaMemcpyAsync(yd, ysrc, cnt * sizeof(double), cudaMemcpyHostToDevice, H->stream));
    if (stage_y) ONK_CUDA(cudaEventRecord(sent_y[si].e, H->stream));
    if (is_axpy) ONK_CUDA(cudaMemcpyAsync(xd, xsrc, cnt * sizeof(double), cudaMemcpyHostToDevice, H->stream));
    if (stage_x) ONK_CUDA(cudaEventRecord(sent_x[si].e, H->stream));
    ONK_CUDA(cudaEventRecord(ready[b].e, H->stream));
    ONK_CUDA(cudaStreamWaitEvent(K->stream, ready[b].e, 0));
    if (is_axpy) { if (int rc = onk_axpy_f64(yd, xd, a, cnt, LANE_COMPUTE)) return rc; }
    else         { if (int rc = onk_value_add_f64(yd, 1, cnt, a, LANE_COMPUTE)) return rc; }
    ONK_CUDA(cudaEventRecord(done[b].e, K->stream));
    ONK_CUDA(cudaStreamWaitEvent(D->stream, done[b].e, 0));
    if (stage_y)
    {
      const int so = (int)(k % PIN_OUT_SLOTS);            // free: chunk k-2 was unstaged in the previous iteration
      ONK_CUDA(cudaMemcpyAsync(g_pin[PIN_OUT0 + so], yd, cnt * sizeof(double), cudaMemcpyDeviceToHost, D->stream));
      ONK_CUDA(cudaEventRecord(landed[so].e, D->stream));
    }
    else ONK_CUDA(cudaMemcpyAsync(y_host + k * chunk, yd, cnt * sizeof(double), cudaMemcpyDeviceToHost, D->stream));
    ONK_CUDA(cudaEventRecord(drained[b].e, D->stream));
    if (stage_y && k >= 1) { if (int rc = unstage(k - 1)) return rc; }          // overlaps the transfers of chunk k
  }
  if (stage_y && nchunks >= 1) { if (int rc = unstage(nchunks - 1)) return rc; }
  ONK_CUDA(cudaStreamSynchronize(D->stream));
  ONK_CUDA(cudaStreamSynchronize(K->stream));
  return ONK_OK;
}

int onk_host_axpy_f64(double* y_host, const double* x_host, double a, uint64_t n)
{
  if (n == 0) return ONK_OK;
  ONK_REQUIRE(y_host != nullptr && x_host != nullptr, "onk_host_axpy_f64: null pointer");
  return host_update(y_host, x_host, n, a, true);
}

int onk_host_value_add_f64(double* array_host, uint64_t rows, uint64_t cols, double value)
{
  const uint64_t n = rows * cols;
  if (n == 0) return ONK_OK;
  ONK_REQUIRE(array_host != nullptr, "onk_host_value_add_f64: null pointer");
  return host_update(array_host, nullptr, n, value, false);
}

} // extern "C"
