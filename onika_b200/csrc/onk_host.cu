// onk_host.cu -- host-buffer entry points: what a plugin that holds its data in HOST memory calls.
//
// Reference behaviour: the reference keeps everything in CUDA managed memory (src/cuda/allocator.cpp:157) and lets
// UVM page-fault the data across on first touch ("distributed communication backend").  Here the
// transfers are explicit: the input is cut in chunks, chunk k+1 crosses PCIe (copy engine, lane ONK_LANE_H2D) while
// the kernel consumes chunk k (lane ONK_LANE_COMPUTE) and, for in-place updates, chunk k-1 travels back (lane
// ONK_LANE_D2H); device staging is double-buffered and recycled through events.  Results are identical to the
// device-resident entry points because the very same kernels run on the chunks (reductions: per-chunk partials are
// combined in chunk order by onk_combine_f64).
#include "onk_internal.cuh"

namespace onk
{
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
  Ev ready[2], freed[2];
  for (int i = 0; i < 2; i++) { if (int rc = ready[i].init()) return rc; if (int rc = freed[i].init()) return rc; }
  for (uint64_t k = 0; k < nchunks; k++)
  {
    const int b = (int)(k & 1);
    const uint64_t cnt = (k + 1 == nchunks) ? n - k * chunk : chunk;
    if (k >= 2) ONK_CUDA(cudaStreamWaitEvent(H->stream, freed[b].e, 0));
    ONK_CUDA(cudaMemcpyAsync(r.stage_dev[b], in_host + k * chunk, cnt * sizeof(double), cudaMemcpyHostToDevice, H->stream));
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
  Ev ready[2], done[2], drained[2];
  for (int i = 0; i < 2; i++) { if (int rc = ready[i].init()) return rc; if (int rc = done[i].init()) return rc; if (int rc = drained[i].init()) return rc; }
  for (uint64_t k = 0; k < nchunks; k++)
  {
    const int b = (int)(k & 1);
    const uint64_t cnt = (k + 1 == nchunks) ? n - k * chunk : chunk;
    double* yd = (double*)r.stage_dev[b];
    double* xd = is_axpy ? (double*)r.stage_dev[2 + b] : nullptr;
    if (k >= 2) ONK_CUDA(cudaStreamWaitEvent(H->stream, drained[b].e, 0));      // buffer b has been copied back
    ONK_CUDA(cudaMemcpyAsync(yd, y_host + k * chunk, cnt * sizeof(double), cudaMemcpyHostToDevice, H->stream));
    if (is_axpy) ONK_CUDA(cudaMemcpyAsync(xd, x_host + k * chunk, cnt * sizeof(double), cudaMemcpyHostToDevice, H->stream));
    ONK_CUDA(cudaEventRecord(ready[b].e, H->stream));
    ONK_CUDA(cudaStreamWaitEvent(K->stream, ready[b].e, 0));
    if (is_axpy) { if (int rc = onk_axpy_f64(yd, xd, a, cnt, LANE_COMPUTE)) return rc; }
    else         { if (int rc = onk_value_add_f64(yd, 1, cnt, a, LANE_COMPUTE)) return rc; }
    ONK_CUDA(cudaEventRecord(done[b].e, K->stream));
    ONK_CUDA(cudaStreamWaitEvent(D->stream, done[b].e, 0));
    ONK_CUDA(cudaMemcpyAsync(y_host + k * chunk, yd, cnt * sizeof(double), cudaMemcpyDeviceToHost, D->stream));
    ONK_CUDA(cudaEventRecord(drained[b].e, D->stream));
  }
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
