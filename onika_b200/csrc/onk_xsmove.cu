// onk_xsmove.cu -- id-based redistribution of elements between the GPUs of a node.
//
// Reference behaviour: communication_scheme_from_ids (src/mpi/xs_data_move.cpp:42-462) builds, through a rendezvous rank
// per id range and three MPI_Alltoall(v) rounds, the tables that data_move (include/onika/mpi/xs_data_move_impl.h:44-113)
// then uses to pack, MPI_Alltoallv and unpack every array: after[d][j] = before[s][i] wherever ids_after[d][j] ==
// ids_before[s][i].
//
// B200 design: no pack / exchange / unpack.  Every rank owns
//   * a DIRECTORY SEGMENT: one 64-bit word per id of the id range it is the rendezvous rank of (same range split as the
//     reference: xs_data_move_range_utils.h:39-63), and
//   * one OUTPUT WINDOW per moved array,
// both plain device memory mapped into every peer through CUDA IPC (NVLink / NVSwitch peer access).
//   plan : publish -- each rank CASes (rank, j) of every `after` id into the id's directory word on its rendezvous rank;
//          barrier; lookup -- each rank fetches the word of every `before` id (marking it consumed) = its ROUTE (d, j);
//          barrier; verify -- each rank checks that every word of its segment was published once and consumed once (the
//          reference aborts on an id without source or destination: .cpp:286-289,331-334,366-369); the error code is
//          max-combined over the ranks so every rank returns the same status.
//   move : barrier; one kernel stores element i straight into window[d] at j (peer stores, 16/8/4/1-byte pieces, coalesced
//          when neighbours travel together, which is the common case for particles that migrate a little); barrier.
// Barriers are the device-side flag exchange of onk_reduce.cu (st.release.sys / ld.acquire.sys), in stream order.
#include "onk_internal.cuh"

namespace onk
{
  constexpr unsigned long long XS_CONSUMED = 1ull << 63;
  constexpr int XS_OWNER_SHIFT = 44;                              // word = consumed | (owner+1) << 44 | index (44 bits: 16 Ti elements)
  constexpr unsigned long long XS_INDEX_MASK = (1ull << XS_OWNER_SHIFT) - 1ull;

  // several defects can show at once; codes are max-combined (kernels and ranks), so the most fundamental one has the highest value
  enum { XS_OK = 0, XS_ERR_NO_SOURCE = 1, XS_ERR_NO_DEST = 2, XS_ERR_DUP_BEFORE = 3, XS_ERR_DUP_AFTER = 4, XS_ERR_RANGE = 5 };

  struct XsGeometry
  {
    unsigned long long id_min, id_max;
    int rank, world;
    unsigned long long* dir[ONK_XCHG_MAX_WORLD];                   // dir[p] = rank p's directory segment as seen from this device
  };

  // xs_data_move_range_utils.h:39-55
  __host__ __device__ __forceinline__ int xs_process_for_id(unsigned long long id, const XsGeometry& g)
  { return (int)(((id - g.id_min + 1ull) * (unsigned long long)g.world - 1ull) / (g.id_max - g.id_min)); }
  __host__ __device__ __forceinline__ unsigned long long xs_range_start(int p, const XsGeometry& g)
  { return g.id_min + ((unsigned long long)p * (g.id_max - g.id_min)) / (unsigned long long)g.world; }

  __global__ void __launch_bounds__(256)
  xs_publish_kernel(const unsigned long long* __restrict__ ids_after, unsigned long long n_after, const __grid_constant__ XsGeometry g, int* err)
  {
    for (unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; j < n_after; j += (unsigned long long)gridDim.x * blockDim.x)
    {
      const unsigned long long id = ids_after[j];
      if (id < g.id_min || id >= g.id_max) { atomicMax(err, XS_ERR_RANGE); continue; }
      const int p = xs_process_for_id(id, g);
      const unsigned long long word = ((unsigned long long)(g.rank + 1) << XS_OWNER_SHIFT) | j;
      const unsigned long long old = atomicCAS_system(g.dir[p] + (id - xs_range_start(p, g)), 0ull, word);
      if (old != 0ull) atomicMax(err, XS_ERR_DUP_AFTER);
    }
  }

  __global__ void __launch_bounds__(256)
  xs_lookup_kernel(const unsigned long long* __restrict__ ids_before, unsigned long long n_before, const __grid_constant__ XsGeometry g,
                   unsigned long long* __restrict__ route, unsigned long long* __restrict__ send_count, int* err)
  {
    __shared__ unsigned long long s_count[ONK_XCHG_MAX_WORLD];
    if (threadIdx.x < ONK_XCHG_MAX_WORLD) s_count[threadIdx.x] = 0ull;
    __syncthreads();
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_before; i += (unsigned long long)gridDim.x * blockDim.x)
    {
      const unsigned long long id = ids_before[i];
      unsigned long long r = 0ull;
      if (id < g.id_min || id >= g.id_max) atomicMax(err, XS_ERR_RANGE);
      else
      {
        const int p = xs_process_for_id(id, g);
        const unsigned long long old = atomicOr_system(g.dir[p] + (id - xs_range_start(p, g)), XS_CONSUMED);
        if (old == 0ull) atomicMax(err, XS_ERR_NO_DEST);
        else if (old & XS_CONSUMED) atomicMax(err, XS_ERR_DUP_BEFORE);
        else { r = old; atomicAdd(&s_count[(int)(old >> XS_OWNER_SHIFT) - 1], 1ull); }
      }
      route[i] = r;
    }
    __syncthreads();
    if (threadIdx.x < g.world && s_count[threadIdx.x] != 0ull) atomicAdd(&send_count[threadIdx.x], s_count[threadIdx.x]);
  }

  // every word of the local segment must be (published, consumed); err_f64 receives the code for the cross-rank combine
  __global__ void __launch_bounds__(256)
  xs_verify_kernel(const unsigned long long* __restrict__ segment, unsigned long long n, int* err)
  {
    for (unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (unsigned long long)gridDim.x * blockDim.x)
    {
      const unsigned long long w = segment[k];
      if ((w & XS_CONSUMED) == 0ull) atomicMax(err, XS_ERR_NO_SOURCE);            // nobody holds this id before
      else if ((w & ~XS_CONSUMED) == 0ull) atomicMax(err, XS_ERR_NO_DEST);       // held before, wanted by nobody after
    }
  }
  __global__ void xs_err_to_f64_kernel(const int* err, double* out) { *out = (double)*err; }

  struct XsWindows { unsigned char* out[ONK_XCHG_MAX_WORLD]; };

  // 32-byte piece moved with one 256-bit load and one 256-bit store
  struct alignas(32) XsPiece32 { unsigned long long a, b, c, d; };
  template<class P> __device__ __forceinline__ P xs_load(const P* p) { return *p; }
  template<class P> __device__ __forceinline__ void xs_store(P* p, const P& v) { *p = v; }
  template<> __device__ __forceinline__ XsPiece32 xs_load<XsPiece32>(const XsPiece32* p)
  {
    XsPiece32 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v.a), "=l"(v.b), "=l"(v.c), "=l"(v.d) : "l"(p));
    return v;
  }
  template<> __device__ __forceinline__ void xs_store<XsPiece32>(XsPiece32* p, const XsPiece32& v)
  {
    asm volatile("st.global.L1::no_allocate.v4.u64 [%0], {%1,%2,%3,%4};" :: "l"(p), "l"(v.a), "l"(v.b), "l"(v.c), "l"(v.d) : "memory");
  }

  // element i of `in` -> window[d] at index j, as pieces of sizeof(P) bytes; thread t handles piece t of the flat stream.
  // XS_UNROLL pieces per thread and step are loaded (route and payload) before the first store: the stores go to other
  // GPUs, the loads keep the local HBM busy meanwhile.
  constexpr int XS_UNROLL = 4;
  template<class P>
  __global__ void __launch_bounds__(256)
  xs_scatter_kernel(const P* __restrict__ in, unsigned long long n_before, unsigned pieces_per_elem,
                    const unsigned long long* __restrict__ route, const __grid_constant__ XsWindows w)
  {
    const unsigned long long total = n_before * pieces_per_elem;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long t0 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t0 < total; t0 += stride * XS_UNROLL)
    {
      P v[XS_UNROLL]; P* dst[XS_UNROLL];
#pragma unroll
      for (int u = 0; u < XS_UNROLL; u++)
      {
        const unsigned long long t = t0 + (unsigned long long)u * stride;
        dst[u] = nullptr;
        if (t < total)
        {
          const unsigned long long i = (pieces_per_elem == 1u) ? t : t / pieces_per_elem;
          const unsigned c = (unsigned)(t - i * pieces_per_elem);
          const unsigned long long r = __ldg(route + i);
          dst[u] = reinterpret_cast<P*>(w.out[(int)(r >> XS_OWNER_SHIFT) - 1]) + (r & XS_INDEX_MASK) * pieces_per_elem + c;
          v[u] = xs_load(in + t);
        }
      }
#pragma unroll
      for (int u = 0; u < XS_UNROLL; u++) if (dst[u] != nullptr) xs_store(dst[u], v[u]);
    }
  }
}

using namespace onk;

/* ---------------------------------------------- peer-mapped windows ---------------------------------------------- */

struct onk_window
{
  void*  local = nullptr;
  size_t bytes = 0;
  int    rank = 0, world = 1;
  void*  peer[ONK_XCHG_MAX_WORLD] = {};
  void*  opened[ONK_XCHG_MAX_WORLD] = {};
  bool   connected = false;
};

struct onk_xs_plan
{
  onk_xchg*   x = nullptr;
  XsGeometry  geo{};
  onk_window* dir = nullptr;
  unsigned long long  seg_size = 0;
  unsigned long long* route = nullptr;        size_t route_cap = 0;
  unsigned long long* send_count = nullptr;   // world counters (device)
  int*        err = nullptr;                  // device
  double*     err_f64 = nullptr;              // device: [0] local code, [1] max over ranks
  unsigned long long n_before = 0, n_after = 0;
  bool        built = false;
};

extern "C" {

int onk_window_create(size_t bytes, int rank, int world, onk_window** w)
{
  ONK_REQUIRE(w != nullptr, "onk_window_create: null argument");
  ONK_REQUIRE(world >= 1 && world <= ONK_XCHG_MAX_WORLD && rank >= 0 && rank < world, "onk_window_create: bad rank/world %d/%d", rank, world);
  if (int rc = require_init()) return rc;
  onk_window* h = new onk_window();
  h->bytes = bytes; h->rank = rank; h->world = world;
  const size_t alloc = bytes < 256 ? 256 : bytes;
  cudaError_t e = cudaMalloc(&h->local, alloc);               // plain cudaMalloc: exportable through CUDA IPC
  if (e == cudaSuccess) e = cudaMemset(h->local, 0, alloc);
  if (e != cudaSuccess) { delete h; ONK_CUDA(e); }
  h->peer[rank] = h->local;
  h->connected = (world == 1);
  *w = h;
  return ONK_OK;
}

int onk_window_export(onk_window* w, unsigned char handle[ONK_IPC_HANDLE_BYTES])
{
  ONK_REQUIRE(w != nullptr && handle != nullptr, "onk_window_export: null argument");
  cudaIpcMemHandle_t h;
  ONK_CUDA(cudaIpcGetMemHandle(&h, w->local));
  memcpy(handle, &h, sizeof(h));
  return ONK_OK;
}

int onk_window_connect(onk_window* w, const unsigned char* handles)
{
  ONK_REQUIRE(w != nullptr && handles != nullptr, "onk_window_connect: null argument");
  for (int r = 0; r < w->world; r++)
  {
    if (r == w->rank || w->opened[r] != nullptr) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * ONK_IPC_HANDLE_BYTES, sizeof(h));
    void* p = nullptr;
    ONK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    w->opened[r] = p; w->peer[r] = p;
  }
  w->connected = true;
  return ONK_OK;
}

int onk_window_ptr(onk_window* w, void** local_dev, size_t* bytes)
{
  ONK_REQUIRE(w != nullptr, "onk_window_ptr: null window");
  if (local_dev) *local_dev = w->local;
  if (bytes) *bytes = w->bytes;
  return ONK_OK;
}

int onk_window_destroy(onk_window* w)
{
  if (!w) return ONK_OK;
  cudaDeviceSynchronize();
  for (int r = 0; r < ONK_XCHG_MAX_WORLD; r++) if (w->opened[r]) cudaIpcCloseMemHandle(w->opened[r]);
  if (w->local) cudaFree(w->local);
  delete w;
  return ONK_OK;
}

/* ---------------------------------------------- redistribution plan ---------------------------------------------- */

int onk_xs_plan_create(onk_xchg* x, uint64_t id_min, uint64_t id_max, onk_xs_plan** plan)
{
  ONK_REQUIRE(x != nullptr && plan != nullptr, "onk_xs_plan_create: null argument");
  ONK_REQUIRE(id_max > id_min, "onk_xs_plan_create: empty id range [%llu,%llu)", (unsigned long long)id_min, (unsigned long long)id_max);
  int rank = 0, world = 1;
  if (int rc = onk_xchg_info(x, &rank, &world)) return rc;
  ONK_REQUIRE((id_max - id_min) <= (~0ull) / (unsigned long long)(world + 1), "onk_xs_plan_create: id range too large");
  onk_xs_plan* p = new onk_xs_plan();
  p->x = x;
  p->geo.id_min = id_min; p->geo.id_max = id_max; p->geo.rank = rank; p->geo.world = world;
  p->seg_size = xs_range_start(rank + 1, p->geo) - xs_range_start(rank, p->geo);     // range_start(world) == id_max
  int rc = onk_window_create((size_t)p->seg_size * 8, rank, world, &p->dir);
  cudaError_t e = cudaSuccess;
  if (rc == ONK_OK)
  {
    e = cudaMalloc((void**)&p->send_count, sizeof(unsigned long long) * ONK_XCHG_MAX_WORLD);
    if (e == cudaSuccess) e = cudaMalloc((void**)&p->err, 256);
    if (e == cudaSuccess) e = cudaMalloc((void**)&p->err_f64, 256);
  }
  if (rc != ONK_OK || e != cudaSuccess)
  {
    if (p->dir) onk_window_destroy(p->dir);
    if (p->send_count) cudaFree(p->send_count);
    if (p->err) cudaFree(p->err);
    if (p->err_f64) cudaFree(p->err_f64);
    delete p;
    if (rc != ONK_OK) return rc;
    ONK_CUDA(e);
  }
  p->geo.dir[rank] = (unsigned long long*)p->dir->local;
  *plan = p;
  return ONK_OK;
}

int onk_xs_plan_export(onk_xs_plan* plan, unsigned char handle[ONK_IPC_HANDLE_BYTES])
{
  ONK_REQUIRE(plan != nullptr, "onk_xs_plan_export: null plan");
  return onk_window_export(plan->dir, handle);
}

int onk_xs_plan_connect(onk_xs_plan* plan, const unsigned char* handles)
{
  ONK_REQUIRE(plan != nullptr, "onk_xs_plan_connect: null plan");
  if (int rc = onk_window_connect(plan->dir, handles)) return rc;
  for (int r = 0; r < plan->geo.world; r++) plan->geo.dir[r] = (unsigned long long*)plan->dir->peer[r];
  return ONK_OK;
}

int onk_xs_plan_build(onk_xs_plan* plan, const uint64_t* ids_before_dev, uint64_t n_before,
                      const uint64_t* ids_after_dev, uint64_t n_after, int lane, int* status)
{
  ONK_REQUIRE(plan != nullptr && plan->dir->connected, "onk_xs_plan_build: plan is not connected");
  ONK_REQUIRE((n_before == 0 || ids_before_dev != nullptr) && (n_after == 0 || ids_after_dev != nullptr), "onk_xs_plan_build: null id array");
  ONK_REQUIRE(n_before <= XS_INDEX_MASK && n_after <= XS_INDEX_MASK, "onk_xs_plan_build: too many elements");
  Lane* L; if (int rc = lane_get(lane, &L)) return rc;
  cudaStream_t st = L->stream;
  if (plan->route_cap < n_before)
  {
    if (plan->route) { ONK_CUDA(cudaStreamSynchronize(st)); ONK_CUDA(cudaFree(plan->route)); plan->route = nullptr; plan->route_cap = 0; }
    ONK_CUDA(cudaMalloc((void**)&plan->route, sizeof(unsigned long long) * (size_t)(n_before ? n_before : 1)));
    plan->route_cap = (size_t)n_before;
  }
  plan->built = false;
  plan->n_before = n_before; plan->n_after = n_after;
  // 1. clean directory segment and counters; nobody may publish into a segment that is still being cleared
  ONK_CUDA(cudaMemsetAsync(plan->dir->local, 0, (size_t)(plan->seg_size ? plan->seg_size * 8 : 8), st));
  ONK_CUDA(cudaMemsetAsync(plan->send_count, 0, sizeof(unsigned long long) * ONK_XCHG_MAX_WORLD, st));
  ONK_CUDA(cudaMemsetAsync(plan->err, 0, 4, st));
  if (int rc = onk_xchg_barrier(plan->x, lane)) return rc;
  // 2. publish destinations
  if (n_after > 0)
  {
    xs_publish_kernel<<<grid_for(n_after, 256, 8), 256, 0, st>>>((const unsigned long long*)ids_after_dev, n_after, plan->geo, plan->err);
    ONK_CHECK_LAUNCH(); count_launch();
  }
  if (int rc = onk_xchg_barrier(plan->x, lane)) return rc;
  // 3. look routes up
  if (n_before > 0)
  {
    xs_lookup_kernel<<<grid_for(n_before, 256, 8), 256, 0, st>>>((const unsigned long long*)ids_before_dev, n_before, plan->geo, plan->route, plan->send_count, plan->err);
    ONK_CHECK_LAUNCH(); count_launch();
  }
  if (int rc = onk_xchg_barrier(plan->x, lane)) return rc;
  // 4. verify the local segment, share the verdict
  if (plan->seg_size > 0)
  {
    xs_verify_kernel<<<grid_for(plan->seg_size, 256, 8), 256, 0, st>>>((const unsigned long long*)plan->dir->local, plan->seg_size, plan->err);
    ONK_CHECK_LAUNCH(); count_launch();
  }
  xs_err_to_f64_kernel<<<1, 1, 0, st>>>(plan->err, plan->err_f64);
  ONK_CHECK_LAUNCH(); count_launch();
  if (int rc = onk_reduce_xchg_f64(ONK_OP_MAX, plan->err_f64, 1, plan->err_f64 + 1, plan->x, lane)) return rc;
  double verdict[2] = {0.0, 0.0};
  ONK_CUDA(cudaMemcpyAsync(verdict, plan->err_f64, sizeof(verdict), cudaMemcpyDeviceToHost, st));
  ONK_CUDA(cudaStreamSynchronize(st));
  const int code = (int)verdict[1];
  if (status) *status = code;
  if (code != XS_OK)
  {
    static const char* what[] = { "", "an id of the range has no source", "an id has no destination", "an id appears twice in the distribution before",
                                  "an id appears twice in the distribution after", "an id lies outside [id_min,id_max)" };
    set_error("onk_xs_plan_build: %s (code %d, same on every rank)", what[code >= 1 && code <= 5 ? code : 0], code);
    return ONK_ERR_ARG;
  }
  plan->built = true;
  return ONK_OK;
}

int onk_xs_plan_send_counts(onk_xs_plan* plan, uint64_t* send_count_host)
{
  ONK_REQUIRE(plan != nullptr && plan->built && send_count_host != nullptr, "onk_xs_plan_send_counts: plan was not built");
  unsigned long long tmp[ONK_XCHG_MAX_WORLD];
  ONK_CUDA(cudaMemcpy(tmp, plan->send_count, sizeof(tmp), cudaMemcpyDeviceToHost));
  for (int r = 0; r < plan->geo.world; r++) send_count_host[r] = tmp[r];
  return ONK_OK;
}

// argument checks and the scatter launch of one array (no barriers)
static int xs_check_one(onk_xs_plan* plan, onk_window* out, uint64_t elem_bytes, const void* in_dev)
{
  ONK_REQUIRE(out != nullptr && out->connected && out->world == plan->geo.world && out->rank == plan->geo.rank, "onk_xs_move: output window is not connected to the plan's ranks");
  ONK_REQUIRE(elem_bytes > 0 && (plan->n_before == 0 || in_dev != nullptr), "onk_xs_move: bad element size or null input");
  ONK_REQUIRE(out->bytes >= plan->n_after * elem_bytes, "onk_xs_move: output window too small (%zu < %llu)", out->bytes, (unsigned long long)(plan->n_after * elem_bytes));
  return ONK_OK;
}

static int xs_scatter_one(onk_xs_plan* plan, onk_window* out, uint64_t elem_bytes, const void* in_dev, Lane* L)
{
  if (plan->n_before == 0) return ONK_OK;
  XsWindows w{};
  for (int r = 0; r < plan->geo.world; r++) w.out[r] = (unsigned char*)out->peer[r];
  const uintptr_t a = (uintptr_t)in_dev;
  const unsigned piece = (elem_bytes % 32 == 0 && a % 32 == 0) ? 32u : (elem_bytes % 16 == 0 && a % 16 == 0) ? 16u
                       : (elem_bytes % 8 == 0 && a % 8 == 0) ? 8u : (elem_bytes % 4 == 0 && a % 4 == 0) ? 4u : 1u;
  const unsigned ppe = (unsigned)(elem_bytes / piece);
  const unsigned grid = grid_for(plan->n_before * ppe, 256 * XS_UNROLL, 8);
  switch (piece)
  {
    case 32: xs_scatter_kernel<XsPiece32><<<grid, 256, 0, L->stream>>>((const XsPiece32*)in_dev, plan->n_before, ppe, plan->route, w); break;
    case 16: xs_scatter_kernel<uint4><<<grid, 256, 0, L->stream>>>((const uint4*)in_dev, plan->n_before, ppe, plan->route, w); break;
    case 8:  xs_scatter_kernel<unsigned long long><<<grid, 256, 0, L->stream>>>((const unsigned long long*)in_dev, plan->n_before, ppe, plan->route, w); break;
    case 4:  xs_scatter_kernel<unsigned><<<grid, 256, 0, L->stream>>>((const unsigned*)in_dev, plan->n_before, ppe, plan->route, w); break;
    default: xs_scatter_kernel<unsigned char><<<grid, 256, 0, L->stream>>>((const unsigned char*)in_dev, plan->n_before, ppe, plan->route, w); break;
  }
  ONK_CHECK_LAUNCH(); count_launch();
  return ONK_OK;
}

int onk_xs_move_fields(onk_xs_plan* plan, int nfields, onk_window* const* outs, const uint64_t* elem_bytes, const void* const* ins_dev, int lane)
{
  ONK_REQUIRE(plan != nullptr && plan->built, "onk_xs_move: plan was not built");
  ONK_REQUIRE(nfields >= 1 && outs != nullptr && elem_bytes != nullptr && ins_dev != nullptr, "onk_xs_move_fields: null argument");
  // every argument is checked before the first collective step: a rank that rejects its arguments must not leave the others at a barrier
  for (int k = 0; k < nfields; k++) if (int rc = xs_check_one(plan, outs[k], elem_bytes[k], ins_dev[k])) return rc;
  Lane* L; if (int rc = lane_get(lane, &L)) return rc;
  // the peers' windows may still be read by what they queued before this call
  if (int rc = onk_xchg_barrier(plan->x, lane)) return rc;
  for (int k = 0; k < nfields; k++) if (int rc = xs_scatter_one(plan, outs[k], elem_bytes[k], ins_dev[k], L)) return rc;
  // everything every rank sent has landed when the lane continues
  return onk_xchg_barrier(plan->x, lane);
}

int onk_xs_move(onk_xs_plan* plan, onk_window* out, uint64_t elem_bytes, const void* in_dev, int lane)
{
  onk_window* outs[1] = { out };
  const void* ins[1] = { in_dev };
  return onk_xs_move_fields(plan, 1, outs, &elem_bytes, ins, lane);
}

int onk_xs_plan_destroy(onk_xs_plan* plan)
{
  if (!plan) return ONK_OK;
  cudaDeviceSynchronize();
  if (plan->dir) onk_window_destroy(plan->dir);
  if (plan->route) cudaFree(plan->route);
  if (plan->send_count) cudaFree(plan->send_count);
  if (plan->err) cudaFree(plan->err);
  if (plan->err_f64) cudaFree(plan->err_f64);
  delete plan;
  return ONK_OK;
}

}
