// tools/tune.cu -- parameter sweep of the streaming kernel shapes on a B200 (not part of the product).
// Variants of the reduce (read-only) and read-modify-write tile loops: threads, unroll, CTAs per SM, cache hints.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tune tools/tune.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <algorithm>

#define CK(x) do{ cudaError_t e=(x); if(e!=cudaSuccess){ printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

struct alignas(32) d4 { double a,b,c,d; };
template<int H> __device__ __forceinline__ d4 ld(const double* p)
{
  d4 v;
  if constexpr (H==0) asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.a),"=d"(v.b),"=d"(v.c),"=d"(v.d) : "l"(p));
  else if constexpr (H==1) asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.a),"=d"(v.b),"=d"(v.c),"=d"(v.d) : "l"(p));
  else if constexpr (H==2) asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.a),"=d"(v.b),"=d"(v.c),"=d"(v.d) : "l"(p));
  else if constexpr (H==3) asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.L2::256B.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.a),"=d"(v.b),"=d"(v.c),"=d"(v.d) : "l"(p));
  else asm volatile("ld.global.cs.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.a),"=d"(v.b),"=d"(v.c),"=d"(v.d) : "l"(p));
  return v;
}
template<int H> __device__ __forceinline__ void st(double* p, const d4& v)
{
  if constexpr (H==0) asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p),"d"(v.a),"d"(v.b),"d"(v.c),"d"(v.d) : "memory");
  else if constexpr (H==1) asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p),"d"(v.a),"d"(v.b),"d"(v.c),"d"(v.d) : "memory");
  else if constexpr (H==2) asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p),"d"(v.a),"d"(v.b),"d"(v.c),"d"(v.d) : "memory");
  else asm volatile("st.global.L1::no_allocate.L2::evict_first.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p),"d"(v.a),"d"(v.b),"d"(v.c),"d"(v.d) : "memory");
}

template<int T, int U, int C, int H>
__global__ void __launch_bounds__(T, C) k_reduce(const double* __restrict__ in, uint64_t ntiles, double* out)
{
  double acc[4] = {1e308,1e308,1e308,1e308};
  constexpr unsigned TILE = T*U*4;
  for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x)
  {
    const double* p = in + t*TILE + threadIdx.x*4;
    d4 v[U];
#pragma unroll
    for (int u=0;u<U;u++) v[u] = ld<H>(p + u*(T*4));
#pragma unroll
    for (int u=0;u<U;u++) { acc[0] = v[u].a<acc[0]?v[u].a:acc[0]; acc[1] = v[u].b<acc[1]?v[u].b:acc[1]; acc[2] = v[u].c<acc[2]?v[u].c:acc[2]; acc[3] = v[u].d<acc[3]?v[u].d:acc[3]; }
  }
  double m = fmin(fmin(acc[0],acc[1]),fmin(acc[2],acc[3]));
  for (int o=16;o>0;o>>=1) m = fmin(m, __shfl_xor_sync(0xffffffffu,m,o));
  if ((threadIdx.x&31)==0 && m < 0.0) out[blockIdx.x] = m;   // data is positive: never writes, keeps the loads alive
}

template<int T, int U, int C, int HL, int HS>
__global__ void __launch_bounds__(T, C) k_rmw(double* __restrict__ a, uint64_t ntiles, double s, double c)
{
  constexpr unsigned TILE = T*U*4;
  for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x)
  {
    double* p = a + t*TILE + threadIdx.x*4;
    d4 v[U];
#pragma unroll
    for (int u=0;u<U;u++) { if constexpr (HL==9) asm volatile("ld.global.L1::no_allocate.L2::evict_first.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[u].a),"=d"(v[u].b),"=d"(v[u].c),"=d"(v[u].d) : "l"(p + u*(T*4)) : "memory"); else v[u] = ld<HL>(p + u*(T*4)); }
#pragma unroll
    for (int u=0;u<U;u++) { v[u].a = v[u].a*s+c; v[u].b = v[u].b*s+c; v[u].c = v[u].c*s+c; v[u].d = v[u].d*s+c; st<HS>(p + u*(T*4), v[u]); }
  }
}

template<int T, int U, int C>
__global__ void __launch_bounds__(T, C) k_copy(double* __restrict__ dst, const double* __restrict__ src, uint64_t ntiles)
{
  constexpr unsigned TILE = T*U*4;
  for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x)
  {
    d4 v[U];
#pragma unroll
    for (int u=0;u<U;u++) v[u] = ld<0>(src + t*TILE + threadIdx.x*4 + u*(T*4));
#pragma unroll
    for (int u=0;u<U;u++) st<0>(dst + t*TILE + threadIdx.x*4 + u*(T*4), v[u]);
  }
}

// strided line read: `useful` bytes (multiple of 32) at the start of every `stride` bytes -- the cell-grid access pattern
template<int T, int C>
__global__ void __launch_bounds__(T, C) k_strided(const double* __restrict__ in, uint64_t nlines, unsigned stride_d, unsigned vec_per_line, double* out)
{
  double acc = 1e308;
  const uint64_t nvec = nlines * vec_per_line;
  for (uint64_t i = (uint64_t)blockIdx.x*T + threadIdx.x; i < nvec; i += (uint64_t)gridDim.x*T*4)
  {
    d4 v[4];
#pragma unroll
    for (int u=0;u<4;u++) { uint64_t j = i + (uint64_t)u*gridDim.x*T; if (j < nvec) v[u] = ld<0>(in + (j/vec_per_line)*stride_d + (j%vec_per_line)*4); else v[u] = d4{1e308,1e308,1e308,1e308}; }
#pragma unroll
    for (int u=0;u<4;u++) acc = fmin(acc, fmin(fmin(v[u].a,v[u].b),fmin(v[u].c,v[u].d)));
  }
  if (acc < 0.0) out[blockIdx.x] = acc;
}

// TMA-bulk staged read-modify-write: one elected thread streams 16 KiB tiles global->shared (cp.async.bulk + mbarrier),
// all threads update the tile in shared memory, one thread streams it back (cp.async.bulk shared->global)
template<int T, int STAGES, int TILE_BYTES, int C>
__global__ void __launch_bounds__(T, C) k_rmw_tma(double* __restrict__ a, uint64_t ntiles, double s, double c)
{
  extern __shared__ __align__(128) unsigned char smem[];
  double* buf = reinterpret_cast<double*>(smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES*TILE_BYTES);
  constexpr int TILE_D = TILE_BYTES/8;
  const unsigned tid = threadIdx.x;
  if (tid == 0) { for (int i=0;i<STAGES;i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"((unsigned)__cvta_generic_to_shared(full+i))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  // my tiles: blockIdx.x, +grid, ...
  uint64_t issued = 0, t_issue = blockIdx.x;
  auto issue = [&](int stage, uint64_t tile) {
    unsigned bar = (unsigned)__cvta_generic_to_shared(full+stage);
    unsigned dst = (unsigned)__cvta_generic_to_shared(buf + (size_t)stage*TILE_D);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(TILE_BYTES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(dst), "l"(a + tile*TILE_D), "r"(TILE_BYTES), "r"(bar) : "memory");
  };
  if (tid == 0) { for (int sidx=0; sidx<STAGES-1 && t_issue < ntiles; sidx++, t_issue += gridDim.x, issued++) issue(sidx, t_issue); }
  uint64_t k = 0;
  for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x, k++)
  {
    const int stage = k % STAGES; const unsigned phase = (k / STAGES) & 1;
    // keep the pipeline full: the stage being refilled was stored (and its store drained) one iteration ago
    if (tid == 0 && t_issue < ntiles)
    {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous store has finished reading shared memory
      issue((int)(issued % STAGES), t_issue); t_issue += gridDim.x; issued++;
    }
    unsigned bar = (unsigned)__cvta_generic_to_shared(full+stage);
    unsigned ok = 0;
    while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p; }" : "=r"(ok) : "r"(bar), "r"(phase) : "memory");
    double* tile = buf + (size_t)stage*TILE_D;
    for (int i = tid*2; i < TILE_D; i += T*2) { double2 v = *reinterpret_cast<double2*>(tile+i); v.x = v.x*s+c; v.y = v.y*s+c; *reinterpret_cast<double2*>(tile+i) = v; }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0)
    {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(a + t*TILE_D), "r"((unsigned)__cvta_generic_to_shared(tile)), "r"(TILE_BYTES) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// contiguous-chunk assignment instead of grid-stride tiles
template<int T, int U, int C>
__global__ void __launch_bounds__(T, C) k_rmw_chunked(double* __restrict__ a, uint64_t ntiles, double s, double c)
{
  constexpr unsigned TILE = T*U*4;
  const uint64_t per = (ntiles + gridDim.x - 1) / gridDim.x;
  const uint64_t t0 = blockIdx.x * per, t1 = (t0 + per < ntiles) ? t0 + per : ntiles;
  for (uint64_t t = t0; t < t1; t++)
  {
    double* p = a + t*TILE + threadIdx.x*4;
    d4 v[U];
#pragma unroll
    for (int u=0;u<U;u++) asm volatile("ld.global.L1::no_allocate.L2::evict_first.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[u].a),"=d"(v[u].b),"=d"(v[u].c),"=d"(v[u].d) : "l"(p + u*(T*4)) : "memory");
#pragma unroll
    for (int u=0;u<U;u++) { v[u].a = v[u].a*s+c; v[u].b = v[u].b*s+c; v[u].c = v[u].c*s+c; v[u].d = v[u].d*s+c; st<0>(p + u*(T*4), v[u]); }
  }
}

// width sweep: W doubles per access (1 = 8 B scalar, 2 = 16 B, 4 = 32 B), EPT independent accesses per thread per tile
template<int W> struct vec_t;
template<> struct vec_t<1> { double v[1]; };
template<> struct alignas(16) vec_t<2> { double v[2]; };
template<> struct alignas(32) vec_t<4> { double v[4]; };
template<int W> __device__ __forceinline__ vec_t<W> ldw(const double* p)
{
  vec_t<W> r;
  if constexpr (W==1) asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(r.v[0]) : "l"(p) : "memory");
  else if constexpr (W==2) asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.v[0]),"=d"(r.v[1]) : "l"(p) : "memory");
  else asm volatile("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.v[0]),"=d"(r.v[1]),"=d"(r.v[2]),"=d"(r.v[3]) : "l"(p) : "memory");
  return r;
}
template<int W> __device__ __forceinline__ void stw(double* p, const vec_t<W>& r)
{
  if constexpr (W==1) asm volatile("st.global.f64 [%0], %1;" :: "l"(p),"d"(r.v[0]) : "memory");
  else if constexpr (W==2) asm volatile("st.global.v2.f64 [%0], {%1,%2};" :: "l"(p),"d"(r.v[0]),"d"(r.v[1]) : "memory");
  else asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p),"d"(r.v[0]),"d"(r.v[1]),"d"(r.v[2]),"d"(r.v[3]) : "memory");
}
template<int T, int W, int EPT, int C>
__global__ void __launch_bounds__(T, C) k_rmw_w(double* __restrict__ a, uint64_t ntiles, double s, double c)
{
  constexpr unsigned TILE = T*W*EPT;
  for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x)
  {
    double* p = a + t*TILE + threadIdx.x*W;
    vec_t<W> v[EPT];
#pragma unroll
    for (int e=0;e<EPT;e++) v[e] = ldw<W>(p + e*(T*W));
#pragma unroll
    for (int e=0;e<EPT;e++) { for (int k=0;k<W;k++) v[e].v[k] = v[e].v[k]*s+c; stw<W>(p + e*(T*W), v[e]); }
  }
}
template<int T, int W, int EPT, int C>
__global__ void __launch_bounds__(T, C) k_axpy_w(double* __restrict__ y, const double* __restrict__ x, uint64_t ntiles, double a)
{
  constexpr unsigned TILE = T*W*EPT;
  for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x)
  {
    const uint64_t o = t*TILE + threadIdx.x*W;
    vec_t<W> vx[EPT], vy[EPT];
#pragma unroll
    for (int e=0;e<EPT;e++) { vx[e] = ldw<W>(x + o + e*(T*W)); vy[e] = ldw<W>(y + o + e*(T*W)); }
#pragma unroll
    for (int e=0;e<EPT;e++) { for (int k=0;k<W;k++) vy[e].v[k] = vy[e].v[k] + a*vx[e].v[k]; stw<W>(y + o + e*(T*W), vy[e]); }
  }
}

template<int T, int W, int EPT, int C, int HS>
__global__ void __launch_bounds__(T, C) k_set_w(double* __restrict__ a, uint64_t ntiles, double v)
{
  constexpr unsigned TILE = T*W*EPT;
  vec_t<W> r; for (int k=0;k<W;k++) r.v[k] = v;
  for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x)
  {
    double* p = a + t*TILE + threadIdx.x*W;
#pragma unroll
    for (int e=0;e<EPT;e++)
    {
      if constexpr (HS==0) stw<W>(p + e*(T*W), r);
      else { static_assert(W==4 || HS==0); d4 q{v,v,v,v}; st<HS-1>(p + e*(T*W), q); }
    }
  }
}

// memset through TMA bulk stores: the tile is written to shared memory once, then one thread per CTA streams it out
template<int T, int TILE_BYTES, int C, int DEPTH>
__global__ void __launch_bounds__(T, C) k_set_tma(double* __restrict__ a, uint64_t ntiles, double v)
{
  extern __shared__ __align__(128) unsigned char smem[];
  double* tile = reinterpret_cast<double*>(smem);
  for (int i = threadIdx.x; i < TILE_BYTES/8; i += T) tile[i] = v;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0)
  {
    int pending = 0;
    for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x)
    {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(a + t*(TILE_BYTES/8)), "r"((unsigned)__cvta_generic_to_shared(tile)), "r"(TILE_BYTES) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (++pending >= DEPTH) { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(DEPTH-1) : "memory"); pending = DEPTH-1; }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

static cudaEvent_t e0, e1;
template<class F> double timeit(F f, int reps = 20)
{
  for (int i=0;i<3;i++) f();
  CK(cudaEventRecord(e0));
  for (int i=0;i<reps;i++) f();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms,e0,e1));
  CK(cudaGetLastError());
  return ms/reps*1e-3;
}

// TMA-bulk staged read-only reduction: one elected thread streams TB-byte tiles global->shared (cp.async.bulk + mbarrier)
// into a ring of ST stages, all threads fold the tile from shared memory (conflict-free LDS.128)
template<int T, int ST, int TB, int C>
__global__ void __launch_bounds__(T, C) k_reduce_tma(const double* __restrict__ in, uint64_t ntiles, double* out)
{
  extern __shared__ __align__(128) unsigned char rsm[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(rsm + (size_t)ST * TB);
  const unsigned tid = threadIdx.x;
  if (tid == 0)
  {
    for (int s = 0; s < ST; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"((unsigned)__cvta_generic_to_shared(bars + s)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint64_t mine = (blockIdx.x < ntiles) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto issue = [&](uint64_t k)
  {
    const int s = (int)(k % ST);
    const unsigned bar = (unsigned)__cvta_generic_to_shared(bars + s);
    const unsigned dst = (unsigned)__cvta_generic_to_shared(rsm + (size_t)s * TB);
    const uint64_t tile = blockIdx.x + k * gridDim.x;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"((unsigned)TB) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(dst), "l"(in + tile * (TB / 8)), "r"((unsigned)TB), "r"(bar) : "memory");
  };
  if (tid == 0) for (uint64_t k = 0; k < (uint64_t)ST && k < mine; k++) issue(k);
  double a0 = 1e308, a1 = 1e308, a2 = 1e308, a3 = 1e308;
  for (uint64_t k = 0; k < mine; k++)
  {
    const int s = (int)(k % ST);
    const unsigned phase = (unsigned)((k / ST) & 1);
    const unsigned bar = (unsigned)__cvta_generic_to_shared(bars + s);
    unsigned done = 0;
    while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(phase) : "memory");
    const unsigned base = (unsigned)__cvta_generic_to_shared(rsm + (size_t)s * TB) + tid * 32u;
#pragma unroll
    for (int i = 0; i < TB / (T * 32); i++)
    {
      double x0, x1, x2, x3;
      asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x0), "=d"(x1) : "r"(base + i * (T * 32u)) : "memory");
      asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x2), "=d"(x3) : "r"(base + i * (T * 32u) + 16u) : "memory");
      a0 = x0 < a0 ? x0 : a0; a1 = x1 < a1 ? x1 : a1; a2 = x2 < a2 ? x2 : a2; a3 = x3 < a3 ? x3 : a3;
    }
    __syncthreads();
    if (tid == 0 && k + ST < mine) issue(k + ST);
  }
  double m = fmin(fmin(a0, a1), fmin(a2, a3));
  for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((tid & 31) == 0 && m < 0.0) out[blockIdx.x] = m;
}

// read side through TMA (bulk tiles into a 2-stage ring), update in registers, 256-bit stores straight from registers
template<int T, int ST, int TB, int C>
__global__ void __launch_bounds__(T, C) k_rmw_tmaload(double* __restrict__ a, uint64_t ntiles, double s, double c)
{
  extern __shared__ __align__(128) unsigned char rsm[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(rsm + (size_t)ST * TB);
  const unsigned tid = threadIdx.x;
  if (tid == 0)
  {
    for (int k = 0; k < ST; k++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"((unsigned)__cvta_generic_to_shared(bars + k)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint64_t mine = (blockIdx.x < ntiles) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto issue = [&](uint64_t k)
  {
    const int st_ = (int)(k % ST);
    const unsigned bar = (unsigned)__cvta_generic_to_shared(bars + st_);
    const unsigned dst = (unsigned)__cvta_generic_to_shared(rsm + (size_t)st_ * TB);
    const uint64_t tile = blockIdx.x + k * gridDim.x;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"((unsigned)TB) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(dst), "l"(a + tile * (TB / 8)), "r"((unsigned)TB), "r"(bar) : "memory");
  };
  if (tid == 0) for (uint64_t k = 0; k < (uint64_t)ST && k < mine; k++) issue(k);
  for (uint64_t k = 0; k < mine; k++)
  {
    const int st_ = (int)(k % ST);
    const unsigned phase = (unsigned)((k / ST) & 1);
    const unsigned bar = (unsigned)__cvta_generic_to_shared(bars + st_);
    unsigned done = 0;
    while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(phase) : "memory");
    const unsigned base = (unsigned)__cvta_generic_to_shared(rsm + (size_t)st_ * TB) + tid * 32u;
    double* out = a + (blockIdx.x + k * gridDim.x) * (TB / 8) + tid * 4;
#pragma unroll
    for (int i = 0; i < TB / (T * 32); i++)
    {
      d4 v;
      asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.a), "=d"(v.b) : "r"(base + i * (T * 32u)) : "memory");
      asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.c), "=d"(v.d) : "r"(base + i * (T * 32u) + 16u) : "memory");
      v.a = v.a * s + c; v.b = v.b * s + c; v.c = v.c * s + c; v.d = v.d * s + c;
      st<0>(out + i * (T * 4), v);
    }
    __syncthreads();
    if (tid == 0 && k + ST < mine) issue(k + ST);
  }
}

int main(int argc, char** argv)
{
  const uint64_t n = 1ull<<28;
  double *a, *b, *out;
  CK(cudaMalloc(&a, n*8)); CK(cudaMalloc(&b, n*8)); CK(cudaMalloc(&out, 1<<20));
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  // positive data
  { std::vector<double> h(1<<20); for (size_t i=0;i<h.size();i++) h[i] = 0.25 + (i%1000)*1e-3; for (uint64_t o=0;o<n;o+=h.size()) CK(cudaMemcpy(a+o,h.data(),h.size()*8,cudaMemcpyHostToDevice)); }
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  printf("SMs %d\n", sms);
#define RTMA(T,ST,TB,C) { const uint64_t nt = (n*8)/TB; size_t sm = (size_t)ST*TB + 128; CK(cudaFuncSetAttribute(k_reduce_tma<T,ST,TB,C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); double t = timeit([&]{ k_reduce_tma<T,ST,TB,C><<<sms*C,T,sm>>>(a,nt,out); }); printf("reduce_tma T=%d stages=%d tile=%d C=%d : %8.1f GB/s\n",T,ST,TB,C, n*8/t/1e9); }
#define RTL(T,ST,TB,C) { const uint64_t nt = (n*8)/TB; size_t sm = (size_t)ST*TB + 128; CK(cudaFuncSetAttribute(k_rmw_tmaload<T,ST,TB,C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); double t = timeit([&]{ k_rmw_tmaload<T,ST,TB,C><<<sms*C,T,sm>>>(a,nt,1.0,0.0); }, 10); printf("rmw_tmaload T=%d stages=%d tile=%d C=%d : %8.1f GB/s\n",T,ST,TB,C, 2*n*8/t/1e9); }
  if (argc > 1 && argv[1][0] == 'w')
  {
    for (int rep = 0; rep < 2; rep++)
    {
      { const uint64_t nt = n/(256*1*4); double t = timeit([&]{ k_rmw<256,1,8,9,0><<<sms*8,256>>>(a,nt,1.0,0.0); }, 10); printf("rmw T=256 U=1 C=8 (product shape) : %8.1f GB/s\n", 2*n*8/t/1e9); }
      RTL(256,2,65536,1) RTL(512,2,65536,1) RTL(256,2,32768,2) RTL(256,4,32768,1) RTL(512,4,32768,1) RTL(256,2,32768,3) RTL(256,3,32768,2) RTL(128,2,16384,4) RTL(256,2,16384,4)
    }
    return 0;
  }
  if (argc > 1 && argv[1][0] == 'r')
  {
    for (int rep = 0; rep < 2; rep++)
    {
      { const uint64_t nt = n/(256*4*4); double t = timeit([&]{ k_reduce<256,4,4,0><<<sms*4,256>>>(a,nt,out); }); printf("reduce T=256 U=4 C=4 H=0 : %8.1f GB/s\n", n*8/t/1e9); }
      RTMA(256,2,65536,1) RTMA(512,2,65536,1) RTMA(1024,2,65536,1) RTMA(128,2,65536,1) RTMA(256,3,65536,1) RTMA(512,3,65536,1) RTMA(256,2,98304,1) RTMA(512,2,98304,1) RTMA(256,2,114688,1) RTMA(512,4,32768,1) RTMA(512,6,32768,1) RTMA(1024,4,32768,1) RTMA(256,4,49152,1) RTMA(512,4,49152,1) RTMA(256,2,49152,2) RTMA(256,2,32768,2) RTMA(256,2,32768,3) RTMA(512,2,32768,2)
    }
    return 0;
  }
  { double t = timeit([&]{ CK(cudaMemcpyAsync(b,a,n*8,cudaMemcpyDeviceToDevice)); }, 10); printf("cudaMemcpy D2D           : %8.1f GB/s (r+w)\n", 2*n*8/t/1e9); }

#define RED(T,U,C,H) { const uint64_t nt = n/(T*U*4); double t = timeit([&]{ k_reduce<T,U,C,H><<<sms*C,T>>>(a,nt,out); }); printf("reduce T=%d U=%d C=%d H=%d : %8.1f GB/s\n",T,U,C,H, n*8/t/1e9); }
  RED(256,4,4,0) RED(256,4,2,0) RED(256,4,3,0) RED(256,4,6,0) RED(256,4,8,0)
  RED(256,2,4,0) RED(256,2,8,0) RED(256,8,2,0) RED(256,8,4,0) RED(256,1,8,0)
  RED(512,4,2,0) RED(512,2,4,0) RED(512,4,1,0) RED(128,4,8,0) RED(128,8,4,0) RED(128,4,12,0) RED(1024,2,1,0) RED(1024,2,2,0)
  RED(256,4,4,1) RED(256,4,4,2) RED(256,4,4,3) RED(256,4,4,4) RED(256,8,2,1) RED(256,8,2,3) RED(256,4,6,3)

#define RMW(T,U,C,HL,HS) { const uint64_t nt = n/(T*U*4); double t = timeit([&]{ k_rmw<T,U,C,HL,HS><<<sms*C,T>>>(a,nt,1.0,0.0); }, 10); printf("rmw T=%d U=%d C=%d HL=%d HS=%d : %8.1f GB/s\n",T,U,C,HL,HS, 2*n*8/t/1e9); }
  RMW(256,4,4,9,0) RMW(256,4,2,9,0) RMW(256,4,6,9,0) RMW(256,4,8,9,0) RMW(256,2,4,9,0) RMW(256,2,8,9,0) RMW(256,8,2,9,0) RMW(256,8,4,9,0) RMW(256,1,8,9,0)
  RMW(512,4,2,9,0) RMW(512,2,4,9,0) RMW(128,4,8,9,0) RMW(1024,2,2,9,0)
  RMW(256,4,4,9,1) RMW(256,4,4,9,2) RMW(256,4,4,9,3) RMW(256,4,4,2,1) RMW(256,4,4,4,2) RMW(256,2,8,2,1) RMW(256,8,2,9,3)

#define CPY(T,U,C) { const uint64_t nt = n/(T*U*4); double t = timeit([&]{ k_copy<T,U,C><<<sms*C,T>>>(b,a,nt); }, 10); printf("copy T=%d U=%d C=%d : %8.1f GB/s (r+w)\n",T,U,C, 2*n*8/t/1e9); }
  CPY(256,4,4) CPY(256,2,8) CPY(256,8,2) CPY(512,4,2) CPY(256,4,8)

#define RW(T,W,EPT,C) { const uint64_t nt = n/(T*W*EPT); double t = timeit([&]{ k_rmw_w<T,W,EPT,C><<<sms*C,T>>>(a,nt,1.0,0.0); }, 10); printf("rmw_w T=%d W=%d EPT=%d C=%d : %8.1f GB/s\n",T,W,EPT,C, 2*n*8/t/1e9); }
  RW(128,1,4,8) RW(128,4,4,4) RW(128,1,4,8) RW(128,4,4,4) RW(128,1,4,8) RW(256,4,4,4) RW(128,1,4,6) RW(128,1,4,10) RW(128,1,4,7) RW(128,1,4,9) RW(64,1,4,16) RW(64,1,8,16) RW(128,1,6,8) RW(128,1,3,8) RW(128,1,5,8) RW(128,2,2,8) RW(128,2,4,8) RW(128,2,4,4) RW(128,4,1,8) RW(128,4,2,8) RW(128,4,2,6) RW(128,4,4,8) RW(256,1,2,8) RW(256,1,4,4)
#define SET(T,W,EPT,C,HS) { const uint64_t nt = n/(T*W*EPT); double t = timeit([&]{ k_set_w<T,W,EPT,C,HS><<<sms*C,T>>>(a,nt,1.5); }, 10); printf("set_w T=%d W=%d EPT=%d C=%d HS=%d : %8.1f GB/s\n",T,W,EPT,C,HS, n*8/t/1e9); }
  SET(256,4,4,4,0) SET(256,4,4,4,1) SET(256,4,4,4,2) SET(256,4,4,4,3) SET(256,4,4,4,4) SET(256,4,1,4,0) SET(256,4,1,8,0) SET(256,4,2,4,0) SET(128,4,1,8,0) SET(128,1,4,8,0) SET(128,2,4,8,0) SET(256,4,8,2,0) SET(256,4,4,8,0) SET(512,4,4,2,0) SET(128,4,4,16,0) SET(256,4,1,2,0) SET(256,4,16,1,0)
#define SETT(T,TB,C,D) { const uint64_t nt = (n*8)/TB; CK(cudaFuncSetAttribute(k_set_tma<T,TB,C,D>, cudaFuncAttributeMaxDynamicSharedMemorySize, TB)); double t = timeit([&]{ k_set_tma<T,TB,C,D><<<sms*C,T,TB>>>(a,nt,1.5); }, 10); printf("set_tma T=%d tile=%d C=%d depth=%d : %8.1f GB/s\n",T,TB,C,D, n*8/t/1e9); }
  SETT(128,16384,2,4) SETT(128,32768,2,4) SETT(128,65536,1,4) SETT(128,8192,4,8) SETT(128,16384,4,8) SETT(128,32768,1,2) SETT(128,4096,8,8) SETT(128,16384,1,16) SETT(32,16384,4,8)
  { double t = timeit([&]{ CK(cudaMemsetAsync(a, 0x3c, n*8)); }, 10); printf("cudaMemset(0x3c) : %8.1f GB/s\n", n*8/t/1e9); }
  { double t = timeit([&]{ CK(cudaMemsetAsync(a, 0, n*8)); }, 10); printf("cudaMemset : %8.1f GB/s\n", n*8/t/1e9); }

#define AX(T,W,EPT,C) { const uint64_t nh = n/2; const uint64_t nt = nh/(T*W*EPT); double t = timeit([&]{ k_axpy_w<T,W,EPT,C><<<sms*C,T>>>(a,a+nh,nt,0.5); }, 10); printf("axpy_w T=%d W=%d EPT=%d C=%d : %8.1f GB/s\n",T,W,EPT,C, 3*nh*8/t/1e9); }
  AX(128,1,4,16) AX(128,1,4,8) AX(128,1,8,8) AX(256,1,4,8) AX(128,2,4,8) AX(128,2,2,12) AX(256,2,4,4) AX(256,4,2,4) AX(256,4,1,8) AX(128,4,2,8) AX(512,1,4,4) AX(1024,1,2,2)

#define STR(STRIDE,USEFUL) { const uint64_t nl = (n*8)/STRIDE; double t = timeit([&]{ k_strided<256,4><<<sms*4,256>>>(a,nl,STRIDE/8,USEFUL/32,out); }, 10); printf("strided useful=%d of stride=%d : %8.1f GB/s useful (%.1f us for %llu lines)\n",USEFUL,STRIDE, nl*(double)USEFUL/t/1e9, t*1e6, (unsigned long long)nl); }
  STR(128,128) STR(256,128) STR(512,128) STR(1024,128) STR(512,256) STR(1024,256) STR(512,64) STR(2048,128)

#define TMA(T,ST,TB,C) { const uint64_t nt = (n*8)/TB; size_t sm = (size_t)ST*TB + 64; CK(cudaFuncSetAttribute(k_rmw_tma<T,ST,TB,C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); double t = timeit([&]{ k_rmw_tma<T,ST,TB,C><<<sms*C,T,sm>>>(a,nt,1.0,0.0); }, 10); printf("rmw_tma T=%d stages=%d tile=%d C=%d : %8.1f GB/s\n",T,ST,TB,C, 2*n*8/t/1e9); }
  TMA(128,4,16384,2) TMA(128,3,16384,3) TMA(256,4,16384,2) TMA(128,4,32768,1) TMA(256,6,16384,2) TMA(128,8,8192,2) TMA(256,3,32768,2) TMA(128,2,32768,3)

#define CHK(T,U,C) { const uint64_t nt = n/(T*U*4); double t = timeit([&]{ k_rmw_chunked<T,U,C><<<sms*C,T>>>(a,nt,1.0,0.0); }, 10); printf("rmw_chunked T=%d U=%d C=%d : %8.1f GB/s\n",T,U,C, 2*n*8/t/1e9); }
  CHK(256,4,4) CHK(256,4,2) CHK(256,8,2) CHK(256,2,8)
  return 0;
}
