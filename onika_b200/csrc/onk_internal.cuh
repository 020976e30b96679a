// onk_internal.cuh -- shared internals of libonika_b200.so (not part of the ABI).
#pragma once

#include "../../include/onika_b200.h"

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

namespace onk
{
  // ---- error plumbing -------------------------------------------------------------------------------------
  void set_error(const char* fmt, ...);
  int  fail_cuda(cudaError_t e, const char* what, const char* file, int line);

#define ONK_CUDA(call)                                                              \
  do { cudaError_t _e = (call);                                                     \
       if (_e != cudaSuccess) return ::onk::fail_cuda(_e, #call, __FILE__, __LINE__); } while (0)

#define ONK_REQUIRE(cond, ...)                                                      \
  do { if (!(cond)) { ::onk::set_error(__VA_ARGS__); return ONK_ERR_ARG; } } while (0)

#define ONK_CHECK_LAUNCH() ONK_CUDA(cudaGetLastError())

  // ---- per-lane state -------------------------------------------------------------------------------------
  // Reduction workspace: one partial per CTA + a self-resetting ticket, private to the lane so reductions queued on
  // different lanes never share scratch (the reference's per-PEC GPUKernelExecutionScratch plays this role).
  // Concurrency contract: the workspace is only ever touched by KERNELS, and kernels of one lane execute in stream order,
  // so any number of host threads may queue reductions on the same lane (the launches serialise on the stream; each kernel
  // leaves the ticket at 0).  What is NOT allowed is running two kernels that use the same lane's workspace concurrently:
  // (a) rebinding a lane to another stream (onk_lane_bind_stream) while work is pending on it, and (b) replaying a captured
  // graph (onk_graph_launch) on lane X while reductions are queued directly on one of the lanes the graph was CAPTURED on --
  // the graph's kernels keep the workspace pointers of their capture lanes.  onk_graph_launch documents (b) in the header.
  struct LaneWorkspace
  {
    double*   partials = nullptr;   // MAX_PARTIALS doubles
    unsigned* ticket   = nullptr;   // 1 counter, left at 0 by every kernel that uses it
    double*   result   = nullptr;   // 8 doubles of device scratch for results
  };

  struct Lane
  {
    cudaStream_t  stream = nullptr;
    bool          created = false;   // stream exists
    bool          owned = false;     // we created it (destroy at finalize)
    cudaEvent_t   edge = nullptr;    // re-used by onk_lane_wait_lane: 'what this lane holds now' for another lane to wait on
    LaneWorkspace ws;
  };

  constexpr int      MAX_PARTIALS = 8192;
  constexpr unsigned THREADS = 256;

  struct Runtime
  {
    std::mutex mutex;
    bool       initialised = false;
    int        device = -1;
    int        sm_count = 0;
    cudaDeviceProp prop{};
    std::vector<Lane> lanes;                       // indexed by lane id
    std::unordered_map<const void*, std::pair<size_t,int>> allocs;   // ptr -> (bytes, kind)
    std::atomic<uint64_t> launches{0};
    // device staging for the host-buffer entry points (lazily allocated, onk_host.cu)
    void*  stage_dev[4]  = {nullptr, nullptr, nullptr, nullptr};
    size_t stage_bytes   = 0;
  };

  Runtime& rt();
  int  require_init();                              // ONK_OK or ONK_ERR_STATE / ONK_ERR_NO_DEVICE
  int  lane_get(int lane, Lane** out);              // resolves ONK_LANE_DEFAULT, creates stream + workspace lazily
  inline void count_launch(uint64_t n = 1) { rt().launches.fetch_add(n, std::memory_order_relaxed); }

  // grid sizing: persistent-style grids, a multiple of the SM count
  inline unsigned grid_for(uint64_t work_items, uint64_t items_per_cta, unsigned ctas_per_sm)
  {
    const uint64_t need = (work_items + items_per_cta - 1) / items_per_cta;
    const uint64_t cap  = (uint64_t)rt().sm_count * ctas_per_sm;
    uint64_t g = need < cap ? need : cap;
    if (g < 1) g = 1;
    return (unsigned)g;
  }

  // ---- device helpers -------------------------------------------------------------------------------------
#ifdef __CUDACC__
  struct alignas(32) d4 { double a, b, c, d; };

  // 256-bit streaming load, read-only data (never written by the running kernel)
  __device__ __forceinline__ d4 ld_stream_ro(const double* p)
  {
    d4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v.a), "=d"(v.b), "=d"(v.c), "=d"(v.d) : "l"(p));
    return v;
  }
  // 256-bit streaming load of data the same kernel writes back (no .nc)
  __device__ __forceinline__ d4 ld_stream_rw(const double* p)
  {
    d4 v;
    asm volatile("ld.global.L1::no_allocate.L2::evict_first.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v.a), "=d"(v.b), "=d"(v.c), "=d"(v.d) : "l"(p) : "memory");
    return v;
  }
  __device__ __forceinline__ void st_stream(double* p, const d4& v)
  {
    asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "d"(v.a), "d"(v.b), "d"(v.c), "d"(v.d) : "memory");
  }
  __device__ __forceinline__ double ld_stream_ro1(const double* p)
  {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
  }

  // reduction operators with the reference's comparison semantics (NaN never wins, ties keep the accumulator)
  template<int OP> struct Red;
  template<> struct Red<ONK_OP_MIN>
  {
    __host__ __device__ static constexpr double identity() { return 1.7976931348623157e308; }
    __device__ __forceinline__ static double apply(double acc, double v) { return (v < acc) ? v : acc; }
  };
  template<> struct Red<ONK_OP_MAX>
  {
    __host__ __device__ static constexpr double identity() { return -1.7976931348623157e308; }
    __device__ __forceinline__ static double apply(double acc, double v) { return (v > acc) ? v : acc; }
  };
  template<> struct Red<ONK_OP_SUM>
  {
    __host__ __device__ static constexpr double identity() { return 0.0; }
    __device__ __forceinline__ static double apply(double acc, double v) { return __dadd_rn(acc, v); }
  };

  // butterfly warp reduction then warp partials through shared memory; result valid in thread 0.
  // Fixed combination order => reproducible sums.
  template<int OP, unsigned NTHREADS>
  __device__ __forceinline__ double block_reduce(double v)
  {
    __shared__ double warp_part[NTHREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = Red<OP>::apply(v, __shfl_xor_sync(0xffffffffu, v, o));
    const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    if (lane == 0) warp_part[w] = v;
    __syncthreads();
    double r = Red<OP>::identity();
    if (w == 0)
    {
      r = (lane < NTHREADS / 32) ? warp_part[lane] : Red<OP>::identity();
#pragma unroll
      for (int o = (NTHREADS / 64); o > 0; o >>= 1) r = Red<OP>::apply(r, __shfl_xor_sync(0xffffffffu, r, o));
    }
    __syncthreads();
    return r;
  }
#endif
}
