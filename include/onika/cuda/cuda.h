// onika/cuda/cuda.h -- the ONIKA_CU_* vocabulary user functors are written in (B200-native front end).
//
// Same macro spellings and meanings as the reference layer (include/onika/cuda/cuda.h there), for ONE back end:
// NVIDIA sm_100a.  Inside device code the macros map to CUDA built-ins; in the host compilation pass they have the
// "one block of one thread" meaning the reference gives them (BLOCK_SIZE 1, THREAD_IDX 0, BLOCK_SIMD_FOR = plain
// loop), which is what user-declared host-only functors (CudaCompatible = false) execute with.
// Differences on purpose: ONIKA_CU_WARP_SIZE is 32 on the device (the reference swaps the CUDA/HIP values,
// SURVEY appendix A.1); there is no HIP branch and no "vgpu" shim.
#pragma once

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <type_traits>

#include <onika/macro_utils.h>
#include <onika/atomic.h>
#if defined(_OPENMP)
#include <omp.h>
#endif

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#ifndef ONIKA_CUDA_VERSION
#define ONIKA_CUDA_VERSION __CUDACC_VER_MAJOR__
#endif
#endif

#ifndef ONIKA_CU_MAX_THREADS_PER_BLOCK
#define ONIKA_CU_MAX_THREADS_PER_BLOCK 128
#endif
#ifndef ONIKA_CU_MIN_BLOCKS_PER_SM
#define ONIKA_CU_MIN_BLOCKS_PER_SM 4
#endif
#ifndef ONIKA_CU_ENABLE_KERNEL_BOUNDS
#define ONIKA_CU_ENABLE_KERNEL_BOUNDS 0
#endif

#define ONIKA_ALWAYS_INLINE inline __attribute__((always_inline))

// ------------------------------------------------------------------------------------------------ language keywords
#if defined(__CUDACC__)
#  define ONIKA_HOST_FUNC             __host__
#  define ONIKA_DEVICE_FUNC           __device__
#  define ONIKA_HOST_DEVICE_FUNC      __host__ __device__
#  define ONIKA_DEVICE_KERNEL_FUNC    __global__
#  define ONIKA_STATIC_INLINE_KERNEL  [[maybe_unused]] static
#  define ONIKA_CU_GRID_CONSTANT      __grid_constant__
#  define ONIKA_BUILTIN_ASSUME_ALIGNED(p,a) p
#  if ONIKA_CU_ENABLE_KERNEL_BOUNDS == 1
#    define ONIKA_DEVICE_KERNEL_BOUNDS(T,B) __launch_bounds__(T,B)
#  else
#    define ONIKA_DEVICE_KERNEL_BOUNDS(T,B)
#  endif
   // user kernels launched directly on a lane's stream (STREAM = ParallelExecutionStream::m_cu_stream / getThreadStream(lane))
#  define ONIKA_CU_LAUNCH_KERNEL(GDIM,BDIM,SHMEM,STREAM,FUNC, ... ) FUNC <<< GDIM , BDIM , SHMEM , STREAM >>> ( __VA_ARGS__ )
#else
#  define ONIKA_HOST_FUNC
#  define ONIKA_DEVICE_FUNC
#  define ONIKA_HOST_DEVICE_FUNC
#  define ONIKA_DEVICE_KERNEL_FUNC
#  define ONIKA_STATIC_INLINE_KERNEL  static inline
#  define ONIKA_CU_GRID_CONSTANT
#  define ONIKA_BUILTIN_ASSUME_ALIGNED(p,a) __builtin_assume_aligned(p,a)
#  define ONIKA_DEVICE_KERNEL_BOUNDS(T,B)
   // translation units compiled by the host compiler cannot launch kernels: same diagnostic + abort as the reference
#  define ONIKA_CU_LAUNCH_KERNEL(GDIM,BDIM,SHMEM,STREAM,FUNC, ... ) do { \
     std::printf("Illegal call to Kernel %s<<<...>>>(...) with no GPU support, in %s at %s:%d\n", #FUNC, __FUNCTION__, __FILE__, __LINE__); \
     std::abort(); } while(false)
#endif

namespace onika
{
  // compile-time booleans (complete definitions live in onika/integral_constant.h; forward use only here)
  template<class T, T V> struct IntegralConst;

  namespace cuda
  {
    struct onika_cu_atomic_flag_t { uint32_t m_word = 0; };
    static_assert(sizeof(onika_cu_atomic_flag_t) == sizeof(uint32_t), "atomic flag must be one 32-bit word");
#   define ONIKA_CU_ATOMIC_FLAG_INIT ::onika::cuda::onika_cu_atomic_flag_t{0}
    using onika_cu_memory_order_t = std::memory_order;

    // host implementations of the read-modify-write atomics (what ONIKA_CU_ATOMIC_* mean in host code)
    // (the reference's names live in onika/atomic.h: capture_atomic_add / _sub / _min / _max, inplace_atomic_add / _sub)
    template<class T> inline T host_atomic_add(T& x, T a) { return ::onika::capture_atomic_add( x , a ); }
    template<class T> inline T host_atomic_sub(T& x, T a) { return ::onika::capture_atomic_sub( x , a ); }
    template<class T> inline T host_atomic_min(T& x, T a) { return ::onika::capture_atomic_min( x , a ); }
    template<class T> inline T host_atomic_max(T& x, T a) { return ::onika::capture_atomic_max( x , a ); }
  }
}

// ------------------------------------------------------------------------------------------------ execution vocabulary
#define ONIKA_CU_MEM_ORDER_RELEASE std::memory_order_release
#define ONIKA_CU_MEM_ORDER_RELAXED std::memory_order_relaxed
#define ONIKA_CU_MEM_ORDER_ACQUIRE std::memory_order_acquire
#define ONIKA_CU_MEM_ORDER_SEQ_CST std::memory_order_seq_cst

#if defined(__CUDA_ARCH__)
// ---------------- device pass ----------------
#  define ONIKA_GPU_DEVICE_COMPILE 1
#  define ONIKA_CU_WARP_SIZE           32
#  define ONIKA_DEVICE_CONSTANT_MEMORY __constant__
#  define ONIKA_CU_THREAD_LOCAL        __local__
#  define ONIKA_CU_BLOCK_SHARED        __shared__
#  define ONIKA_CU_GLOBAL_VARIABLE     __managed__

// A CTA is either ONE block of the reference's launch (the usual case) or -- when the runtime packs small items, see
// block_parallel_for_adapter.h "sub-block items" -- blockDim.y independent SUB-BLOCKS of blockDim.x <= 32 threads, each of which
// stands for one block: the block vocabulary below then speaks about the calling thread's sub-block (size blockDim.x, thread
// index threadIdx.x, barrier = the sub-block's lanes of the warp).  Such CTAs are launched as explicit one-CTA clusters
// (onk_launch_ex, ONK_LAUNCH_SUB_BLOCKS), which device code reads from a launch constant: one uniform predicate, no memory.
namespace onika { namespace cuda {
  __device__ __forceinline__ bool sub_block_cta()
  { unsigned int r; asm("{\n\t.reg .pred p;\n\tmov.pred p, %%is_explicit_cluster;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(r)); return r != 0u; }
  __device__ __forceinline__ unsigned int lane_id() { unsigned int r; asm("mov.u32 %0, %%laneid;" : "=r"(r)); return r; }
  // lanes of the calling thread's group of `n` consecutive lanes (n a power of two <= 32)
  __device__ __forceinline__ unsigned int lane_group_mask(unsigned int n)
  { return ( n >= 32u ) ? 0xffffffffu : ( ( ( 1u << n ) - 1u ) << ( lane_id() & ~( n - 1u ) ) ); }
  __device__ __forceinline__ unsigned int block_size()  { return sub_block_cta() ? blockDim.x : ( blockDim.x * blockDim.y * blockDim.z ); }
  __device__ __forceinline__ unsigned int thread_idx()  { return sub_block_cta() ? threadIdx.x : ( ( ( threadIdx.z * blockDim.y ) + threadIdx.y ) * blockDim.x + threadIdx.x ); }
  __device__ __forceinline__ unsigned int block_idx()   { const unsigned int b = ( ( blockIdx.z * gridDim.y ) + blockIdx.y ) * gridDim.x + blockIdx.x; return sub_block_cta() ? ( b * blockDim.y + threadIdx.y ) : b; }
  __device__ __forceinline__ unsigned int grid_size()   { const unsigned int g = gridDim.x * gridDim.y * gridDim.z; return sub_block_cta() ? ( g * blockDim.y ) : g; }
  __device__ __forceinline__ dim3 block_dims()          { return sub_block_cta() ? dim3( blockDim.x , 1 , 1 ) : dim3( blockDim.x , blockDim.y , blockDim.z ); }
  __device__ __forceinline__ dim3 thread_coord()        { return sub_block_cta() ? dim3( threadIdx.x , 0 , 0 ) : dim3( threadIdx.x , threadIdx.y , threadIdx.z ); }
  __device__ __forceinline__ void block_sync()          { if( sub_block_cta() ) __syncwarp( lane_group_mask( blockDim.x ) ); else __syncthreads(); }
  __device__ __forceinline__ int  block_sync_or(int p)  { return sub_block_cta() ? __any_sync( lane_group_mask( blockDim.x ) , p ) : __syncthreads_or( p ); }
  __device__ __forceinline__ int  block_sync_and(int p) { return sub_block_cta() ? __all_sync( lane_group_mask( blockDim.x ) , p ) : __syncthreads_and( p ); }
} }

#  define ONIKA_CU_GRID_DIMS    gridDim
#  define ONIKA_CU_GRID_SIZE    ::onika::cuda::grid_size()
#  define ONIKA_CU_BLOCK_COORD  blockIdx
#  define ONIKA_CU_BLOCK_IDX    ::onika::cuda::block_idx()
#  define ONIKA_CU_BLOCK_DIMS   ::onika::cuda::block_dims()
#  define ONIKA_CU_BLOCK_SIZE   ::onika::cuda::block_size()
#  define ONIKA_CU_THREAD_COORD ::onika::cuda::thread_coord()
#  define ONIKA_CU_THREAD_IDX   ::onika::cuda::thread_idx()

// threads of the block share the iterations [s,e)
#  define ONIKA_CU_BLOCK_SIMD_FOR(T,i,s,e)           for(T i=(s)+ONIKA_CU_THREAD_IDX ; i<(e) ; i+=ONIKA_CU_BLOCK_SIZE )
#  define ONIKA_CU_BLOCK_SIMD_FOR_UNGUARDED(T,i,s,e) for(T _onika_j=(s) , i=(s)+ONIKA_CU_THREAD_IDX ; _onika_j<(e) ; _onika_j+=ONIKA_CU_BLOCK_SIZE, i+=ONIKA_CU_BLOCK_SIZE )

#  define ONIKA_CU_BLOCK_SYNC()        ::onika::cuda::block_sync()
#  define ONIKA_CU_BLOCK_SYNC_OR(p)    ::onika::cuda::block_sync_or(p)
#  define ONIKA_CU_BLOCK_SYNC_AND(p)   ::onika::cuda::block_sync_and(p)
#  define ONIKA_CU_BLOCK_FENCE()       __threadfence_block()
#  define ONIKA_CU_DEVICE_FENCE()      __threadfence()
#  define ONIKA_CU_SYSTEM_FENCE()      __threadfence_system()
#  define ONIKA_CU_WARP_SYNC()         __syncwarp()
#  define ONIKA_CU_WARP_SHFL_DOWN_SYNC(mask,var,delta,width) __shfl_down_sync(mask,var,delta,width)
#  define ONIKA_CU_BLOCK_ACTIVE_MASK() __activemask()
#  define ONIKA_CU_WARP_ACTIVE_THREADS_ALL(pred) __all_sync( __activemask(), int(pred) )
#  define ONIKA_CU_WARP_ACTIVE_THREADS_ANY(pred) __any_sync( __activemask(), int(pred) )

#  define ONIKA_CU_ATOMIC_STORE(x,a,...)       ( * (volatile std::remove_reference_t<decltype(x)> *) &(x) ) = (a)
#  define ONIKA_CU_ATOMIC_LOAD(x,...)          ( * (volatile const std::remove_reference_t<decltype(x)> *) &(x) )
#  define ONIKA_CU_ATOMIC_ADD(x,a,...)         atomicAdd( &(x) , static_cast<std::remove_reference_t<decltype(x)> >(a) )
#  define ONIKA_CU_INPLACE_ATOMIC_ADD(x,a,...) atomicAdd( &(x) , static_cast<std::remove_reference_t<decltype(x)> >(a) )
#  define ONIKA_CU_BLOCK_ATOMIC_ADD(x,a,...)   atomicAdd( &(x) , static_cast<std::remove_reference_t<decltype(x)> >(a) )   /* device scope, as in the reference: callers also use it on global memory */
#  define ONIKA_CU_ATOMIC_SUB(x,a,...)         atomicSub( &(x) , static_cast<std::remove_reference_t<decltype(x)> >(a) )
#  define ONIKA_CU_INPLACE_ATOMIC_SUB(x,a,...) atomicSub( &(x) , static_cast<std::remove_reference_t<decltype(x)> >(a) )
#  define ONIKA_CU_BLOCK_ATOMIC_SUB(x,a,...)   atomicSub( &(x) , static_cast<std::remove_reference_t<decltype(x)> >(a) )
#  define ONIKA_CU_ATOMIC_MIN(x,a,...)         atomicMin( &(x) , static_cast<std::remove_reference_t<decltype(x)> >(a) )
#  define ONIKA_CU_ATOMIC_MAX(x,a,...)         atomicMax( &(x) , static_cast<std::remove_reference_t<decltype(x)> >(a) )
#  define ONIKA_CU_ATOMIC_FLAG_TEST_AND_SET(f) ( ! atomicCAS((uint32_t*)&(f),0u,1u) )
#  define ONIKA_CU_ATOMIC_FLAG_CLEAR(f)        * (volatile uint32_t *) &(f) = 0

#  define ONIKA_CU_NANOSLEEP(ns)       __nanosleep(ns)
#  define ONIKA_CU_CLOCK()             clock64()
#  define ONIKA_CU_CLOCK_ELAPSED(a,b)  ((b)-(a))
#  define ONIKA_CU_VALUE_IF_CUDA(a,b)  (a)
#  define ONIKA_CU_ABORT()             do{ __threadfence(); __trap(); }while(0)
namespace onika { namespace cuda { using onika_cu_clock_t = long long int; } }

#else
// ---------------- host pass: one block of one thread ----------------
#  define ONIKA_CU_WARP_SIZE           1
#  define ONIKA_DEVICE_CONSTANT_MEMORY
#  define ONIKA_CU_THREAD_LOCAL
#  define ONIKA_CU_BLOCK_SHARED
#  define ONIKA_CU_GLOBAL_VARIABLE

// host items run inside an OpenMP region: the "grid" is the team and the "block index" the thread's rank in it, as in the
// reference (cuda.h:221-224 there) -- functors index per-block scratch with it.  Without OpenMP: one block.
#  if defined(_OPENMP)
#    define ONIKA_CU_GRID_SIZE    static_cast<unsigned int>(omp_get_num_threads())
#    define ONIKA_CU_GRID_DIMS    onikaDim3_t{ static_cast<unsigned int>(omp_get_num_threads()) , 1 , 1 }
#    define ONIKA_CU_BLOCK_IDX    static_cast<unsigned int>(omp_get_thread_num())
#    define ONIKA_CU_BLOCK_COORD  onikaDim3_t{ static_cast<unsigned int>(omp_get_thread_num()) , 0 , 0 }
#  else
#    define ONIKA_CU_GRID_SIZE    1u
#    define ONIKA_CU_GRID_DIMS    onikaDim3_t{1,1,1}
#    define ONIKA_CU_BLOCK_IDX    0u
#    define ONIKA_CU_BLOCK_COORD  onikaDim3_t{0,0,0}
#  endif
#  define ONIKA_CU_BLOCK_SIZE   1
#  define ONIKA_CU_BLOCK_DIMS   onikaDim3_t{1,1,1}
#  define ONIKA_CU_THREAD_IDX   0
#  define ONIKA_CU_THREAD_COORD onikaDim3_t{0,0,0}

#  define ONIKA_CU_BLOCK_SIMD_FOR(T,i,s,e)           for(T i=(s) ; i<(e) ; ++i)
#  define ONIKA_CU_BLOCK_SIMD_FOR_UNGUARDED(T,i,s,e) for(T i=(s) ; i<(e) ; ++i)

#  define ONIKA_CU_BLOCK_SYNC()        (void)0
#  define ONIKA_CU_BLOCK_SYNC_OR(p)    (p)
#  define ONIKA_CU_BLOCK_SYNC_AND(p)   (p)
#  define ONIKA_CU_BLOCK_FENCE()       (void)0
#  define ONIKA_CU_DEVICE_FENCE()      std::atomic_thread_fence(std::memory_order_seq_cst)
#  define ONIKA_CU_SYSTEM_FENCE()      std::atomic_thread_fence(std::memory_order_seq_cst)
#  define ONIKA_CU_WARP_SYNC()         (void)0
#  define ONIKA_CU_WARP_SHFL_DOWN_SYNC(mask,var,delta,width) (var)
#  define ONIKA_CU_BLOCK_ACTIVE_MASK() 1u
#  define ONIKA_CU_WARP_ACTIVE_THREADS_ALL(pred) (pred)
#  define ONIKA_CU_WARP_ACTIVE_THREADS_ANY(pred) (pred)

#  define ONIKA_CU_ATOMIC_STORE(x,a,...)       __atomic_store_n( &(x) , (a) , __ATOMIC_SEQ_CST )
#  define ONIKA_CU_ATOMIC_LOAD(x,...)          __atomic_load_n( &(x) , __ATOMIC_SEQ_CST )
#  define ONIKA_CU_ATOMIC_ADD(x,a,...)         ::onika::cuda::host_atomic_add( x , static_cast<std::remove_reference_t<decltype(x)> >(a) )
#  define ONIKA_CU_INPLACE_ATOMIC_ADD(x,a,...) ::onika::cuda::host_atomic_add( x , static_cast<std::remove_reference_t<decltype(x)> >(a) )
#  define ONIKA_CU_BLOCK_ATOMIC_ADD(x,a,...)   (x)+=(a)
#  define ONIKA_CU_ATOMIC_SUB(x,a,...)         ::onika::cuda::host_atomic_sub( x , static_cast<std::remove_reference_t<decltype(x)> >(a) )
#  define ONIKA_CU_INPLACE_ATOMIC_SUB(x,a,...) ::onika::cuda::host_atomic_sub( x , static_cast<std::remove_reference_t<decltype(x)> >(a) )
#  define ONIKA_CU_BLOCK_ATOMIC_SUB(x,a,...)   (x)-=(a)
#  define ONIKA_CU_ATOMIC_MIN(x,a,...)         ::onika::cuda::host_atomic_min( x , static_cast<std::remove_reference_t<decltype(x)> >(a) )
#  define ONIKA_CU_ATOMIC_MAX(x,a,...)         ::onika::cuda::host_atomic_max( x , static_cast<std::remove_reference_t<decltype(x)> >(a) )
#  define ONIKA_CU_ATOMIC_FLAG_TEST_AND_SET(f) ( __atomic_exchange_n( &(f).m_word , 1u , __ATOMIC_ACQUIRE ) == 0u )
#  define ONIKA_CU_ATOMIC_FLAG_CLEAR(f)        __atomic_store_n( &(f).m_word , 0u , __ATOMIC_RELEASE )

#  define ONIKA_CU_NANOSLEEP(ns)       std::this_thread::sleep_for(std::chrono::nanoseconds(ns))
#  define ONIKA_CU_CLOCK()             std::chrono::high_resolution_clock::now()
#  define ONIKA_CU_CLOCK_ELAPSED(a,b)  std::chrono::duration<double,std::micro>((b)-(a)).count()
#  define ONIKA_CU_VALUE_IF_CUDA(a,b)  (b)
#  define ONIKA_CU_ABORT()             std::abort()
namespace onika { namespace cuda { using onika_cu_clock_t = decltype(std::chrono::high_resolution_clock::now()); } }
#endif

// ------------------------------------------------------------------------------------------------ hardware intrinsics
namespace onika { namespace cuda {
#if defined(__CUDA_ARCH__)
  __device__ inline unsigned int get_smid() { unsigned int r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
  __device__ inline long long int onika_double_as_longlong(double x) { return __double_as_longlong(x); }
  __device__ inline double onika_longlong_as_double(long long int x) { return __longlong_as_double(x); }
#else
  inline constexpr unsigned int get_smid() { return 0; }
  ONIKA_HOST_DEVICE_FUNC inline long long int onika_double_as_longlong(double x) { long long int r; std::memcpy(&r, &x, sizeof r); return r; }
  ONIKA_HOST_DEVICE_FUNC inline double onika_longlong_as_double(long long int x) { double r; std::memcpy(&r, &x, sizeof r); return r; }
#endif
} }

// fp64 atomicMin / atomicMax: sm_100a has no native instruction; compare-and-swap on the bit pattern, reading the current
// value first so that no CAS is issued when the stored value already wins (the common case in a reduction)
#if defined(__CUDA_ARCH__)
__device__ inline double atomicMin(double* addr, double v)
{
  unsigned long long int* a = reinterpret_cast<unsigned long long int*>(addr);
  unsigned long long int seen = *reinterpret_cast<volatile unsigned long long int*>(a);
  while (__longlong_as_double((long long)seen) > v)
  {
    const unsigned long long int prev = atomicCAS(a, seen, (unsigned long long int)__double_as_longlong(v));
    if (prev == seen) break;
    seen = prev;
  }
  return __longlong_as_double((long long)seen);
}
__device__ inline double atomicMax(double* addr, double v)
{
  unsigned long long int* a = reinterpret_cast<unsigned long long int*>(addr);
  unsigned long long int seen = *reinterpret_cast<volatile unsigned long long int*>(a);
  while (__longlong_as_double((long long)seen) < v)
  {
    const unsigned long long int prev = atomicCAS(a, seen, (unsigned long long int)__double_as_longlong(v));
    if (prev == seen) break;
    seen = prev;
  }
  return __longlong_as_double((long long)seen);
}
#endif

#include <onika/integral_constant.h>
#if defined(__CUDA_ARCH__)
#  define ONIKA_GPU_DEVICE_EXECUTION_TYPE ::onika::TrueType
#else
#  define ONIKA_GPU_DEVICE_EXECUTION_TYPE ::onika::FalseType
#endif
#if defined(__CUDACC__)
#  define ONIKA_GPU_FRONTEND_COMPILER ::onika::TrueType
#else
#  define ONIKA_GPU_FRONTEND_COMPILER ::onika::FalseType
#endif
#define gpu_device_execution() ONIKA_GPU_DEVICE_EXECUTION_TYPE{}
#define gpu_frontend_compiler() ONIKA_GPU_FRONTEND_COMPILER{}
