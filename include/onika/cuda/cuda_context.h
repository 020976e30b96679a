// onika/cuda/cuda_context.h -- runtime-API macro table, handle types and CudaContext for the B200 front end.
// Device selection, lanes->streams and memory go through the C ABI (include/onika_b200.h); the macros below keep the
// reference's spellings for user code that calls the runtime directly (tests/gpu_reduce_min_fp64.cu style).
#pragma once

#include <onika/cuda/cuda.h>
#include <onika/cuda/cuda_error.h>
#include <onika_b200.h>

#include <functional>
#include <iostream>
#include <list>
#include <memory>
#include <string>
#include <vector>
#include <sys/types.h>

// Device description shared by every translation unit of a plugin, whichever compiler built it: ONE layout in nvcc and in
// host-compiler TUs (the inline default_cuda_ctx() below is the same object for both, so a type that changes with the
// compiler would be an ODR violation).  Field names follow cudaDeviceProp for the members the reference reads
// (src/scg-builtin-components/init_cuda.cpp:179-195, include/onika/parallel/block_parallel_for.h:124).
struct onikaDeviceProp_t
{
  char   name[256] = {0};
  int    multiProcessorCount = 0;
  int    warpSize = 32;
  size_t totalGlobalMem = 0;
  size_t sharedMemPerBlock = 0;
  size_t sharedMemPerMultiprocessor = 0;
  int    maxThreadsPerBlock = 1024;
  int    maxThreadsPerMultiProcessor = 2048;
  int    major = 0, minor = 0;
  int    clockRate = 0;                 // kHz
  int    managedMemory = 1;
  int    concurrentManagedAccess = 1;
  int    persistingL2CacheMaxSize = 0;
  int    l2CacheSize = 0;
};

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#include <cuda_runtime_api.h>
// fills the layout-stable description from the runtime's own (what ONIKA_CU_GET_DEVICE_PROPERTIES means here)
static inline cudaError_t onika_get_device_properties( onikaDeviceProp_t* p , int id )
{
  cudaDeviceProp q;
  const cudaError_t e = cudaGetDeviceProperties( &q , id );
  if( e != cudaSuccess ) return e;
  std::snprintf( p->name , sizeof(p->name) , "%s" , q.name );
  p->multiProcessorCount = q.multiProcessorCount; p->warpSize = q.warpSize; p->totalGlobalMem = q.totalGlobalMem;
  p->sharedMemPerBlock = q.sharedMemPerBlock; p->sharedMemPerMultiprocessor = q.sharedMemPerMultiprocessor;
  p->maxThreadsPerBlock = q.maxThreadsPerBlock; p->maxThreadsPerMultiProcessor = q.maxThreadsPerMultiProcessor;
  p->major = q.major; p->minor = q.minor; p->managedMemory = q.managedMemory; p->concurrentManagedAccess = q.concurrentManagedAccess;
  p->persistingL2CacheMaxSize = q.persistingL2CacheMaxSize; p->l2CacheSize = q.l2CacheSize;
  int khz = 0; if( cudaDeviceGetAttribute( &khz , cudaDevAttrClockRate , id ) == cudaSuccess ) p->clockRate = khz;
  return cudaSuccess;
}
using onikaStream_t     = cudaStream_t;
using onikaEvent_t      = cudaEvent_t;
using onikaError_t      = cudaError_t;
using onikaLimit_t      = cudaLimit;
using onikaDim3_t       = dim3;
static inline constexpr auto onikaSuccess            = cudaSuccess;
static inline constexpr auto onikaErrorNotReady      = cudaErrorNotReady;
static inline constexpr auto onikaMemcpyDeviceToHost = cudaMemcpyDeviceToHost;
static inline constexpr auto onikaMemcpyHostToDevice = cudaMemcpyHostToDevice;
static inline constexpr auto onikaSharedMemBankSizeFourByte  = cudaSharedMemBankSizeFourByte;
static inline constexpr auto onikaSharedMemBankSizeEightByte = cudaSharedMemBankSizeEightByte;
static inline constexpr auto onikaSharedMemBankSizeDefault   = cudaSharedMemBankSizeDefault;
static inline constexpr auto onikaLimitStackSize             = cudaLimitStackSize;
static inline constexpr auto onikaLimitPrintfFifoSize        = cudaLimitPrintfFifoSize;
static inline constexpr auto onikaLimitMallocHeapSize        = cudaLimitMallocHeapSize;
static inline constexpr auto onikaDevAttrClockRate           = cudaDevAttrClockRate;
#define ONIKA_CU_NAME_STR                              "cuda-sm100a"
// NVTX ranges (header-only NVTX3: a no-op unless a profiler injects its library), as the reference does around operators
#if __has_include(<nvtx3/nvToolsExt.h>) && ! defined(ONIKA_B200_NO_NVTX)
#  include <nvtx3/nvToolsExt.h>
#  define ONIKA_CU_PROF_RANGE_PUSH(s)                  nvtxRangePush(s)
#  define ONIKA_CU_PROF_RANGE_POP()                    nvtxRangePop()
#else
#  define ONIKA_CU_PROF_RANGE_PUSH(s)                  ((void)(s))
#  define ONIKA_CU_PROF_RANGE_POP()                    ((void)0)
#endif
#define ONIKA_CU_PROF_START()                          ((void)0)
#define ONIKA_CU_PROF_STOP()                           ((void)0)
#define ONIKA_CU_MEM_PREFETCH(ptr,sz,d,st)             cudaMemPrefetchAsync((const void*)(ptr),sz,d,st)
#define ONIKA_CU_CREATE_STREAM_NON_BLOCKING(streamref) cudaStreamCreateWithFlags( & streamref, cudaStreamNonBlocking )
#define ONIKA_CU_STREAM_ADD_CALLBACK(stream,cb,udata)  cudaStreamAddCallback(stream,cb,udata,0u)
#define ONIKA_CU_STREAM_SYNCHRONIZE(STREAM)            cudaStreamSynchronize(STREAM)
#define ONIKA_CU_STREAM_QUERY(STREAM)                  cudaStreamQuery(STREAM)
#define ONIKA_CU_DESTROY_STREAM(streamref)             cudaStreamDestroy(streamref)
#define ONIKA_CU_EVENT_QUERY(evt)                      cudaEventQuery(evt)
#define ONIKA_CU_MALLOC(devPtrPtr,N)                   cudaMalloc(devPtrPtr,N)
#define ONIKA_CU_MALLOC_MANAGED(devPtrPtr,N)           cudaMallocManaged(devPtrPtr,N)
#define ONIKA_CU_FREE(devPtr)                          cudaFree(devPtr)
#define ONIKA_CU_CREATE_EVENT(EVT)                     cudaEventCreate(&EVT)
#define ONIKA_CU_DESTROY_EVENT(EVT)                    cudaEventDestroy(EVT)
#define ONIKA_CU_STREAM_EVENT(EVT,STREAM)              cudaEventRecord(EVT,STREAM)
#define ONIKA_CU_STREAM_HOST_FUNC(s,f,p)               cudaLaunchHostFunc(s,f,p)
#define ONIKA_CU_EVENT_ELAPSED(T,EVT1,EVT2)            cudaEventElapsedTime(&T,EVT1,EVT2)
#define ONIKA_CU_MEMSET(p,v,n,...)                     cudaMemsetAsync(p,v,n OPT_COMMA_VA_ARGS(__VA_ARGS__) )
#define ONIKA_CU_MEMCPY(d,s,n,...)                     cudaMemcpyAsync(d,s,n,cudaMemcpyDefault OPT_COMMA_VA_ARGS(__VA_ARGS__) )
#define ONIKA_CU_MEMCPY_KIND(d,s,n,k,...)              cudaMemcpyAsync(d,s,n,k OPT_COMMA_VA_ARGS(__VA_ARGS__) )
#define ONIKA_CU_MEMCPY_TO_SYMBOL(d,s,n,...)           cudaMemcpyToSymbol(d,s,n,0,cudaMemcpyHostToDevice)
#define ONIKA_CU_GET_DEVICE_COUNT(iPtr)                cudaGetDeviceCount(iPtr)
#define ONIKA_CU_SET_DEVICE(id)                        cudaSetDevice(id)
#define ONIKA_CU_SET_SHARED_MEM_CONFIG(shmc)           cudaDeviceSetSharedMemConfig(shmc)
#define ONIKA_CU_SET_LIMIT(l,v)                        cudaDeviceSetLimit(l,v)
#define ONIKA_CU_GET_LIMIT(vptr,l)                     cudaDeviceGetLimit(vptr,l)
#define ONIKA_CU_GET_DEVICE_PROPERTIES(propPtr,id)     onika_get_device_properties(propPtr,id)
#define ONIKA_CU_GET_DEVICE_ATTRIBUTE(valPtr,attr,id)  cudaDeviceGetAttribute(valPtr,attr,id)
#define ONIKA_CU_DEVICE_SYNCHRONIZE()                  cudaDeviceSynchronize()
#define ONIKA_CU_GET_ERROR_STRING(c)                   cudaGetErrorString(c)
#else
// host-only translation units still see the handle types (opaque) so that headers parse; they cannot launch kernels
struct onikaDim3_t { unsigned int x = 1, y = 1, z = 1; };      // same layout as dim3 (static_assert in the nvcc branch of CudaDevice below)
using onikaStream_t = void*;
using onikaEvent_t  = void*;
using onikaError_t  = int;
static inline constexpr int onikaSuccess = 0;
#define ONIKA_CU_NAME_STR "cuda-sm100a(host TU)"
#define ONIKA_CU_PROF_RANGE_PUSH(s)                    ((void)(s))
#define ONIKA_CU_PROF_RANGE_POP()                      ((void)0)
#endif

struct onikaInt3_t
{
  ssize_t x = 0;
  ssize_t y = 0;
  ssize_t z = 0;
  ONIKA_HOST_DEVICE_FUNC inline constexpr onikaInt3_t operator + (const onikaDim3_t& r) const { return { x + ssize_t(r.x) , y + ssize_t(r.y) , z + ssize_t(r.z) }; }
};

namespace onika { namespace cuda {

  inline constexpr long onika_dim3_size(const onikaDim3_t& d) { return long(d.x) * d.y * d.z; }
  inline constexpr long onika_dim3_size(long d) { return d; }

# if defined(__CUDACC__)
  static_assert( sizeof(onikaDim3_t) == 3 * sizeof(unsigned int) && alignof(onikaDim3_t) == alignof(unsigned int) , "dim3 layout assumed by host-compiler translation units" );
# endif
  struct CudaDevice
  {
    onikaDeviceProp_t m_deviceProp;
    std::list< std::function<void()> > m_finalize_destructors;
    int device_id = 0;
    int m_clock_rate = 0;
  };

  // One process drives one device: the context describes that device; streams belong to lanes owned by
  // the runtime library (onk_lane_stream).
  struct CudaContext
  {
    std::vector<CudaDevice> m_devices;
    std::vector<onikaStream_t> m_threadStream;          // streams handed out so far, by lane (the reference's lazily filled table)

    static inline bool s_global_gpu_enable = true;
    static inline std::shared_ptr<CudaContext> s_default_cuda_ctx = nullptr;

    inline bool has_devices() const { return ! m_devices.empty(); }
    inline unsigned int device_count() const { return m_devices.size(); }
    inline onikaStream_t getThreadStream(unsigned int lane)
    {
      void* s = nullptr;
      ONIKA_B200_CHECK( onk_lane_stream( int(lane) , &s ) );
      if( lane >= m_threadStream.size() ) m_threadStream.resize( lane + 1 , onikaStream_t{} );
      m_threadStream[lane] = (onikaStream_t) s;
      return (onikaStream_t) s;
    }

    static inline void set_global_gpu_enable(bool yn) { s_global_gpu_enable = yn; }
    static inline bool global_gpu_enable() { return s_global_gpu_enable; }

    // device `dev` of this process (default 0, or ONIKA_B200_DEVICE / LOCAL_RANK when set by a launcher)
    static inline CudaContext* default_cuda_ctx()
    {
      if( ! global_gpu_enable() ) return nullptr;
      if( s_default_cuda_ctx == nullptr )
      {
        int dev = 0;
        if( const char* e = std::getenv("ONIKA_B200_DEVICE") ) dev = std::atoi(e);
        else if( const char* e = std::getenv("LOCAL_RANK") ) dev = std::atoi(e);
        ONIKA_B200_CHECK( onk_init(dev) );
        auto ctx = std::make_shared<CudaContext>();
        ctx->m_devices.resize(1);
        auto & d = ctx->m_devices[0];
        int sms = 0; size_t mem = 0;
        ONIKA_B200_CHECK( onk_device_info( & d.device_id , & sms , & mem , d.m_deviceProp.name , sizeof(d.m_deviceProp.name) ) );
        d.m_deviceProp.multiProcessorCount = sms; d.m_deviceProp.totalGlobalMem = mem;       // from the runtime library: any compiler
#       if defined(__CUDACC__)
        ONIKA_CU_CHECK_ERRORS( onika_get_device_properties( & d.m_deviceProp , d.device_id ) );
#       endif
        s_default_cuda_ctx = ctx;
      }
      return s_default_cuda_ctx.get();
    }

    template<class StreamT> inline StreamT& to_stream(StreamT& out)
    {
      out << "=========== " << ONIKA_CU_NAME_STR << " ================" << std::endl;
      if( has_devices() )
      {
        const auto & p = m_devices.front().m_deviceProp;
        out << "GPUs : " << m_devices.size() << std::endl << "Type : " << p.name << std::endl
            << "SMs  : " << p.multiProcessorCount << "x" << p.warpSize << " threads" << std::endl
            << "Mem  : " << ( p.totalGlobalMem >> 20 ) << " MiB" << std::endl;
      }
      else out << "No GPU found" << std::endl;
      out << "=================================" << std::endl;
      return out;
    }
  };

} }
