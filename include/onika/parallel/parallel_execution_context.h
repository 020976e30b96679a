// onika/parallel/parallel_execution_context.h -- per-operation state (the "PEC") of the B200 front end.
//
// Same role and member names as the reference structure: execution target, lane, block/grid sizing, return-data
// pointers, callbacks, the in-place (type-erased) functor adapter in m_host_scratch.  What differs: the CUDA side
// (start/stop events, the 1 KiB device scratch with 8 counters + 960 B of return data, the H2D/D2H return-data copies)
// is owned by one runtime-library object per PEC (onk_op, include/onika_b200.h) instead of raw events and a
// CudaDeviceStorage.
#pragma once
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <onika/cuda/cuda.h>
#include <onika/cuda/cuda_context.h>
#include <onika/cuda/cuda_error.h>
#include <onika/cuda/device_storage.h>
#include <onika/memory/allocator.h>
#include <onika/parallel/constants.h>
#include <onika/parallel/parallel_execution_space.h>
#include <onika/parallel/parallel_data_access.h>
#include <onika/flat_tuple.h>
#include <onika_b200.h>

namespace onika { namespace parallel {

  struct HostKernelExecutionScratch
  {
    static constexpr size_t SCRATCH_BUFFER_SIZE = 1024;
    static constexpr size_t MAX_FUNCTOR_SIZE = SCRATCH_BUFFER_SIZE - sizeof( unsigned long long );
    static constexpr size_t MAX_FUNCTOR_ALIGNMENT = onika::memory::DEFAULT_ALIGNMENT;
    alignas(MAX_FUNCTOR_ALIGNMENT) char functor_data[MAX_FUNCTOR_SIZE];
    unsigned long long functor_data_size;
    inline size_t available_data_bytes() const { return MAX_FUNCTOR_SIZE - functor_data_size; }
    inline char* available_data_ptr() { return functor_data + functor_data_size; }
    inline void append_functor_data(const void* p, size_t n) { std::memcpy( available_data_ptr() , p , n ); functor_data_size += n; }
    inline char* alloc_functor_data(size_t n) { char* p = available_data_ptr(); functor_data_size += n; return p; }
  };
  static_assert( sizeof(HostKernelExecutionScratch) == HostKernelExecutionScratch::SCRATCH_BUFFER_SIZE );

  struct GPUKernelExecutionScratch
  {
    static constexpr size_t SCRATCH_BUFFER_SIZE = ONK_OP_SCRATCH_BYTES;
    static constexpr size_t MAX_COUNTERS = 8;
    static constexpr size_t MAX_RETURN_SIZE = SCRATCH_BUFFER_SIZE - MAX_COUNTERS * sizeof(unsigned long long);
    static constexpr unsigned int WORKSTEALING_COUNTER = 0;
    unsigned long long int counters[MAX_COUNTERS];
    char return_data[MAX_RETURN_SIZE];
  };
  static_assert( sizeof(GPUKernelExecutionScratch) == GPUKernelExecutionScratch::SCRATCH_BUFFER_SIZE );
  static_assert( GPUKernelExecutionScratch::MAX_RETURN_SIZE == ONK_OP_MAX_RETURN_BYTES );

  struct ParallelExecutionContext;
  struct ParallelExecutionQueue;
  struct ParallelExecutionStream;

  struct ParallelExecutionCallback
  {
    void(*m_func)(void*) = nullptr;
    void *m_data = nullptr;
    inline bool is_null() const { return m_func == nullptr; }
    inline void execute() const { if( m_func != nullptr ) (*m_func)(m_data); }
  };

  struct ParallelExecutionFinalize
  {
    void(*m_func)(ParallelExecutionContext*,void*) = nullptr;
    void *m_data = nullptr;
  };

  // device pointer to the operation's scratch, with the accessors user code applies to the reference's handle
  struct DeviceScratchHandle
  {
    GPUKernelExecutionScratch* m_ptr = nullptr;
    inline GPUKernelExecutionScratch* get() const { return m_ptr; }
    inline GPUKernelExecutionScratch* operator -> () const { return m_ptr; }   // address arithmetic only: never dereferenced on the host
    inline bool operator == (std::nullptr_t) const { return m_ptr == nullptr; }
    inline bool operator != (std::nullptr_t) const { return m_ptr != nullptr; }
  };

  struct ParallelExecutionContext
  {
    enum ExecutionTarget
    {
      EXECUTION_TARGET_OPENMP ,                     // user-declared host-only functor: runs as a host node in its lane
      EXECUTION_TARGET_CUDA ,
      EXECUTION_TARGET_HOST = EXECUTION_TARGET_OPENMP
    };

    onika::cuda::CudaContext* m_cuda_ctx = nullptr;
    ParallelExecutionQueue* m_default_queue = nullptr;
    ParallelExecutionStream* m_stream = nullptr;    // set once scheduled
    ParallelExecutionContext* m_next = nullptr;     // intrusive list link (queues)
    const char* m_tag = nullptr;
    const char* m_sub_tag = nullptr;

    onk_op* m_native_op = nullptr;                  // events + device scratch + return-data copies, owned by the runtime library
    DeviceScratchHandle m_cuda_scratch;
    HostKernelExecutionScratch m_host_scratch;

    ParallelDataAccessVector m_data_access;
    ParallelExecutionCallback m_execution_end_callback = {};
    ParallelExecutionFinalize m_finalize = {};
    const void * m_return_data_input = nullptr;
    void * m_return_data_output = nullptr;

    OMPScheduling m_omp_sched = OMP_SCHED_DYNAMIC;
    ExecutionTarget m_execution_target = EXECUTION_TARGET_OPENMP;
    uint16_t m_return_data_size = 0;
    uint16_t m_block_threads = ONIKA_CU_MAX_THREADS_PER_BLOCK;
    int16_t m_lane = UNDEFINED_EXECUTION_LANE;
    uint16_t m_omp_num_tasks = 0;

    onikaDim3_t m_block_size = { m_block_threads , 1 , 1 };
    onikaDim3_t m_grid_size = { 0, 0, 0 };          // x>0: fixed grid with work stealing

    double m_total_cpu_execution_time = 0.0;
    double m_total_gpu_execution_time = 0.0;
    bool m_reset_counters = false;
    bool m_scheduled_on_gpu = false;
    uint64_t m_wait_lanes[4] = { 0 , 0 , 0 , 0 };   // lanes (bit set) whose queued work this operation must follow (lane DAG edges)

    inline void add_lane_dependency(int l) { if( l >= 0 && l < 256 ) m_wait_lanes[l>>6] |= ( uint64_t(1) << (l&63) ); }
    inline bool depends_on_lane(int l) const { return l >= 0 && l < 256 && ( ( m_wait_lanes[l>>6] >> (l&63) ) & 1 ) != 0; }

    inline void initialize_stream_events()
    {
      if( m_cuda_ctx != nullptr && m_native_op == nullptr ) ONIKA_B200_CHECK( onk_op_create( & m_native_op ) );
    }

    inline void reset()
    {
      m_cuda_ctx = nullptr; m_default_queue = nullptr; m_stream = nullptr; m_next = nullptr;
      m_lane = UNDEFINED_EXECUTION_LANE; m_omp_num_tasks = 0; m_tag = nullptr; m_sub_tag = nullptr;
      m_data_access.clear();
      m_execution_end_callback = ParallelExecutionCallback{};
      m_finalize = ParallelExecutionFinalize{};
      m_return_data_input = nullptr; m_return_data_output = nullptr; m_return_data_size = 0;
      m_execution_target = EXECUTION_TARGET_OPENMP;
      m_block_threads = ONIKA_CU_MAX_THREADS_PER_BLOCK;
      m_block_size = onikaDim3_t{ m_block_threads , 1 , 1 };
      m_grid_size = onikaDim3_t{ 0, 0, 0 };
      m_omp_sched = OMP_SCHED_DYNAMIC;
      m_reset_counters = false; m_scheduled_on_gpu = false;
      for(auto& w : m_wait_lanes) w = 0;
      m_total_cpu_execution_time = 0.0; m_total_gpu_execution_time = 0.0;
      m_host_scratch.functor_data_size = 0;
    }

    inline ~ParallelExecutionContext()
    {
      if( m_native_op != nullptr ) { onk_op_destroy( m_native_op ); m_native_op = nullptr; }
    }

    inline bool has_gpu_context() const { return m_cuda_ctx != nullptr; }
    inline onika::cuda::CudaContext* gpu_context() const { return m_cuda_ctx; }
    inline const char* tag() const { return ( m_tag != nullptr ) ? m_tag : "<unknown>"; }
    inline const char* sub_tag() const { return ( m_sub_tag != nullptr ) ? m_sub_tag : ""; }

    inline void init_device_scratch()
    {
      if( m_cuda_scratch == nullptr )
      {
        if( m_cuda_ctx == nullptr ) { std::fprintf(stderr,"Fatal error: no Cuda context, cannot initialize device scratch mem\n"); std::abort(); }
        initialize_stream_events();
        void* p = nullptr;
        ONIKA_B200_CHECK( onk_op_device_scratch( m_native_op , &p ) );
        m_cuda_scratch.m_ptr = static_cast<GPUKernelExecutionScratch*>( p );
      }
    }

    inline void* get_device_return_data_ptr() { return m_cuda_scratch->return_data; }

    inline void set_return_data_input( const void* ptr, size_t sz )
    {
      if( sz > GPUKernelExecutionScratch::MAX_RETURN_SIZE ) { std::fprintf(stderr,"Fatal error: return data size too large\n"); std::abort(); }
      m_return_data_input = ptr; m_return_data_size = sz;
    }
    inline void set_return_data_output( void* ptr, size_t sz )
    {
      if( sz > GPUKernelExecutionScratch::MAX_RETURN_SIZE ) { std::fprintf(stderr,"Fatal error: return data size too large\n"); std::abort(); }
      if( m_return_data_input != nullptr && m_return_data_size != 0 && m_return_data_size != sz ) { std::fprintf(stderr,"Fatal error: return data size mismatch\n"); std::abort(); }
      m_return_data_output = ptr; m_return_data_size = sz;
    }
    template<class T> inline void set_return_data_input( const T* init_value )
    {
      static_assert( sizeof(T) <= GPUKernelExecutionScratch::MAX_RETURN_SIZE , "return type size too large" );
      set_return_data_input( init_value , sizeof(T) );
    }
    template<class T> inline void set_return_data_output( T* result )
    {
      static_assert( sizeof(T) <= GPUKernelExecutionScratch::MAX_RETURN_SIZE , "return type size too large" );
      set_return_data_output( result , sizeof(T) );
    }

    // ---- global tuning knobs (configuration.onika.* in the reference application) ----
    static inline int s_parallel_task_core_mult = 4;
    static inline int s_parallel_task_core_add = 0;
    static inline int s_gpu_sm_mult = -1;
    static inline int s_gpu_sm_add = -1;
    static inline int s_gpu_block_size = 128;
    // false (or ONIKA_GPU_TIMING=0 in the environment): operations are not bracketed by timing events, m_total_gpu_execution_time stays 0
    static inline bool s_gpu_timing = []{ const char* e = std::getenv("ONIKA_GPU_TIMING"); return !( e != nullptr && e[0] == '0' ); }();
    static inline onikaDim3_t s_gpu_block_dims = { 8 , 8 , 4 };
    static inline int parallel_task_core_mult() { return s_parallel_task_core_mult; }
    static inline int parallel_task_core_add() { return s_parallel_task_core_add; }
    // fixed grids (work stealing) have gpu_sm_mult x SMs + gpu_sm_add blocks.  Unset, the reference falls back to the CPU task
    // multiplier (4): 4 x 128 threads per SM leave a B200 three quarters empty (value-add rows through the generic work-stealing
    // kernel: 2250 GB/s at 4, 3858 at 8, 5086 at 16 and 32), so the default here fills the SM: 2048 / 128 = 16
    static inline int gpu_sm_mult() { return ( s_gpu_sm_mult >= 0 ) ? s_gpu_sm_mult : int( 2048 / ONIKA_CU_MAX_THREADS_PER_BLOCK ); }
    static inline int gpu_sm_add() { return ( s_gpu_sm_add >= 0 ) ? s_gpu_sm_add : parallel_task_core_add(); }
    static inline int gpu_block_size() { return s_gpu_block_size; }
    static inline onikaDim3_t gpu_block_dims() { return s_gpu_block_dims; }
  };

  inline ParallelExecutionContext* pec_list_append(ParallelExecutionContext* l, ParallelExecutionContext* pec)
  {
    if( l == nullptr ) return pec;
    auto * e = l;
    while( e->m_next != nullptr ) e = e->m_next;
    e->m_next = pec;
    return l;
  }

} }
