// onika/parallel/parallel_execution_stream.h -- one execution stream per lane: operations queued on it run in order.
// The CUDA stream belongs to the runtime library's lane table (onk_lane_stream); host-executed operations are host
// nodes of the same stream, so "wait" is a single lane synchronisation.
#pragma once
#include <atomic>
#include <mutex>
#include <onika/cuda/cuda_context.h>
#include <onika/parallel/stream_utils.h>
#include <onika/log.h>

namespace onika { namespace parallel {

  struct ParallelExecutionStream
  {
    onika::cuda::CudaContext* m_cuda_ctx = nullptr;
    onikaStream_t m_cu_stream = 0;
    uint32_t m_stream_id = 0;                        // == lane
    std::atomic<uint32_t> m_omp_execution_count = 0; // host nodes in flight on this lane
    std::mutex m_mutex;

    inline void wait() { std::lock_guard lk( m_mutex ); wait_nolock(); }
    inline void wait_nolock()
    {
      ONIKA_B200_CHECK( onk_lane_sync( int(m_stream_id) ) );
      if( m_omp_execution_count.load() > 0 )
        fatal_error() << "Internal error : unterminated host operations (" << m_omp_execution_count.load() << ") remain in lane " << m_stream_id << std::endl;
    }
  };

} }
