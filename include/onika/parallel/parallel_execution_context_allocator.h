// onika/parallel/parallel_execution_context_allocator.h -- mutex-guarded free-list pool of PECs.
// create() hands out a reset PEC wired to a queue and a GPU context; the default finalize callback returns it to
// the pool after the queue has synchronised it.
#pragma once
#include <mutex>
#include <string_view>
#include <unordered_set>
#include <vector>
#include <onika/parallel/parallel_execution_context.h>
#include <onika/parallel/parallel_execution_queue.h>
#include <onika/log.h>

namespace onika { namespace parallel {

  struct ParallelExecutionContextAllocator
  {
    std::vector< ParallelExecutionContext* > m_free_parallel_execution_contexts;
    std::unordered_set< ParallelExecutionContext* > m_allocated_parallel_execution_contexts;
    std::mutex m_mutex;

    inline ~ParallelExecutionContextAllocator()
    {
      if( ! m_allocated_parallel_execution_contexts.empty() )
      {
        auto err = fatal_error();
        err << "m_allocated_parallel_execution_contexts is not empty :";
        for(auto pec : m_allocated_parallel_execution_contexts) err << ' ' << pec->tag() << '/' << pec->sub_tag() << "@" << (void*)pec;
        err << std::endl;
      }
      for(auto pec : m_free_parallel_execution_contexts) delete pec;
    }

    inline ParallelExecutionContext* create(
        std::string_view tag
      , std::string_view sub_tag = {""}
      , ParallelExecutionQueue * default_queue_ptr = nullptr
      , onika::cuda::CudaContext* default_cuda_ctx = cuda::CudaContext::default_cuda_ctx()
      , int omp_num_tasks = -1
      , ParallelExecutionFinalize && on_terminate_func = {nullptr,nullptr} )
    {
      if( default_queue_ptr == nullptr ) default_queue_ptr = & ParallelExecutionQueue::default_queue();
      std::lock_guard<std::mutex> lock(m_mutex);
      if( m_free_parallel_execution_contexts.empty() ) m_free_parallel_execution_contexts.push_back( new ParallelExecutionContext() );
      auto pec = m_free_parallel_execution_contexts.back();
      m_free_parallel_execution_contexts.pop_back();
      m_allocated_parallel_execution_contexts.insert( pec );
      pec->reset();
      pec->m_tag = tag.data();
      pec->m_sub_tag = sub_tag.data();
      pec->m_cuda_ctx = default_cuda_ctx;
      pec->m_default_queue = default_queue_ptr;
      pec->m_omp_num_tasks = ( omp_num_tasks > 0 ) ? omp_num_tasks : 0;   // task mode is a host-runtime notion: kept as a hint only
      pec->initialize_stream_events();
      pec->m_finalize = ( on_terminate_func.m_func == nullptr ) ? ParallelExecutionFinalize{ free_cb , this } : on_terminate_func;
      return pec;
    }

    inline void free(ParallelExecutionContext* pec)
    {
      pec->reset();
      std::lock_guard<std::mutex> lock(m_mutex);
      auto it = m_allocated_parallel_execution_contexts.find( pec );
      assert( it != m_allocated_parallel_execution_contexts.end() );
      m_free_parallel_execution_contexts.push_back( pec );
      m_allocated_parallel_execution_contexts.erase( it );
    }
    static inline void free_cb(ParallelExecutionContext* pec, void * self) { static_cast<ParallelExecutionContextAllocator*>(self)->free(pec); }
  };

} }
