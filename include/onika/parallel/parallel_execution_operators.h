// onika/parallel/parallel_execution_operators.h -- the stream-style scheduling vocabulary:
//   queue << set_lane(0) << std::move(op1) << set_lane(1) << std::move(op2) << flush << synchronize;
// and the move-only wrapper returned by block_parallel_for, which runs the operation synchronously on its default
// queue when it is destroyed without having been pushed onto a queue.
#pragma once
#include <concepts>
#include <utility>
#include <onika/parallel/parallel_execution_context.h>
#include <onika/parallel/parallel_execution_queue.h>
#include <onika/parallel/parallel_data_access.h>
#include <onika/log.h>

namespace onika { namespace parallel {

  struct ParallelExecutionWrapper
  {
    ParallelExecutionContext* m_pec = nullptr;
    ParallelExecutionWrapper() = default;
    ParallelExecutionWrapper(const ParallelExecutionWrapper&) = delete;
    ParallelExecutionWrapper& operator = (const ParallelExecutionWrapper &) = delete;
    inline ParallelExecutionWrapper(ParallelExecutionContext* pec) : m_pec(pec) {}
    inline ParallelExecutionWrapper(ParallelExecutionWrapper && o) : m_pec(o.m_pec) { o.m_pec = nullptr; }
    inline ParallelExecutionWrapper& operator = (ParallelExecutionWrapper && o) { m_pec = o.m_pec; o.m_pec = nullptr; return *this; }
    inline ~ParallelExecutionWrapper()
    {
      if( m_pec == nullptr ) return;
      auto * q = m_pec->m_default_queue;
      if( q == nullptr ) fatal_error() << "No default queue to schedule parallel operation" << std::endl;
      q->wait();
      q->reset_data_access();
      q->set_lane( DEFAULT_EXECUTION_LANE );
      q->enqueue( m_pec );
      m_pec = nullptr;
      q->schedule_all();
      q->wait();
    }
  };

  struct flush_t {};       static inline constexpr flush_t flush = {};
  struct synchronize_t {}; static inline constexpr synchronize_t synchronize = {};
  struct set_lane_t { int m_lane = UNDEFINED_EXECUTION_LANE; int m_auto_lane_cycle = ONIKA_AUTO_LANE_CYCLE_HINT; };
  static inline constexpr set_lane_t set_lane(int l, int lane_cycle = ONIKA_AUTO_LANE_CYCLE_HINT ) { return { l , lane_cycle }; }
  static inline constexpr set_lane_t any_lane(int lane_cycle = ONIKA_AUTO_LANE_CYCLE_HINT ) { return { UNDEFINED_EXECUTION_LANE , lane_cycle }; }

  template< std::derived_from<ParallelExecutionQueueBase> Q >
  inline Q & operator << ( Q & q , ParallelExecutionWrapper && w )
  {
    auto * pec = w.m_pec; w.m_pec = nullptr;
    q.enqueue( pec );
    q.reset_data_access();
    return q;
  }
  template< std::derived_from<ParallelExecutionQueueBase> Q >
  inline Q & operator << ( Q & q , ParallelExecutionQueueBase && other )
  {
    std::lock_guard lk( other.m_mutex );
    while( other.m_queue_list != nullptr )
    {
      auto pec = other.m_queue_list;
      other.m_queue_list = pec->m_next;
      pec->m_next = nullptr;
      q.enqueue( pec , true );
    }
    return q;
  }
  template< std::derived_from<ParallelExecutionQueueBase> Q >
  inline Q & operator << ( Q & q , const set_lane_t& sl ) { q.set_lane( sl.m_lane , sl.m_auto_lane_cycle ); return q; }
  template< std::derived_from<ParallelExecutionQueueBase> Q >
  inline Q & operator << ( Q & q , const ParallelDataAccess& pda ) { q.add_data_access( pda ); return q; }

  // queue << capture(g) ... << flush : the flush records into g instead of executing (parallel_execution_graph.h)
  struct capture_t { ParallelExecutionGraph* m_graph = nullptr; };
  static inline capture_t capture( ParallelExecutionGraph& g ) { return { &g }; }
  inline ParallelExecutionQueue& operator << ( ParallelExecutionQueue& q , const capture_t& c )
  {
    if( ! q.empty() ) fatal_error() << "capture : the queue must be empty (flush and synchronize it first)" << std::endl;
    q.m_capture_target = c.m_graph;
    return q;
  }

  inline ParallelExecutionQueue& operator << ( ParallelExecutionQueue& q , flush_t ) { q.schedule_all(); q.set_lane(DEFAULT_EXECUTION_LANE); return q; }
  inline ParallelExecutionQueue& operator << ( ParallelExecutionQueue& q , synchronize_t ) { q.wait(); return q; }

} }
