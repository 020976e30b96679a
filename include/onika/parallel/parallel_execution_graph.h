// onika/parallel/parallel_execution_graph.h -- replayable capture of what a queue flushes (no counterpart in the reference:
// SURVEY 8 f-2, "CUDA-Graph capture of a lane DAG").
//
//   ParallelExecutionGraph g;
//   queue << capture(g) << any_lane() << acc1 << std::move(op1) << acc2 << std::move(op2) ... << flush;   // recorded, not run
//   g.launch(); g.launch(); g.wait();                                                                    // replays
//
// The flush that follows capture(g) assigns lanes as usual (explicit lanes, or the lane DAG from declared accesses),
// then records every operation -- bracket copies, counter reset, kernels, cross-lane edges, host-only functors and
// epilogs as host nodes -- into one CUDA graph through onk_graph_begin / onk_graph_end.  A replay costs one launch
// whatever the number of operations.  The graph keeps the operations' contexts (and thus the functor copies that host
// nodes and epilogs refer to) until it is destroyed; return-data host pointers and the arrays the functors point to are
// borrowed, as for a normal asynchronous operation, for as long as the graph may be launched.
#pragma once
#include <onika/parallel/parallel_execution_context.h>
#include <onika/parallel/block_parallel_for_functor.h>

namespace onika { namespace parallel {

  struct ParallelExecutionGraph
  {
    onk_graph* m_graph = nullptr;
    ParallelExecutionContext* m_pec_list = nullptr;    // operations recorded in the graph, owned until release()
    int m_origin_lane = 0;
    unsigned int m_num_operations = 0;
    unsigned int m_num_lanes = 0;

    ParallelExecutionGraph() = default;
    ParallelExecutionGraph(const ParallelExecutionGraph&) = delete;
    ParallelExecutionGraph& operator = (const ParallelExecutionGraph&) = delete;
    inline ~ParallelExecutionGraph() { release(); }

    inline bool valid() const { return m_graph != nullptr; }
    inline unsigned int number_of_operations() const { return m_num_operations; }
    inline unsigned int number_of_lanes() const { return m_num_lanes; }

    // enqueue one replay in the origin lane (replays are ordered with each other and with that lane's other work)
    inline void launch()
    {
      if( m_graph == nullptr ) fatal_error() << "ParallelExecutionGraph::launch : nothing was captured" << std::endl;
      ONIKA_B200_CHECK( onk_graph_launch( m_graph , m_origin_lane ) );
    }
    inline void wait() { ONIKA_B200_CHECK( onk_lane_sync( m_origin_lane ) ); }

    // waits for pending replays, destroys the graph and hands the recorded operations back to their pools
    inline void release()
    {
      if( m_graph != nullptr )
      {
        onk_lane_sync( m_origin_lane );
        onk_graph_destroy( m_graph );
        m_graph = nullptr;
      }
      while( m_pec_list != nullptr )
      {
        auto* pec = m_pec_list;
        m_pec_list = pec->m_next;
        pec->m_next = nullptr;
        auto finalize = pec->m_finalize;
        reinterpret_cast<BlockParallelForHostFunctor*>( pec->m_host_scratch.functor_data ) -> ~BlockParallelForHostFunctor();
        pec->m_stream = nullptr;
        if( finalize.m_func != nullptr ) ( * finalize.m_func ) ( pec , finalize.m_data );
      }
      m_num_operations = 0; m_num_lanes = 0;
    }
  };

} }
