// onika/parallel/parallel_execution_queue.h -- queues of parallel operations and their scheduling onto lanes.
//
// Semantics kept from the reference: operations are enqueued with the queue's current lane, scheduled at flush in FIFO
// order, ordered inside a lane and free to overlap across lanes; a delay queue (ParallelExecutionQueueBase) can be
// spliced into a schedulable queue keeping each operation's lane; wait() synchronises, runs the finalize callback
// (returning the PEC to its pool) and destroys the type-erased functor in place.
// B200 runtime underneath: lane i <-> stream i of libonika_b200 (onk_lane_*); the CUDA bracket of an operation
// (events, return-data copies, counter reset, epilog callback) is onk_op_begin / onk_op_end; host-only functors run as
// host nodes of the lane's stream (onk_lane_host_func) so they stay ordered with the kernels of that lane.
#pragma once
#include <memory>
#include <mutex>
#include <utility>
#include <vector>
#include <onika/cuda/cuda_context.h>
#include <onika/parallel/parallel_execution_context.h>
#include <onika/parallel/parallel_execution_stream.h>
#include <onika/parallel/parallel_data_access.h>
#include <onika/parallel/constants.h>
#include <onika/parallel/block_parallel_for_functor.h>
#include <onika/parallel/parallel_execution_graph.h>
#include <onika/log.h>

#ifndef ONIKA_AUTO_LANE_CYCLE_HINT
#define ONIKA_AUTO_LANE_CYCLE_HINT 16
#endif
#ifndef ONIKA_PARALLEL_QUEUE_MAX_LANES
#define ONIKA_PARALLEL_QUEUE_MAX_LANES 32
#endif

namespace onika { namespace parallel {

  // callback returning the execution stream of a lane
  struct ParallelExecutionStreamPool
  {
    ParallelExecutionStream* (*m_func) (void*,int) = nullptr;
    void *m_priv = nullptr;
    inline ParallelExecutionStream* operator () (int lane) { return (*m_func)(m_priv,lane); }
  };

  // lazily creates one ParallelExecutionStream per lane
  struct ParallelExecutionStreamAutoAllocator
  {
    onika::cuda::CudaContext * m_cuda_ctx = nullptr;
    std::vector< std::shared_ptr< ParallelExecutionStream > > m_pes;
    std::mutex m_mutex;

    inline ParallelExecutionStream * parallel_execution_stream(int lane = DEFAULT_EXECUTION_LANE)
    {
      if( lane == DEFAULT_EXECUTION_LANE ) lane = 0;
      if( lane < 0 || lane >= MAX_EXECUTION_LANES ) fatal_error() << "Invalid execution stream id (" << lane << ")" << std::endl;
      std::lock_guard<std::mutex> lock(m_mutex);
      if( size_t(lane) >= m_pes.size() ) m_pes.resize( lane+1 );
      if( m_pes[lane] == nullptr )
      {
        auto s = std::make_shared<ParallelExecutionStream>();
        s->m_stream_id = lane;
        if( m_cuda_ctx != nullptr ) { s->m_cuda_ctx = m_cuda_ctx; s->m_cu_stream = m_cuda_ctx->getThreadStream(lane); }
        m_pes[lane] = s;
      }
      return m_pes[lane].get();
    }
    static inline ParallelExecutionStream* parallel_execution_stream_cb(void* self, int lane)
    {
      return static_cast<ParallelExecutionStreamAutoAllocator*>(self)->parallel_execution_stream(lane);
    }
    inline ParallelExecutionStreamPool parallel_execution_stream_pool() { return { parallel_execution_stream_cb , this }; }
  };

  struct ParallelExecutionQueueBase
  {
    int m_lane = DEFAULT_EXECUTION_LANE;
    int m_auto_lane_cycle = ONIKA_AUTO_LANE_CYCLE_HINT;
    ParallelExecutionContext* m_queue_list = nullptr;
    ParallelDataAccessVector m_data_access;
    std::mutex m_mutex;

    inline ~ParallelExecutionQueueBase() { assert( m_queue_list == nullptr ); }

    inline void set_lane(int l , int lane_cycle = ONIKA_AUTO_LANE_CYCLE_HINT)
    {
      m_lane = l;
      m_auto_lane_cycle = ( lane_cycle < int(ONIKA_PARALLEL_QUEUE_MAX_LANES) ) ? lane_cycle : int(ONIKA_PARALLEL_QUEUE_MAX_LANES);
    }
    inline void reset_data_access() { const std::lock_guard lk( m_mutex ); m_data_access.clear(); }
    inline void add_data_access(const ParallelDataAccess& pda) { m_data_access.push_back(pda); }

    inline void enqueue(ParallelExecutionContext* pec, bool from_other_queue = false)
    {
      assert( pec->m_next == nullptr && pec->m_stream == nullptr );
      const std::lock_guard lk( m_mutex );
      if( ! from_other_queue )
      {
        pec->m_lane = m_lane;
        std::swap( pec->m_data_access , m_data_access );
      }
      m_data_access.clear();
      m_queue_list = pec_list_append( m_queue_list , pec );
    }

    // Lane assignment of operations queued under any_lane() (their lane is UNDEFINED_EXECUTION_LANE).
    //  * single tasks and operations without declared data access: round robin over the auto-lane cycle, as the reference
    //    does (src/parallel/parallel_execution_queue.cpp:163-183);
    //  * operations WITH declared accesses: the reference detects potential conflicts and prints them but leaves placement
    //    unfinished (.cpp:185-207), i.e. they end up serialised.  Here the declarations drive a lane DAG: an operation
    //    goes to the lane of the latest earlier operation it conflicts with (stream order gives the dependency for free),
    //    gets a cross-lane edge (m_wait_lanes -> onk_lane_wait_lane at schedule time) for every other lane holding a
    //    conflicting predecessor, and takes the least loaded lane of the cycle when it conflicts with nothing.
    //    `in_flight` = operations already scheduled and not yet waited for; they count as predecessors too.
    //  The outcome equals serial FIFO execution whenever the declared accesses are truthful.
    inline void pre_process_queue( ParallelExecutionContext* head , ParallelExecutionContext* in_flight = nullptr )
    {
      const int ncycle = ( m_auto_lane_cycle > 0 ) ? m_auto_lane_cycle : 1;
      int lane_load[ONIKA_PARALLEL_QUEUE_MAX_LANES];
      for(int l=0;l<int(ONIKA_PARALLEL_QUEUE_MAX_LANES);l++) lane_load[l] = 0;
      auto count_load = [&](const ParallelExecutionContext* p) { if( p->m_lane >= 0 && p->m_lane < int(ONIKA_PARALLEL_QUEUE_MAX_LANES) ) ++ lane_load[p->m_lane]; };
      for(auto* p=in_flight; p!=nullptr; p=p->m_next) count_load(p);

      // phase 1 : operations that cannot be reasoned about keep the reference's placement
      int cycle = 0;
      for(auto* pec=head; pec!=nullptr; pec=pec->m_next)
      {
        if( pec->m_lane == DEFAULT_EXECUTION_LANE ) pec->m_lane = 0;
        else if( pec->m_lane == UNDEFINED_EXECUTION_LANE && ( get_block_parallel_functor(pec).is_single_task() || pec->m_data_access.empty() ) )
        {
          pec->m_lane = cycle;
          cycle = ( cycle + 1 ) % ncycle;
        }
        if( pec->m_lane >= 0 ) count_load(pec);
      }

      // phase 2 : operations with declared accesses, in FIFO order
      for(auto* pec=head; pec!=nullptr; pec=pec->m_next)
      {
        if( pec->m_lane != UNDEFINED_EXECUTION_LANE ) continue;
        int last_conflict_lane = -1;
        auto visit = [&](const ParallelExecutionContext* q)
        {
          if( q->m_lane < 0 || q->m_data_access.empty() ) return;
          if( ! data_access_dependency( q->m_data_access , pec->m_data_access ) ) return;
          pec->add_lane_dependency( q->m_lane );
          last_conflict_lane = q->m_lane;
        };
        for(auto* q=in_flight; q!=nullptr; q=q->m_next) visit(q);
        for(auto* q=head; q!=nullptr && q!=pec; q=q->m_next) visit(q);
        if( last_conflict_lane >= 0 ) pec->m_lane = last_conflict_lane;
        else
        {
          int best = 0;
          for(int l=1;l<ncycle;l++) if( lane_load[l] < lane_load[best] ) best = l;
          pec->m_lane = best;
        }
        pec->m_wait_lanes[pec->m_lane>>6] &= ~( uint64_t(1) << (pec->m_lane&63) );   // same lane: ordered by the stream itself
        count_load(pec);
      }
    }
  };

  struct ParallelExecutionQueue : public ParallelExecutionQueueBase
  {
    ParallelExecutionStreamPool m_stream_pool = {};
    ParallelExecutionContext* m_exec_list = nullptr;
    ParallelExecutionGraph* m_capture_target = nullptr;   // set by `queue << capture(g)`: the next flush records into g
    bool m_capturing = false;

    static inline std::shared_ptr<ParallelExecutionQueue> s_default_queue = nullptr;
    static inline std::shared_ptr<ParallelExecutionStreamAutoAllocator> s_default_stream_allocator = nullptr;

    static inline ParallelExecutionQueue& default_queue()
    {
      if( s_default_stream_allocator == nullptr )
      {
        s_default_stream_allocator = std::make_shared<ParallelExecutionStreamAutoAllocator>();
        s_default_stream_allocator->m_cuda_ctx = onika::cuda::CudaContext::default_cuda_ctx();
      }
      if( s_default_queue == nullptr )
      {
        s_default_queue = std::make_shared<ParallelExecutionQueue>();
        s_default_queue->m_stream_pool = s_default_stream_allocator->parallel_execution_stream_pool();
      }
      return * s_default_queue;
    }

    inline ~ParallelExecutionQueue() { schedule_all(); wait(); assert( m_exec_list == nullptr ); }

    inline void schedule_all(int lane = UNDEFINED_EXECUTION_LANE )
    {
      const std::lock_guard lk( m_mutex );
      // operations pushed under any_lane() carry an undefined lane: give them lanes before scheduling
      bool any_undefined = false;
      for(auto* p=m_queue_list; p!=nullptr; p=p->m_next) any_undefined = any_undefined || ( p->m_lane == UNDEFINED_EXECUTION_LANE );
      if( any_undefined ) pre_process_queue( m_queue_list , m_exec_list );
      if( m_capture_target != nullptr ) { capture_all(); return; }
      // cross-lane edges rely on FIFO submission (an edge records what its source lane holds NOW): when the queue carries
      // such edges a lane-filtered flush would submit a dependent operation before its predecessor, so everything goes
      if( lane >= 0 )
        for(auto* p=m_queue_list; p!=nullptr; p=p->m_next)
          if( ( p->m_wait_lanes[0] | p->m_wait_lanes[1] | p->m_wait_lanes[2] | p->m_wait_lanes[3] ) != 0 ) { lane = UNDEFINED_EXECUTION_LANE; break; }
      ParallelExecutionContext *keep = nullptr, *scheduled = nullptr;
      auto* p = m_queue_list;
      while( p != nullptr )
      {
        auto* next = p->m_next;
        p->m_next = nullptr;
        if( lane < 0 || p->m_lane == lane ) { schedule(p); scheduled = pec_list_append( scheduled , p ); }
        else keep = pec_list_append( keep , p );
        p = next;
      }
      m_queue_list = keep;
      m_exec_list = pec_list_append( m_exec_list , scheduled );
    }

    // records everything queued into m_capture_target instead of running it (see parallel_execution_graph.h)
    inline void capture_all()
    {
      auto & g = * m_capture_target;
      m_capture_target = nullptr;
      if( g.valid() || g.m_pec_list != nullptr ) fatal_error() << "capture : ParallelExecutionGraph already holds a capture, release() it first" << std::endl;
      if( m_queue_list == nullptr ) return;
      // everything that allocates (lane streams, device scratch, events) happens before the capture starts
      int lanes[MAX_EXECUTION_LANES]; int nlanes = 0;
      for(auto* p=m_queue_list; p!=nullptr; p=p->m_next)
      {
        if( p->m_lane == UNDEFINED_EXECUTION_LANE || p->m_lane == DEFAULT_EXECUTION_LANE ) p->m_lane = 0;
        bool seen = false;
        for(int i=0;i<nlanes;i++) seen = seen || ( lanes[i] == p->m_lane );
        if( ! seen ) { lanes[nlanes++] = p->m_lane; m_stream_pool( p->m_lane ); }
        if( p->m_execution_target == ParallelExecutionContext::EXECUTION_TARGET_CUDA ) p->init_device_scratch();
      }
      g.m_origin_lane = lanes[0];
      g.m_num_lanes = nlanes;
      ONIKA_B200_CHECK( onk_graph_begin( lanes[0] , lanes+1 , nlanes-1 ) );
      m_capturing = true;
      auto* p = m_queue_list;
      m_queue_list = nullptr;
      while( p != nullptr )
      {
        auto* next = p->m_next;
        p->m_next = nullptr;
        schedule( p );
        g.m_pec_list = pec_list_append( g.m_pec_list , p );
        ++ g.m_num_operations;
        p = next;
      }
      m_capturing = false;
      ONIKA_B200_CHECK( onk_graph_end( & g.m_graph ) );
    }

    static inline void host_operation_trampoline(void* p)
    {
      auto* pec = static_cast<ParallelExecutionContext*>(p);
      get_block_parallel_functor(pec).execute_omp_parallel_region( pec , pec->m_stream );
    }
    static inline void host_operation_replay_trampoline(void* p)   // host node of a captured graph: runs once per replay
    {
      auto* pec = static_cast<ParallelExecutionContext*>(p);
      pec->m_stream->m_omp_execution_count.fetch_add(1);
      get_block_parallel_functor(pec).execute_omp_parallel_region( pec , pec->m_stream );
    }
    static inline void gpu_epilog_trampoline(void* p)
    {
      auto* pec = static_cast<const ParallelExecutionContext*>(p);
      get_block_parallel_functor(pec).execute_gpu_epilog( pec );
    }

    inline void schedule(ParallelExecutionContext* pec)
    {
      assert( pec->m_stream == nullptr && pec->m_next == nullptr );
      int lane = pec->m_lane;
      if( lane == UNDEFINED_EXECUTION_LANE || lane == DEFAULT_EXECUTION_LANE ) lane = 0;
      pec->m_lane = lane;
      auto exec_stream = m_stream_pool( lane );
      std::lock_guard lk_stream( exec_stream->m_mutex );
      pec->m_stream = exec_stream;
      const auto & func = get_block_parallel_functor( pec );

      // lane DAG edges computed by pre_process_queue: operations are submitted in FIFO order, so everything this one
      // depends on is already in its lane's stream -- one event edge per lane
      for(int w=0;w<4;w++) for(uint64_t bits=pec->m_wait_lanes[w]; bits!=0; bits&=bits-1)
      {
        const int other = w*64 + __builtin_ctzll(bits);
        if( other != lane ) ONIKA_B200_CHECK( onk_lane_wait_lane( lane , other ) );
      }

      if( pec->m_execution_target == ParallelExecutionContext::EXECUTION_TARGET_CUDA )
      {
        if( exec_stream->m_cuda_ctx == nullptr || exec_stream->m_cuda_ctx != pec->m_cuda_ctx )
          fatal_error() << "Cannot schedule GPU parallel operation onto stream with no GPU context" << std::endl;
        pec->init_device_scratch();
        pec->m_scheduled_on_gpu = true;
        ONIKA_B200_CHECK( onk_op_begin( pec->m_native_op , lane , pec->m_return_data_input , pec->m_return_data_size ,
                                        ( ( pec->m_reset_counters || pec->m_grid_size.x > 0 ) ? ONK_OP_RESET_COUNTERS : 0 )
                                        | ( ParallelExecutionContext::s_gpu_timing ? 0 : ONK_OP_NO_TIMING ) ) );
        func.stream_gpu_initialize( pec , exec_stream );
        func.stream_gpu_kernel( pec , exec_stream );
        func.stream_gpu_finalize( pec , exec_stream );
        ONIKA_B200_CHECK( onk_op_end( pec->m_native_op , pec->m_return_data_output , pec->m_return_data_size ,
                                      func.needs_gpu_epilog(pec) ? gpu_epilog_trampoline : nullptr , pec ) );
      }
      else
      {
        // user-declared host-only functor: a host node in the lane's stream keeps it ordered with the lane's kernels
        pec->m_scheduled_on_gpu = false;
        if( m_capturing ) ONIKA_B200_CHECK( onk_lane_host_func( lane , host_operation_replay_trampoline , pec ) );
        else
        {
          exec_stream->m_omp_execution_count.fetch_add(1);
          ONIKA_B200_CHECK( onk_lane_host_func( lane , host_operation_trampoline , pec ) );
        }
      }
    }

    inline ParallelExecutionContext* sync_and_remove(ParallelExecutionContext* head, int lane = UNDEFINED_EXECUTION_LANE)
    {
      ParallelExecutionContext *keep = nullptr;
      auto* pec = head;
      // every operation of the list was submitted before this call and a lane runs in order: one synchronisation per
      // lane covers all of the lane's operations
      uint64_t lane_synced[4] = { 0 , 0 , 0 , 0 };
      while( pec != nullptr )
      {
        auto* next = pec->m_next;
        pec->m_next = nullptr;
        if( lane < 0 || pec->m_lane == lane )
        {
          std::lock_guard lk( pec->m_stream->m_mutex );
          const unsigned l = unsigned( pec->m_lane ) & 255u;
          if( ( ( lane_synced[l>>6] >> (l&63) ) & 1 ) == 0 )
          {
            pec->m_stream->wait_nolock();
            lane_synced[l>>6] |= uint64_t(1) << (l&63);
          }
          if( pec->m_scheduled_on_gpu )
          {
            float ms = 0.0f;
            ONIKA_B200_CHECK( onk_op_elapsed_ms( pec->m_native_op , &ms ) );
            pec->m_total_gpu_execution_time = ms;
          }
          pec->m_stream = nullptr;
          auto finalize = pec->m_finalize;
          reinterpret_cast<BlockParallelForHostFunctor*>( pec->m_host_scratch.functor_data ) -> ~BlockParallelForHostFunctor();
          if( finalize.m_func != nullptr ) ( * finalize.m_func ) ( pec , finalize.m_data );
        }
        else keep = pec_list_append( keep , pec );
        pec = next;
      }
      return keep;
    }

    inline void wait_nolock(int lane = UNDEFINED_EXECUTION_LANE) { m_exec_list = sync_and_remove( m_exec_list , lane ); }
    inline void wait(int lane = UNDEFINED_EXECUTION_LANE) { std::lock_guard lk( m_mutex ); wait_nolock(lane); }

    inline bool query_status(int lane = UNDEFINED_EXECUTION_LANE)
    {
      const std::lock_guard lk( m_mutex );
      if( m_exec_list == nullptr && m_queue_list == nullptr ) return true;
      if( m_queue_list != nullptr ) return false;
      for(auto* pec=m_exec_list; pec!=nullptr; pec=pec->m_next)
      {
        if( lane >= 0 && pec->m_lane != lane ) continue;
        int idle = 0;
        ONIKA_B200_CHECK( onk_lane_query( pec->m_lane , &idle ) );
        if( ! idle ) return false;
      }
      wait_nolock( lane );
      return true;
    }

    inline bool empty() { const std::lock_guard lk( m_mutex ); return m_exec_list == nullptr && m_queue_list == nullptr; }
  };

} }
