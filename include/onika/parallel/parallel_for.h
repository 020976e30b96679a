// onika/parallel/parallel_for.h -- parallel_for(N | space, functor, pec [, options]): one functor call per element.
//
// Built on block_parallel_for like in the reference (an adapter functor turns block items into elements), but the
// mapping is B200-native: a block item is a TILE of BLOCK_SIZE x ELEMS_PER_THREAD consecutive elements, thread t taking
// elements t, t+BLOCK_SIZE, ... of the tile (coalesced, ELEMS_PER_THREAD independent accesses in flight per thread), and
// the persistent item loop of bpf_items_kernel makes the whole thing a grid-stride sweep -- instead of the reference's
// ceil(N/128) blocks of one element per thread.  In the host pass a tile is one element (BLOCK_SIZE = 1).
#pragma once
#include <onika/parallel/block_parallel_for.h>

#ifndef ONIKA_PARFOR_OMPSCHED_DEFAULT
#define ONIKA_PARFOR_OMPSCHED_DEFAULT OMP_SCHED_GUIDED
#endif
#ifndef ONIKA_PARFOR_ELEMS_PER_THREAD
#define ONIKA_PARFOR_ELEMS_PER_THREAD 4
#endif

namespace onika { namespace parallel {

  template<class FuncT> struct ParallelForFunctorTraits { static inline constexpr bool CudaCompatible = false; };

  struct ParallelForOptions
  {
    ParallelExecutionCallback user_cb = {};
    void * return_data = nullptr;
    size_t return_data_size = 0;
    bool enable_gpu = true;
    OMPScheduling omp_scheduling = ONIKA_PARFOR_OMPSCHED_DEFAULT;
  };

  template<class FuncT, class PES> struct ParallelForBlockAdapter;

  template<class FuncT, unsigned int ND>
  struct ParallelForBlockAdapter< FuncT, ParallelExecutionSpace<ND,0> >
  {
    const FuncT m_func;
    const onikaInt3_t m_offset;
    const onikaDim3_t m_dims;

    ONIKA_HOST_DEVICE_FUNC inline void operator () ( ssize_t tile ) const
    {
      if constexpr ( ND == 1 )
      {
        constexpr int EPT = ONIKA_CU_VALUE_IF_CUDA( ONIKA_PARFOR_ELEMS_PER_THREAD , 1 );
        const ssize_t base = tile * ( ssize_t(ONIKA_CU_BLOCK_SIZE) * EPT ) + ONIKA_CU_THREAD_IDX;
#       pragma unroll
        for(int e=0; e<EPT; e++)
        {
          const ssize_t i = base + ssize_t(e) * ONIKA_CU_BLOCK_SIZE;
          if( i >= 0 && i < ssize_t(m_dims.x) ) m_func( m_offset.x + i );
        }
      }
    }

    ONIKA_HOST_DEVICE_FUNC inline void operator () ( const onikaInt3_t& block_coord ) const
    {
      if constexpr ( ND > 1 )
      {
        const ssize_t i = block_coord.x * ssize_t(ONIKA_CU_BLOCK_DIMS.x) + ssize_t(ONIKA_CU_THREAD_COORD.x);
        const ssize_t j = block_coord.y * ssize_t(ONIKA_CU_BLOCK_DIMS.y) + ssize_t(ONIKA_CU_THREAD_COORD.y);
        const ssize_t k = block_coord.z * ssize_t(ONIKA_CU_BLOCK_DIMS.z) + ssize_t(ONIKA_CU_THREAD_COORD.z);
        if( i>=0 && i<ssize_t(m_dims.x) && j>=0 && j<ssize_t(m_dims.y) && k>=0 && k<ssize_t(m_dims.z) )
          m_func( onikaInt3_t{ m_offset.x + i , m_offset.y + j , m_offset.z + k } );
      }
    }
  };

  template<class FuncT, class PES> struct BlockParallelForFunctorTraits< ParallelForBlockAdapter<FuncT,PES> >
  {
    static inline constexpr bool CudaCompatible = ParallelForFunctorTraits<FuncT>::CudaCompatible;
  };

  namespace details
  {
    inline BlockParallelForOptions to_block_options(const ParallelForOptions& o, unsigned int elems_per_thread)
    {
      BlockParallelForOptions b = {};
      b.user_cb = o.user_cb; b.return_data = o.return_data; b.return_data_size = o.return_data_size;
      b.enable_gpu = o.enable_gpu; b.omp_scheduling = o.omp_scheduling;
      b.n_div_blocksize = true; b.n_div_elems_per_thread = elems_per_thread;
      return b;
    }
  }

  template< class FuncT >
  static inline ParallelExecutionWrapper
  parallel_for( uint64_t N , const FuncT& func , ParallelExecutionContext * pec , const ParallelForOptions& opts = ParallelForOptions{} )
  {
    using PES = ParallelExecutionSpace<1,0>;
    const onikaInt3_t offset = {0,0,0};
    const onikaDim3_t dims = { static_cast<unsigned int>(N) , 1 , 1 };
    return block_parallel_for( PES{ {0} , {static_cast<ssize_t>(N)} } , ParallelForBlockAdapter<FuncT,PES>{func,offset,dims} , pec
                             , details::to_block_options(opts, ONIKA_PARFOR_ELEMS_PER_THREAD) );
  }

  template<class FuncT, unsigned int ND, unsigned int ElemND>
  static inline ParallelExecutionWrapper
  parallel_for( ParallelExecutionSpace<ND,ElemND> par_space , const FuncT& func , ParallelExecutionContext * pec , const ParallelForOptions& opts = ParallelForOptions{} )
  {
    using PES = ParallelExecutionSpace<ND,ElemND>;
    static_assert( ElemND==0 , "Index list not supported yet" );
    static_assert( ND <= 3 );
    onikaInt3_t offset = { par_space.m_start[0] ,  0 , 0 };
    onikaDim3_t dims = { static_cast<unsigned int>( par_space.m_end[0] - par_space.m_start[0] ) , 1 , 1 };
    if constexpr ( ND >= 2 ) { offset.y = par_space.m_start[1]; dims.y = par_space.m_end[1] - par_space.m_start[1]; }
    if constexpr ( ND >= 3 ) { offset.z = par_space.m_start[2]; dims.z = par_space.m_end[2] - par_space.m_start[2]; }
    for(unsigned int d=0; d<ND; d++) { par_space.m_end[d] -= par_space.m_start[d]; par_space.m_start[d] = 0; }   // 0-based space; the offset travels in the adapter
    return block_parallel_for( par_space , ParallelForBlockAdapter<FuncT,PES>{func,offset,dims} , pec
                             , details::to_block_options(opts, ND == 1 ? ONIKA_PARFOR_ELEMS_PER_THREAD : 1) );
  }

} }
