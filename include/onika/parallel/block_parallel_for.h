// onika/parallel/block_parallel_for.h -- block_parallel_for(space | N, functor, pec [, options]).
// Fills the ParallelExecutionContext (target, block/grid sizing, return data), placement-constructs the type-erased
// adapter inside it and returns the move-only wrapper; nothing runs until the wrapper is queued or destroyed.
#pragma once
#include <algorithm>
#include <onika/cuda/cuda.h>
#include <onika/cuda/cuda_error.h>
#include <onika/cuda/device_storage.h>
#include <onika/soatl/field_id.h>
#include <onika/parallel/parallel_execution_context.h>
#include <onika/parallel/parallel_execution_stream.h>
#include <onika/parallel/block_parallel_for_adapter.h>
#include <onika/parallel/parallel_execution_operators.h>
#include <onika/lambda_tools.h>
#include <onika/stream_utils.h>

#ifndef ONIKA_BLKPARFOR_OMPSCHED_DEFAULT
#define ONIKA_BLKPARFOR_OMPSCHED_DEFAULT OMP_SCHED_GUIDED
#endif

namespace onika { namespace parallel {

  struct BlockParallelForOptions
  {
    ParallelExecutionCallback user_cb = {};
    void * return_data = nullptr;
    size_t return_data_size = 0;
    unsigned int max_block_size = ONIKA_CU_MAX_THREADS_PER_BLOCK;
    bool enable_gpu = true;
    bool fixed_gpu_grid_size = false;
    bool n_div_blocksize = false;            // divide the (0-based) range by the block size, rounding up (device target only)
    OMPScheduling omp_scheduling = ONIKA_BLKPARFOR_OMPSCHED_DEFAULT;
    unsigned int n_div_elems_per_thread = 1; // with n_div_blocksize: each thread covers this many elements of dimension 0
  };

  namespace details
  {
    // Everything that depends on the run-time situation only (not on the functor type): does this operation go to the
    // device, with which block shape / grid policy, and how does its return data travel.  `extent` is the (mutable) end
    // of the 0-based range per dimension, rescaled here when the caller asked for one item per block of elements.
    template<unsigned int ND>
    static inline bool configure_device_target( ParallelExecutionContext* pec , const BlockParallelForOptions& opts , bool indexed
                                              , const ssize_t* range_start , ssize_t* range_end )
    {
      const bool has_device = opts.enable_gpu && pec->m_cuda_ctx != nullptr && pec->m_cuda_ctx->has_devices();
      if( ! has_device ) return false;

      pec->m_execution_target = ParallelExecutionContext::EXECUTION_TARGET_CUDA;

      // block shape: 1D -> min(option, compile-time cap, configured size) threads; 2D/3D -> configured tile (z=1 in 2D)
      const size_t cap = std::min( size_t(ONIKA_CU_MAX_THREADS_PER_BLOCK) , size_t(ParallelExecutionContext::gpu_block_size()) );
      pec->m_block_threads = uint16_t( std::min( size_t(opts.max_block_size) , cap ) );
      pec->m_block_size = ( ND == 1 ) ? onikaDim3_t{ pec->m_block_threads , 1 , 1 } : ParallelExecutionContext::gpu_block_dims();
      if( ND == 2 ) pec->m_block_size.z = 1;

      if( opts.n_div_blocksize )
      {
        if( indexed ) fatal_error() << "n_div_blocksize option is only valid non indexed elements " << std::endl;
        const unsigned int per_item[3] = { pec->m_block_size.x * std::max( 1u , opts.n_div_elems_per_thread ) , pec->m_block_size.y , pec->m_block_size.z };
        // items are tiles of the configured 2D/3D block shape: the block must keep that shape (no sub-block packing)
        if( ND > 1 ) pec->m_block_threads = uint16_t( pec->m_block_size.x * pec->m_block_size.y * pec->m_block_size.z );
        for(unsigned int d=0; d<ND; d++)
        {
          if( range_start[d] != 0 )
            fatal_error() << "n_div_blocksize option is only valid for parallel ranges starting from 0. range start for coord #"<<d<<" is "<< range_start[d] << std::endl;
          range_end[d] = ( range_end[d] + per_item[d] - 1 ) / per_item[d];
        }
      }

      // grid policy: 0 = persistent item loop sized by occupancy; >0 = fixed grid with work stealing
      pec->m_grid_size = onikaDim3_t{ 0 , 0 , 0 };
      if( opts.fixed_gpu_grid_size )
      {
        const auto & prop = pec->m_cuda_ctx->m_devices[0].m_deviceProp;
        pec->m_grid_size.x = prop.multiProcessorCount * ParallelExecutionContext::gpu_sm_mult() + ParallelExecutionContext::gpu_sm_add();
      }
      pec->m_reset_counters = opts.fixed_gpu_grid_size;

      // return data: the same host location seeds the device accumulator and receives the result
      const bool has_return = ( opts.return_data != nullptr && opts.return_data_size > 0 );
      pec->set_return_data_input( has_return ? opts.return_data : nullptr , has_return ? opts.return_data_size : 0 );
      pec->set_return_data_output( has_return ? opts.return_data : nullptr , has_return ? opts.return_data_size : 0 );
      return true;
    }
  }

  template< class FuncT , unsigned int ND, unsigned int ElemND, class ElementListT>
  static inline ParallelExecutionWrapper
  block_parallel_for( ParallelExecutionSpace<ND,ElemND,ElementListT> par_space , const FuncT& func , ParallelExecutionContext * pec
                    , const BlockParallelForOptions& opts = BlockParallelForOptions{} )
  {
    // the functor must accept the element type of the space (index or 3D coordinate), or be a single-task functor
    static constexpr unsigned int FuncParamDim = ( ElemND==0 ) ? ND : ElemND ;
    using FuncParamType = std::conditional_t< FuncParamDim==1 , ssize_t , onikaInt3_t >;
    static_assert( lambda_is_compatible_with_v<FuncT,void,FuncParamType> || lambda_is_compatible_with_v<FuncT,void,block_parallel_for_single_task_t>
                 , "User defined functor is not compatible with execution space");

    // the type-erased adapter lives inside the PEC: it must fit the functor area and respect its alignment
    using Adapter = BlockParallelForHostAdapter< FuncT , functor_gpu_support_v<FuncT> , ND,ElemND,ElementListT >;
    [[maybe_unused]] static constexpr AssertFunctorSizeFitIn< sizeof(Adapter) , HostKernelExecutionScratch::MAX_FUNCTOR_SIZE , FuncT > _check_functor_size = {};
    static_assert( ( HostKernelExecutionScratch::MAX_FUNCTOR_ALIGNMENT % alignof(FuncT) ) == 0 , "functor_data alignment is not sufficient for user functor" );
    assert( pec != nullptr );

    pec->m_execution_end_callback = opts.user_cb;
    pec->m_omp_sched = opts.omp_scheduling;

    bool on_device = false;
    if constexpr ( functor_gpu_support_v<FuncT> )
      on_device = details::configure_device_target<ND>( pec , opts , ElemND >= 1 , par_space.m_start.data() , par_space.m_end.data() );
    if( ! on_device ) pec->m_execution_target = ParallelExecutionContext::EXECUTION_TARGET_OPENMP;   // host node in the lane

    pec->m_host_scratch.functor_data_size = sizeof(Adapter);
    new(pec->m_host_scratch.functor_data) Adapter( func , par_space );
    return { pec };
  }

  template<class FuncT>
  static inline ParallelExecutionWrapper
  block_parallel_for( size_t N , const FuncT& func , ParallelExecutionContext * pec , const BlockParallelForOptions& opts = BlockParallelForOptions{} )
  {
    return block_parallel_for( ParallelExecutionSpace<1>{ {0} , {ssize_t(N)} } , func , pec , opts );
  }

} }
