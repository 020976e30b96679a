// onika/parallel/reduction.h -- parallel_reduce_min / _max / _sum (fp64) as queueable parallel operations.
//
// The reference offers no reduction entry point: users accumulate into the PEC's return data with atomics from inside a
// functor or, in tests/gpu_reduce_min_fp64.cu, leave the GPU part as a TODO.  This helper packages
// the idiom: the result travels through the standard return-data protocol (host value in before, result out after, see
// ParallelExecutionQueue::schedule), the operation is ordered in its lane like any block_parallel_for, and on the device
// it is the typed sm_100a kernel (onk_reduce_f64: 256-bit streaming loads, two-pass finish inside one launch).
// Semantics: MIN/MAX compare like the reference's OpenMP loop (NaNs ignored); SUM uses a fixed, reproducible tree.
#pragma once
#include <cfloat>
#include <onika/parallel/block_parallel_for.h>
#include <onika/cuda/cuda_reduction.h>
#include <onika_b200.h>

namespace onika { namespace parallel {

  template<int OP>
  struct ReduceF64Functor
  {
    const double * __restrict__ m_data = nullptr;
    size_t m_size = 0;
    double * m_result = nullptr;     // device return-data slot (device target) or the host result (host target)

    // generic body (used when the typed fast path is disabled): one block folds the whole range
    ONIKA_HOST_DEVICE_FUNC inline void operator () ( size_t ) const
    {
      double acc = ( OP == ONK_OP_MIN ) ? DBL_MAX : ( ( OP == ONK_OP_MAX ) ? -DBL_MAX : 0.0 );
      ONIKA_CU_BLOCK_SIMD_FOR(size_t, i, 0, m_size)
      {
        const double v = m_data[i];
        if constexpr ( OP == ONK_OP_MIN ) { if( v < acc ) acc = v; }
        else if constexpr ( OP == ONK_OP_MAX ) { if( v > acc ) acc = v; }
        else acc += v;
      }
      if constexpr ( OP == ONK_OP_MIN ) acc = onika::cuda::block_reduce_min( acc , onika::IntConst<ONIKA_CU_MAX_THREADS_PER_BLOCK>{} );
      else if constexpr ( OP == ONK_OP_MAX ) acc = onika::cuda::block_reduce_max( acc , onika::IntConst<ONIKA_CU_MAX_THREADS_PER_BLOCK>{} );
      else acc = onika::cuda::block_reduce_add( acc , onika::IntConst<ONIKA_CU_MAX_THREADS_PER_BLOCK>{} );
      if( ONIKA_CU_THREAD_IDX == 0 ) *m_result = acc;
    }
  };

  template<int OP> struct BlockParallelForFunctorTraits< ReduceF64Functor<OP> > { static inline constexpr bool CudaCompatible = true; };

  template<int OP>
  struct B200TypedFastPath< ReduceF64Functor<OP> , 1 , 0 >
  {
    static inline constexpr bool available = true;
    template<class SpaceT> static inline bool launch( const ReduceF64Functor<OP>& f , const SpaceT& , int lane )
    {
      ONIKA_B200_CHECK( onk_reduce_f64( OP , f.m_data , f.m_size , f.m_result , lane ) );
      return true;
    }
  };

  namespace details
  {
    template<int OP>
    static inline ParallelExecutionWrapper parallel_reduce_f64( const double* data , size_t n , double* result , ParallelExecutionContext* pec )
    {
      const bool device = ( pec->m_cuda_ctx != nullptr && pec->m_cuda_ctx->has_devices() && functor_gpu_support_v< ReduceF64Functor<OP> > );
      double* slot = result;
      if( device ) { pec->init_device_scratch(); slot = static_cast<double*>( pec->get_device_return_data_ptr() ); }
      BlockParallelForOptions opts;
      opts.return_data = result;
      opts.return_data_size = sizeof(double);
      opts.max_block_size = ONIKA_CU_MAX_THREADS_PER_BLOCK;
      return block_parallel_for( 1 , ReduceF64Functor<OP>{ data , n , slot } , pec , opts );
    }
  }

  // *result receives the reduction once the operation has completed (after synchronize / wrapper destruction)
  static inline ParallelExecutionWrapper parallel_reduce_min( const double* data , size_t n , double* result , ParallelExecutionContext* pec ) { return details::parallel_reduce_f64<ONK_OP_MIN>( data , n , result , pec ); }
  static inline ParallelExecutionWrapper parallel_reduce_max( const double* data , size_t n , double* result , ParallelExecutionContext* pec ) { return details::parallel_reduce_f64<ONK_OP_MAX>( data , n , result , pec ); }
  static inline ParallelExecutionWrapper parallel_reduce_sum( const double* data , size_t n , double* result , ParallelExecutionContext* pec ) { return details::parallel_reduce_f64<ONK_OP_SUM>( data , n , result , pec ); }

} }
