// onika/extras/block_parallel_value_add_functor.h -- the functors the tutorials and lane tests launch:
// BlockParallelValueAddFunctor (a[i][j] += v; memory bound) and BlockParallelIterativeValueAddFunctor (a pow/sin chain
// per element; compute bound).  Row i is one block item; its columns are shared by the threads of the block.
#pragma once
#include <cmath>
#include <onika/cuda/cuda.h>
#include <onika/cuda/cuda_context.h>
#include <onika/extras/array2d.h>
#include <onika/parallel/block_parallel_for_functor.h>
#include <onika/parallel/block_parallel_for_adapter.h>
#include <onika_b200.h>

namespace onika { namespace extras {

  template<bool AllowGpuExec = true, bool SingleTaskMode = false>
  struct BlockParallelValueAddFunctor
  {
    Array2DReference m_array;
    double m_value_to_add = 0.0;

    // whole array in one host call (single-task mode)
    ONIKA_HOST_DEVICE_FUNC inline void operator () ( parallel::block_parallel_for_single_task_t ) const requires(SingleTaskMode)
    {
      const size_t rows = m_array.rows() , cols = m_array.columns();
      for(size_t i=0;i<rows;i++) for(size_t j=0;j<cols;j++) m_array[i][j] += m_value_to_add;
    }
    // one row per block item
    ONIKA_HOST_DEVICE_FUNC inline void operator () (size_t i) const
    {
      const size_t cols = m_array.columns();
      double * __restrict__ row = m_array[i];
      ONIKA_CU_BLOCK_SIMD_FOR(size_t, j, 0, cols) { row[j] += m_value_to_add; }
    }
    // one element per item of a 2D space.  The whole block is called for the item: exactly one thread updates it, so the
    // device result equals the host (one thread per block) result instead of depending on a write race
    ONIKA_HOST_DEVICE_FUNC inline void operator () (const onikaInt3_t& c) const { if( ONIKA_CU_THREAD_IDX == 0 ) m_array[c.x][c.y] += m_value_to_add; }
  };

  template<bool AllowGpuExec = true>
  struct BlockParallelIterativeValueAddFunctor
  {
    Array2DReference m_array;
    double m_value_to_add = 0.0;
    const long m_iterations = 0;

    ONIKA_HOST_DEVICE_FUNC inline double chain(double x) const
    {
      double y = pow( x , sin(x) );
      for(long k=0;k<m_iterations;k++) { x = x + y + m_value_to_add; y = pow( x , sin(x) ); }
      return x * y;
    }
    ONIKA_HOST_DEVICE_FUNC inline void operator () (size_t i) const
    {
      const size_t cols = m_array.columns();
      double * __restrict__ row = m_array[i];
      ONIKA_CU_BLOCK_SIMD_FOR(size_t, j, 0, cols) { row[j] = chain( row[j] ); }
    }
    ONIKA_HOST_DEVICE_FUNC inline void operator () (const onikaInt3_t& c) const { if( ONIKA_CU_THREAD_IDX == 0 ) m_array[c.x][c.y] = chain( m_array[c.x][c.y] ); }
  };

} 

namespace parallel
{
  // rows [start,end) of a dense row-major Array2D (and full-width 2D boxes) are ONE contiguous stream of doubles:
  // onk_value_add_f64 (256-bit vectorised, persistent grid) does the whole operation
  template<unsigned int ND>
  struct B200TypedFastPath< extras::BlockParallelValueAddFunctor<true,false> , ND , 0 >
  {
    static inline constexpr bool available = ( ND <= 2 );
    template<class SpaceT> static inline bool launch( const extras::BlockParallelValueAddFunctor<true,false>& f , const SpaceT& sp , int lane )
    {
      const ssize_t r0 = sp.m_start[0] , r1 = sp.m_end[0];
      if( r0 < 0 || r1 > ssize_t(f.m_array.rows()) ) return false;
      if constexpr ( ND == 2 ) { if( sp.m_start[1] != 0 || sp.m_end[1] != ssize_t(f.m_array.columns()) ) return false; }
      if( r1 <= r0 ) return true;
      ONIKA_B200_CHECK( onk_value_add_f64( f.m_array[r0] , uint64_t(r1-r0) , f.m_array.columns() , f.m_value_to_add , lane ) );
      return true;
    }
  };

  template<bool GpuEnable, bool SingleTaskMode>
  struct BlockParallelForFunctorTraits< extras::BlockParallelValueAddFunctor<GpuEnable,SingleTaskMode> > { static constexpr bool CudaCompatible = GpuEnable && !SingleTaskMode; };
  template<bool GpuEnable>
  struct BlockParallelForFunctorTraits< extras::BlockParallelIterativeValueAddFunctor<GpuEnable> > { static constexpr bool CudaCompatible = GpuEnable; };
}

}
