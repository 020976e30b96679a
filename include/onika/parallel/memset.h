// onika/parallel/memset.h -- parallel_memset(data, N, value, pec): data[i] = value through parallel_for.
#pragma once
#include <onika/parallel/parallel_for.h>

namespace onika { namespace parallel {

  template<class T>
  struct MemSetFunctor
  {
    T * __restrict__ m_data = nullptr;
    T m_value = {};
    ONIKA_HOST_DEVICE_FUNC inline void operator () ( size_t i ) const { m_data[i] = m_value; }
  };
  template<class T> struct ParallelForFunctorTraits< MemSetFunctor<T> > { static inline constexpr bool CudaCompatible = true; };

  // parallel_memset of 1/2/4/8-byte trivially copyable values is the typed onk_parallel_memset kernel (256-bit stores)
  template<class T>
  struct B200TypedFastPath< ParallelForBlockAdapter< MemSetFunctor<T> , ParallelExecutionSpace<1,0> > , 1 , 0 >
  {
    static inline constexpr bool available = std::is_trivially_copyable_v<T> && ( sizeof(T)==1 || sizeof(T)==2 || sizeof(T)==4 || sizeof(T)==8 );
    template<class SpaceT> static inline bool launch( const ParallelForBlockAdapter< MemSetFunctor<T> , ParallelExecutionSpace<1,0> >& a , const SpaceT& , int lane )
    {
      uint64_t pattern = 0;
      std::memcpy( &pattern , & a.m_func.m_value , sizeof(T) );
      ONIKA_B200_CHECK( onk_parallel_memset( a.m_func.m_data + a.m_offset.x , a.m_dims.x , pattern , int(sizeof(T)) , lane ) );
      return true;
    }
  };

  template<class T>
  static inline ParallelExecutionWrapper parallel_memset( T * data , uint64_t N, const T& value, ParallelExecutionContext * pec )
  {
    return parallel_for( N , MemSetFunctor<T>{ data , value } , pec );
  }

} }
