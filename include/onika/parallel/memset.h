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

  template<class T>
  static inline ParallelExecutionWrapper parallel_memset( T * data , uint64_t N, const T& value, ParallelExecutionContext * pec )
  {
    return parallel_for( N , MemSetFunctor<T>{ data , value } , pec );
  }

} }
