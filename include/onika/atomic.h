// onika/atomic.h -- host-side read-modify-write atomics on plain variables (what ONIKA_CU_ATOMIC_* mean in host code).
//
// Same names and return values as the reference helpers (include/onika/atomic.h:14-62 there): capture_* return the value
// the variable held BEFORE the update, inplace_* return nothing.  The reference builds add / sub on `#pragma omp atomic`;
// these are compiler atomics (fetch-add for integers, compare-and-swap for floating point), so they are also atomic with
// respect to threads that are not part of an OpenMP team (the lane worker thread, CUDA callback threads).
#pragma once
#include <type_traits>

namespace onika
{
  namespace atomic_details
  {
    template<class T, class Op>
    static inline T capture_rmw(T& x, Op op)
    {
      static_assert( std::is_trivially_copyable_v<T> && ( sizeof(T)==1 || sizeof(T)==2 || sizeof(T)==4 || sizeof(T)==8 ) , "lock-free host atomics need a 1/2/4/8 byte trivially copyable type" );
      T seen , want;
      __atomic_load( &x , &seen , __ATOMIC_SEQ_CST );
      do { want = op( seen ); } while( ! __atomic_compare_exchange( &x , &seen , &want , false , __ATOMIC_SEQ_CST , __ATOMIC_SEQ_CST ) );
      return seen;
    }
  }

  template<class T>
  static inline T capture_atomic_add(T& x , const T& a)
  {
    if constexpr ( std::is_integral_v<T> ) return __atomic_fetch_add( &x , a , __ATOMIC_SEQ_CST );
    else return atomic_details::capture_rmw( x , [&](const T& v) -> T { return v + a; } );
  }

  template<class T>
  static inline T capture_atomic_sub(T& x , const T& a)
  {
    if constexpr ( std::is_integral_v<T> ) return __atomic_fetch_sub( &x , a , __ATOMIC_SEQ_CST );
    else return atomic_details::capture_rmw( x , [&](const T& v) -> T { return v - a; } );
  }

  template<class T> static inline void inplace_atomic_add(T& x , const T& a) { (void) capture_atomic_add( x , a ); }
  template<class T> static inline void inplace_atomic_sub(T& x , const T& a) { (void) capture_atomic_sub( x , a ); }

  // the stored value becomes min(x,a) / max(x,a); ties and NaN arguments leave it unchanged, as `( xval < a ) ? xval : a` does
  template<class T>
  static inline T capture_atomic_min(T& x , const T& a)
  {
    return atomic_details::capture_rmw( x , [&](const T& v) -> T { return ( v < a ) ? v : a; } );
  }

  template<class T>
  static inline T capture_atomic_max(T& x , const T& a)
  {
    return atomic_details::capture_rmw( x , [&](const T& v) -> T { return ( v > a ) ? v : a; } );
  }
}
