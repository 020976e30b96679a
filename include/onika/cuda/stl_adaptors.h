// onika/cuda/stl_adaptors.h -- data()/size() of std::vector-like containers callable from host-device functions
// (std::vector members are host-only; these read the same three-pointer layout libstdc++ uses).
#pragma once
#include <vector>
#include <onika/cuda/cuda.h>

namespace onika { namespace cuda {

  template<class T, class A>
  ONIKA_HOST_DEVICE_FUNC inline T* vector_data(std::vector<T,A>& v)
  {
    return * reinterpret_cast<T**>( &v );                  // _M_start is the first member of libstdc++'s vector
  }
  template<class T, class A>
  ONIKA_HOST_DEVICE_FUNC inline const T* vector_data(const std::vector<T,A>& v)
  {
    return * reinterpret_cast<T* const*>( &v );
  }
  template<class T, class A>
  ONIKA_HOST_DEVICE_FUNC inline size_t vector_size(const std::vector<T,A>& v)
  {
    T* const* p = reinterpret_cast<T* const*>( &v );
    return size_t( p[1] - p[0] );
  }

  // pointer + size view usable in kernels
  // std::pair / std::lower_bound stand-ins usable in device code
  template<class T, class U> struct pair
  {
    T first; U second;
    ONIKA_HOST_DEVICE_FUNC inline bool operator == ( const pair & o ) const { return first == o.first && second == o.second; }
    ONIKA_HOST_DEVICE_FUNC inline bool operator != ( const pair & o ) const { return ! ( *this == o ); }
    ONIKA_HOST_DEVICE_FUNC inline bool operator <  ( const pair & o ) const { return first < o.first || ( first == o.first && second < o.second ); }
    ONIKA_HOST_DEVICE_FUNC inline bool operator >= ( const pair & o ) const { return ! ( *this < o ); }
  };
  // first position in [first,last) whose element is not less than x; the range must be sorted (binary search)
  template<class Iterator, class T> ONIKA_HOST_DEVICE_FUNC inline Iterator lower_bound( Iterator first , Iterator last , const T & x )
  {
    auto n = last - first;
    while( n > 0 )
    {
      const auto half = n / 2;
      Iterator mid = first + half;
      if( *mid < x ) { first = mid + 1; n -= half + 1; } else n = half;
    }
    return first;
  }

  template<class T> struct span
  {
    T* m_ptr = nullptr; size_t m_size = 0;
    ONIKA_HOST_DEVICE_FUNC inline T* data() const { return m_ptr; }
    ONIKA_HOST_DEVICE_FUNC inline size_t size() const { return m_size; }
    ONIKA_HOST_DEVICE_FUNC inline T& operator [] (size_t i) const { return m_ptr[i]; }
    ONIKA_HOST_DEVICE_FUNC inline T* begin() const { return m_ptr; }
    ONIKA_HOST_DEVICE_FUNC inline T* end() const { return m_ptr + m_size; }
  };
  template<class T, class A> inline span<T> make_span(std::vector<T,A>& v) { return { v.data() , v.size() }; }
  template<class T, class A> inline span<const T> make_const_span(const std::vector<T,A>& v) { return { v.data() , v.size() }; }

} }
