// onika/cuda/stl_adaptors.h -- the small std:: stand-ins kernel functors use: data()/size() of a std::vector from
// host-device functions, pair, lower_bound, span, a printf-backed cout, a forward memmove.
//
// Same names, members and meaning as the reference header (include/onika/cuda/stl_adaptors.h there).  Two deliberate
// differences: pair::operator== compares BOTH members (the reference joins its two comparisons with `||`, so that == and !=
// are not complements there), and memmove also copies backwards (the reference aborts when dst > src).
#pragma once
#include <cstdint>
#include <cstdio>
#include <ranges>
#include <vector>
#include <onika/cuda/cuda.h>
#include <onika/cuda/ro_shallow_copy.h>
#include <onika/type_utils.h>

namespace onika { namespace cuda {

  template<class T, class U> struct pair
  {
    T first; U second;
    ONIKA_HOST_DEVICE_FUNC inline bool operator == ( const pair & o ) const { return first == o.first && second == o.second; }
    ONIKA_HOST_DEVICE_FUNC inline bool operator != ( const pair & o ) const { return first != o.first || second != o.second; }
    ONIKA_HOST_DEVICE_FUNC inline bool operator <  ( const pair & o ) const { return first < o.first || ( first == o.first && second < o.second ); }
    ONIKA_HOST_DEVICE_FUNC inline bool operator >= ( const pair & o ) const { return first > o.first || ( first == o.first && second >= o.second ); }
  };

  // byte copy between possibly overlapping ranges, one thread; 8 bytes at a time when both ends allow it
  ONIKA_HOST_DEVICE_FUNC inline void memmove( void* dst , const void* src , unsigned int n )
  {
    volatile uint8_t* d1 = (volatile uint8_t*) dst;
    const volatile uint8_t* s1 = (const volatile uint8_t*) src;
    const bool words = ( ( (uintptr_t) dst | (uintptr_t) src ) & 7u ) == 0;
    if( dst < src )
    {
      unsigned int i = 0;
      if( words ) { volatile uint64_t* d8 = (volatile uint64_t*) dst; const volatile uint64_t* s8 = (const volatile uint64_t*) src; for(; i<n/8; i++) d8[i] = s8[i]; i *= 8; }
      for(; i<n; i++) d1[i] = s1[i];
    }
    else if( dst > src )
    {
      unsigned int i = n;
      for(; i>0 && ( !words || (i&7u) != 0 ); i--) d1[i-1] = s1[i-1];
      if( words ) { volatile uint64_t* d8 = (volatile uint64_t*) dst; const volatile uint64_t* s8 = (const volatile uint64_t*) src; for(unsigned int k=i/8; k>0; k--) d8[k-1] = s8[k-1]; }
    }
  }

  // std::vector members are host-only: these read the begin / end pointers libstdc++'s vector starts with
  template<class T, class A> ONIKA_HOST_DEVICE_FUNC inline T* vector_data( std::vector<T,A>& v ) { return * reinterpret_cast<T**>( &v ); }
  template<class T, class A> ONIKA_HOST_DEVICE_FUNC inline const T* vector_data( const std::vector<T,A>& v ) { return * reinterpret_cast<T* const*>( &v ); }
  template<class T, class A> ONIKA_HOST_DEVICE_FUNC inline size_t vector_size( const std::vector<T,A>& v )
  {
    static_assert( sizeof(std::vector<T,A>) >= 2 * sizeof(T*) );
    T* const* p = reinterpret_cast<T* const*>( &v );
    return size_t( p[1] - p[0] );
  }
  template<class T> ONIKA_HOST_DEVICE_FUNC inline size_t vector_size( const VectorShallowCopy<T>& v ) { return v.size(); }
  template<class T> ONIKA_HOST_DEVICE_FUNC inline const T* vector_data( const VectorShallowCopy<T>& v ) { return v.data(); }

  // first position whose element is not less than x, scanning from `begin` (the reference's definition: it does not require
  // a sorted range, so neither does this; on a sorted range it is std::lower_bound)
  template<class Iterator, class T> ONIKA_HOST_DEVICE_FUNC inline Iterator lower_bound( Iterator begin , Iterator end , const T & x )
  {
    for(; begin != end ; ++begin) { if( ! ( (*begin) < x ) ) return begin; }
    return end;
  }

  // pointer + size view usable in kernels
  template<class T> struct span
  {
    using value_type = T;
    T * m_start = nullptr;
    size_t m_size = 0;
    ONIKA_HOST_DEVICE_FUNC inline T * data() { return m_start; }
    ONIKA_HOST_DEVICE_FUNC inline const T * data() const { return m_start; }
    ONIKA_HOST_DEVICE_FUNC inline T& operator [] (size_t i) { return m_start[i]; }
    ONIKA_HOST_DEVICE_FUNC inline const T& operator [] (size_t i) const { return m_start[i]; }
    ONIKA_HOST_DEVICE_FUNC inline size_t size() const { return m_size; }
    ONIKA_HOST_DEVICE_FUNC inline bool empty() const { return m_size == 0; }
    ONIKA_HOST_DEVICE_FUNC inline auto begin() const { return m_start; }
    ONIKA_HOST_DEVICE_FUNC inline auto end() const { return m_start + m_size; }
  };

  // `onika::cuda::cout << "text" << value` from host or device code (printf underneath)
  struct PrintfBaseStdOutStream
  {
    ONIKA_HOST_DEVICE_FUNC inline PrintfBaseStdOutStream operator << ( const char* s ) const { printf("%s",s); return {}; }
    ONIKA_HOST_DEVICE_FUNC inline PrintfBaseStdOutStream operator << ( uint64_t i ) const { printf("%llu",static_cast<unsigned long long>(i)); return {}; }
    ONIKA_HOST_DEVICE_FUNC inline PrintfBaseStdOutStream operator << ( int64_t i ) const { printf("%lld",static_cast<long long>(i)); return {}; }
    ONIKA_HOST_DEVICE_FUNC inline PrintfBaseStdOutStream operator << ( uint32_t i ) const { printf("%llu",static_cast<unsigned long long>(i)); return {}; }
    ONIKA_HOST_DEVICE_FUNC inline PrintfBaseStdOutStream operator << ( int32_t i ) const { printf("%lld",static_cast<long long>(i)); return {}; }
    ONIKA_HOST_DEVICE_FUNC inline PrintfBaseStdOutStream operator << ( uint16_t i ) const { printf("%llu",static_cast<unsigned long long>(i)); return {}; }
    ONIKA_HOST_DEVICE_FUNC inline PrintfBaseStdOutStream operator << ( int16_t i ) const { printf("%lld",static_cast<long long>(i)); return {}; }
    ONIKA_HOST_DEVICE_FUNC inline PrintfBaseStdOutStream operator << ( double d ) const { printf("% .9e",d); return {}; }
    ONIKA_HOST_DEVICE_FUNC inline PrintfBaseStdOutStream operator << ( float d ) const { printf("% .9e",double(d)); return {}; }
    ONIKA_HOST_DEVICE_FUNC static inline void flush() {}
  };
  static inline constexpr PrintfBaseStdOutStream cout = {};

  template< std::ranges::contiguous_range R > inline span<const typename R::value_type> make_const_span( const R& r ) { return { r.data() , r.size() }; }
  template< std::ranges::contiguous_range R > inline span<typename R::value_type> make_span( R& r ) { return { r.data() , r.size() }; }

} }

namespace onika
{
  template<class T> struct is_span_t< ::onika::cuda::span<T> > : public std::true_type {};
}
