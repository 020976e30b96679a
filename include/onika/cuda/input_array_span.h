// onika/cuda/input_array_span.h -- InputArraySpan<T,N>: read-only view of an input array handed to a kernel functor.
// Arrays of at most N elements are COPIED INTO the span, so they travel inside the kernel parameter block
// (__grid_constant__ functor) and the kernel never touches the original (possibly managed) memory; larger arrays are
// referenced by pointer.
#pragma once
#include <cstring>
#include <ranges>
#include <type_traits>
#include <onika/cuda/cuda.h>
#include <onika/cuda/uninitialized_place_holder.h>
#include <onika/type_utils.h>

namespace onika {

  namespace cuda {

  template<class T, size_t N = 1>
  struct InputArraySpan
  {
    static_assert( std::is_trivially_copyable_v<T> , "InputArraySpan elements must be trivially copyable to ride in kernel parameters" );
    using value_type = T;
    static inline constexpr size_t InternalArraySize = N;

    const T* m_data_pointer = nullptr;
    size_t m_size = 0;
    UnitializedArrayPlaceHolder<T,N> m_internal_copy;

    InputArraySpan() = default;
    template< std::ranges::contiguous_range R >
    inline explicit InputArraySpan(const R& r) : m_size( r.size() )
    {
      if( m_size <= N ) { if( m_size > 0 ) std::memcpy( m_internal_copy.get_array() , r.data() , sizeof(T) * m_size ); }
      else m_data_pointer = r.data();
    }
    ONIKA_HOST_DEVICE_FUNC inline void reset() { m_size = 0; m_data_pointer = nullptr; }
    ONIKA_HOST_DEVICE_FUNC inline const T* data() const { return ( m_size <= N ) ? m_internal_copy.get_array() : m_data_pointer; }
    ONIKA_HOST_DEVICE_FUNC inline const T& operator [] (size_t i) const { return data()[i]; }
    ONIKA_HOST_DEVICE_FUNC inline size_t size() const { return m_size; }
    ONIKA_HOST_DEVICE_FUNC inline bool empty() const { return m_size == 0; }
    ONIKA_HOST_DEVICE_FUNC inline const T* begin() const { return data(); }
    ONIKA_HOST_DEVICE_FUNC inline const T* end() const { return data() + m_size; }
  };

  template< std::ranges::contiguous_range R , size_t N = 1 >
  inline InputArraySpan< typename R::value_type , N > make_input_array_span( const R& r , std::integral_constant<size_t,N> = {} )
  {
    return InputArraySpan< typename R::value_type , N >( r );
  }

  }

  template<class T, size_t N> struct is_span_t< ::onika::cuda::InputArraySpan<T,N> > : std::true_type {};
}
