// onika/cuda/uninitialized_place_holder.h -- raw, correctly aligned storage for objects that may not be initialised at
// declaration (shared-memory variables, in-kernel scratch arrays).
#pragma once
#include <cstddef>
#include <onika/cuda/cuda.h>

namespace onika { namespace cuda {

  template<class T>
  struct UnitializedPlaceHolder
  {
    using value_type = T;
    alignas(T) unsigned char m_bytes[ sizeof(T) ];
    ONIKA_HOST_DEVICE_FUNC inline T& get_ref() { return * reinterpret_cast<T*>( m_bytes ); }
    ONIKA_HOST_DEVICE_FUNC inline const T& get_ref() const { return * reinterpret_cast<const T*>( m_bytes ); }
  };

  template<class T, size_t N>
  struct UnitializedArrayPlaceHolder
  {
    using value_type = T;
    static inline constexpr size_t ArraySize = N;
    alignas(T) unsigned char m_bytes[ sizeof(T) * ( N > 0 ? N : 1 ) ];
    ONIKA_HOST_DEVICE_FUNC inline T* get_array() { return reinterpret_cast<T*>( m_bytes ); }
    ONIKA_HOST_DEVICE_FUNC inline const T* get_array() const { return reinterpret_cast<const T*>( m_bytes ); }
    ONIKA_HOST_DEVICE_FUNC inline operator T* () { return get_array(); }
    ONIKA_HOST_DEVICE_FUNC inline operator const T* () const { return get_array(); }
    ONIKA_HOST_DEVICE_FUNC inline T& operator [] (size_t i) { return get_array()[i]; }
    ONIKA_HOST_DEVICE_FUNC inline const T& operator [] (size_t i) const { return get_array()[i]; }
  };

} }
