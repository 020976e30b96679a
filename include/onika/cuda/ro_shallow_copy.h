// onika/cuda/ro_shallow_copy.h -- VectorShallowCopy<T>: pointer+size stand-in for a std::vector inside kernel functors
// (std::vector itself cannot be copied into kernel parameters); ro_shallow_copy_t<X> maps vector types to it.
#pragma once
#include <cassert>
#include <cstdlib>
#include <vector>
#include <onika/cuda/cuda.h>

namespace onika { namespace cuda {

  template<class T>
  struct VectorShallowCopy
  {
    const T* m_data = nullptr;
    size_t m_size = 0;
    VectorShallowCopy() = default;
    template<class A> inline VectorShallowCopy(const std::vector<T,A>& v) : m_data( v.data() ) , m_size( v.size() ) {}
    inline VectorShallowCopy(const T* p, size_t n) : m_data(p) , m_size(n) {}
    ONIKA_HOST_DEVICE_FUNC inline const T* data() const { return m_data; }
    ONIKA_HOST_DEVICE_FUNC inline bool empty() const { return m_size == 0; }
    ONIKA_HOST_DEVICE_FUNC inline size_t size() const { return m_size; }
    ONIKA_HOST_DEVICE_FUNC inline void resize(size_t n) const { assert( n == m_size ); (void)n; }
    ONIKA_HOST_DEVICE_FUNC inline const T& operator [] (size_t i) const { return m_data[i]; }
    ONIKA_HOST_DEVICE_FUNC inline const T* begin() const { return m_data; }
    ONIKA_HOST_DEVICE_FUNC inline const T* end() const { return m_data + m_size; }
  };

  template<class T> struct ReadOnlyShallowCopyType { using type = T; };
  template<class T, class A> struct ReadOnlyShallowCopyType< std::vector<T,A> > { using type = VectorShallowCopy<T>; };
  template<class T> using ro_shallow_copy_t = typename ReadOnlyShallowCopyType<T>::type;

} }
