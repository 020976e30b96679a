// onika/oarray.h -- fixed-size array and in-place vector usable in kernels (no heap, trivially copyable).
#pragma once
#include <cstddef>
#include <cstdint>
#include <compare>
#include <concepts>
#include <functional>
#include <initializer_list>
#include <type_traits>
#include <utility>
#include <onika/cuda/cuda.h>

namespace onika
{
  template<class T, size_t N>
  struct oarray_t
  {
    static inline constexpr size_t array_size = N;
    using value_type = T;
    T m_data[ N > 0 ? N : 1 ];
    ONIKA_HOST_DEVICE_FUNC inline T& operator [] (size_t i) { return m_data[i]; }
    ONIKA_HOST_DEVICE_FUNC inline const T& operator [] (size_t i) const { return m_data[i]; }
    ONIKA_HOST_DEVICE_FUNC inline T* data() { return m_data; }
    ONIKA_HOST_DEVICE_FUNC inline const T* data() const { return m_data; }
    static inline constexpr size_t size() { return N; }
    ONIKA_HOST_DEVICE_FUNC inline T* begin() { return m_data; }
    ONIKA_HOST_DEVICE_FUNC inline const T* begin() const { return m_data; }
    ONIKA_HOST_DEVICE_FUNC inline T* end() { return m_data + N; }
    ONIKA_HOST_DEVICE_FUNC inline const T* end() const { return m_data + N; }

    // lexicographic order, coordinate 0 most significant (include/onika/oarray.h:49-58 in the reference); host and device
    template<class U> requires std::totally_ordered_with<T,U>
    ONIKA_HOST_DEVICE_FUNC inline std::strong_ordering operator <=> (const oarray_t<U,N>& o) const
    {
      for (size_t k = 0; k < N; k++)
      {
        if (m_data[k] < o.m_data[k]) return std::strong_ordering::less;
        if (m_data[k] > o.m_data[k]) return std::strong_ordering::greater;
      }
      return std::strong_ordering::equal;
    }
    template<class U> requires std::equality_comparable_with<T,U>
    ONIKA_HOST_DEVICE_FUNC inline bool operator == (const oarray_t<U,N>& o) const
    {
      for (size_t k = 0; k < N; k++) if (!(m_data[k] == o.m_data[k])) return false;
      return true;
    }
  };

  template<class T, size_t MaxSize>
  struct inplace_vector
  {
    static inline constexpr uint32_t max_size = MaxSize;
    using value_type = T;
    T m_data[MaxSize];
    uint32_t m_size = 0;
    inplace_vector() = default;
    inline inplace_vector(std::initializer_list<T> l) { for (const auto& x : l) push_back(x); }
    ONIKA_HOST_DEVICE_FUNC inline T& operator [] (size_t i) { return m_data[i]; }
    ONIKA_HOST_DEVICE_FUNC inline const T& operator [] (size_t i) const { return m_data[i]; }
    ONIKA_HOST_DEVICE_FUNC inline T* data() { return m_data; }
    ONIKA_HOST_DEVICE_FUNC inline const T* data() const { return m_data; }
    ONIKA_HOST_DEVICE_FUNC inline size_t size() const { return m_size; }
    ONIKA_HOST_DEVICE_FUNC inline bool empty() const { return m_size == 0; }
    ONIKA_HOST_DEVICE_FUNC inline void push_back(const T& v) { m_data[m_size++] = v; }
    ONIKA_HOST_DEVICE_FUNC inline void emplace_back(T&& v) { m_data[m_size++] = static_cast<T&&>(v); }
    ONIKA_HOST_DEVICE_FUNC inline void pop_back() { --m_size; }
    ONIKA_HOST_DEVICE_FUNC inline void clear() { m_size = 0; }
    ONIKA_HOST_DEVICE_FUNC inline T* begin() { return m_data; }
    ONIKA_HOST_DEVICE_FUNC inline const T* begin() const { return m_data; }
    ONIKA_HOST_DEVICE_FUNC inline T* end() { return m_data + m_size; }
    ONIKA_HOST_DEVICE_FUNC inline const T* end() const { return m_data + m_size; }
  };

  template<class T> struct IsOArray : std::false_type {};
  template<class T, size_t N> struct IsOArray< oarray_t<T,N> > : std::true_type {};
  template<class T> static inline constexpr bool is_oarray_v = IsOArray<T>::value;
  template<class T> static inline constexpr bool is_integral_oarray_v = IsOArray<T>::value;
  template<class T> concept some_oarray_t = is_oarray_v<T>;
  template<class T> concept some_integral_oarray_t = is_oarray_v<T> && std::is_integral_v<typename T::value_type>;
}

namespace std
{
  template<onika::some_integral_oarray_t A> struct hash<A>
  {
    inline size_t operator () (const A& a) const
    {
      size_t h = 1469598103934665603ull;
      for (auto x : a) { h ^= std::hash<typename A::value_type>{}(x); h *= 1099511628211ull; }
      return h;
    }
  };
}
