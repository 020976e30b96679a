// onika/type_utils.h -- the few type traits of the reference's type_utils.h that headers of the data-parallel path use
// (span detection for element lists and field/span tuples, pointer stripping that understands __restrict__).
#pragma once
#include <cstddef>
#include <span>
#include <type_traits>

namespace onika
{
  template<class T> struct remove_pointer { typedef T type; };
  template<class T> struct remove_pointer<T*> { typedef T type; };
  template<class T> struct remove_pointer<T* const> { typedef T type; };
  template<class T> struct remove_pointer<T* volatile> { typedef T type; };
  template<class T> struct remove_pointer<T* const volatile> { typedef T type; };
  template<class T> struct remove_pointer<T* __restrict__ > { typedef T type; };
  template<class T> struct remove_pointer<T* __restrict__ const> { typedef T type; };
  template<class T> struct remove_pointer<T* __restrict__ volatile> { typedef T type; };
  template<class T> struct remove_pointer<T* __restrict__ const volatile> { typedef T type; };
  template<class T> using remove_pointer_t = typename ::onika::remove_pointer<T>::type ;

  // std::span, onika::cuda::span (cuda/stl_adaptors.h) and onika::cuda::InputArraySpan (cuda/input_array_span.h) count as spans
  template<class T> struct is_span_t : public std::false_type {};
  template<class T, std::size_t Extent> struct is_span_t< std::span<T,Extent> > : public std::true_type {};
  template<class T> static inline constexpr bool is_span_v = is_span_t<T>::value;
  template<typename T> concept SpanCompatible = is_span_v<T>;
}
