// onika/flat_tuple.h -- a trivially copyable tuple that can cross the host/device boundary inside a kernel argument.
#pragma once
#include <cstddef>
#include <type_traits>
#include <utility>
#include <onika/cuda/cuda.h>

namespace onika
{
  template<size_t I> struct TupleIndex { static constexpr size_t value = I; };
  template<size_t I> static inline constexpr TupleIndex<I> tuple_index = {};

  namespace flat_tuple_details
  {
    template<size_t I, class T> struct Leaf { T v; };
    template<class Seq, class... T> struct Storage;
    template<size_t... I, class... T> struct Storage<std::index_sequence<I...>, T...> : Leaf<I,T>... {};
  }

  template<class... T>
  struct FlatTuple : flat_tuple_details::Storage<std::index_sequence_for<T...>, T...>
  {
    static constexpr size_t tuple_size = sizeof...(T);
    FlatTuple() = default;
    ONIKA_HOST_DEVICE_FUNC constexpr FlatTuple(const T&... a) requires (sizeof...(T) > 0) { assign(std::index_sequence_for<T...>{}, a...); }
    template<size_t I> ONIKA_HOST_DEVICE_FUNC constexpr auto& get(TupleIndex<I> = {}) { return leaf<I>(*this).v; }
    template<size_t I> ONIKA_HOST_DEVICE_FUNC constexpr const auto& get(TupleIndex<I> = {}) const { return leaf<I>(*this).v; }
    template<size_t I> ONIKA_HOST_DEVICE_FUNC constexpr auto& get_nth() { return leaf<I>(*this).v; }
    template<size_t I> ONIKA_HOST_DEVICE_FUNC constexpr const auto& get_nth() const { return leaf<I>(*this).v; }
    static constexpr size_t size() { return sizeof...(T); }
  private:
    template<size_t I, class U> ONIKA_HOST_DEVICE_FUNC static constexpr flat_tuple_details::Leaf<I,U>& leaf(flat_tuple_details::Leaf<I,U>& l) { return l; }
    template<size_t I, class U> ONIKA_HOST_DEVICE_FUNC static constexpr const flat_tuple_details::Leaf<I,U>& leaf(const flat_tuple_details::Leaf<I,U>& l) { return l; }
    template<size_t... I> ONIKA_HOST_DEVICE_FUNC constexpr void assign(std::index_sequence<I...>, const T&... a) { ( ... , ( leaf<I>(*this).v = a ) ); }
  };
  template<> struct FlatTuple<> { static constexpr size_t tuple_size = 0; static constexpr size_t size() { return 0; } };

  template<class Tp, size_t I> struct FlatTupleElement;
  template<size_t I, class... T> struct FlatTupleElement<FlatTuple<T...>,I> { using type = std::tuple_element_t<I,std::tuple<T...>>; };
  template<class Tp, size_t I> using flat_tuple_element_t = typename FlatTupleElement<Tp,I>::type;

  template<class... T> ONIKA_HOST_DEVICE_FUNC constexpr FlatTuple<T...> make_flat_tuple(const T&... a) { return FlatTuple<T...>(a...); }
}
