// onika/flat_tuple.h -- a trivially copyable tuple that can cross the host/device boundary inside a kernel argument.
#pragma once
#include <cstddef>
#include <type_traits>
#include <utility>
#include <onika/cuda/cuda.h>

namespace onika
{
  template<size_t I> struct TupleIndex { static constexpr size_t value = I; };
  template<size_t I> using tuple_index_t = TupleIndex<I>;                       // the reference's spelling of the index tag type
  template<size_t I> static inline constexpr TupleIndex<I> tuple_index = {};
  static inline constexpr size_t FlatTupleMaxElements = 256;                    // the reference generates 40 specialisations; no such limit here

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
    template<size_t I> ONIKA_HOST_DEVICE_FUNC constexpr const auto& get_nth_const() const { return leaf<I>(*this).v; }
    template<size_t I> ONIKA_HOST_DEVICE_FUNC constexpr auto get_copy(TupleIndex<I> = {}) const { return leaf<I>(*this).v; }
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

  // element types are the decayed argument types (arrays and functions become pointers, cv and references drop)
  template<class... Types> ONIKA_HOST_DEVICE_FUNC constexpr FlatTuple< std::decay_t<Types>... > make_flat_tuple(Types&&... a)
  {
    return FlatTuple< std::decay_t<Types>... >( static_cast<Types&&>(a)... );
  }

  // the elements Is... of a tuple, as a new tuple
  template<class _Tuple, size_t... Is>
  static inline FlatTuple< flat_tuple_element_t<_Tuple,Is> ... > flat_tuple_subset( const _Tuple& v , std::integer_sequence<std::size_t,Is...> )
  {
    return FlatTuple< flat_tuple_element_t<_Tuple,Is> ... >( v.get( tuple_index<Is> ) ... );
  }

  template<class FlatTupleT> struct TupleSizeConst { static inline constexpr size_t value = FlatTupleT::size(); };
  template<class FlatTupleT> static inline constexpr size_t tuple_size_const_v = TupleSizeConst<FlatTupleT>::value;
}
