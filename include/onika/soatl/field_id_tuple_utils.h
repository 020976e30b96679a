// onika/soatl/field_id_tuple_utils.h -- packed byte sizes of fields: how much room one value of a field (or a span of
// fields, i.e. several values of the same field type) takes when every entry is padded to a pack alignment.
#pragma once
#include <span>
#include <type_traits>
#include <utility>
#include <onika/cuda/cuda.h>
#include <onika/flat_tuple.h>

namespace onika { namespace soatl {

  static inline constexpr size_t DEFAULT_FIELD_PACK_ALIGNMENT = 8;

  namespace details
  {
    template<class T> struct is_std_span : std::false_type {};
    template<class T, size_t E> struct is_std_span< std::span<T,E> > : std::true_type {};
  }

  // one field descriptor (anything with a ::value_type) or a span of descriptors
  template<class FieldOrSpanT>
  ONIKA_HOST_DEVICE_FUNC inline size_t field_id_size_bytes( const FieldOrSpanT & f , size_t pack_alignment = DEFAULT_FIELD_PACK_ALIGNMENT )
  {
    size_t bytes = 0;
    if constexpr ( details::is_std_span< std::remove_cv_t<FieldOrSpanT> >::value )
      bytes = sizeof( typename FieldOrSpanT::value_type::value_type ) * f.size();
    else
      bytes = sizeof( typename FieldOrSpanT::value_type );
    return ( ( bytes + pack_alignment - 1 ) / pack_alignment ) * pack_alignment;
  }

  template<class FieldTupleT, size_t... I>
  ONIKA_HOST_DEVICE_FUNC inline size_t field_id_tuple_size_bytes( const FieldTupleT & ft , std::index_sequence<I...> )
  {
    size_t total = 0;
    ( ... , ( total += field_id_size_bytes( ft.template get_nth<I>() ) ) );
    return total;
  }

  template<class... FieldT>
  ONIKA_HOST_DEVICE_FUNC inline size_t field_id_tuple_size_bytes( const onika::FlatTuple<FieldT...> & ft )
  {
    return field_id_tuple_size_bytes( ft , std::make_index_sequence< sizeof...(FieldT) >{} );
  }

} }
