// onika/soatl/field_tuple.h -- one element of a field-array container as a value (AoS view): FieldTuple<ids...>.
// Default construction zeroes every field; tuples over different field sets convert into each other field by field
// (fields the source lacks become zero with copy_or_zero_fields, stay untouched with copy_existing_fields).
#pragma once
#include <cstring>
#include <tuple>
#include <type_traits>
#include <onika/cuda/cuda.h>
#include <onika/flat_tuple.h>
#include <onika/soatl/field_id.h>

namespace onika { namespace soatl {

  namespace details
  {
    // "zero" of a field value: 0 for arithmetic types, nullptr for pointers, T() when default constructible, zero bytes otherwise
    template<class T> ONIKA_HOST_DEVICE_FUNC static inline void zero_data( T & x )
    {
      if constexpr ( std::is_arithmetic_v<T> ) x = 0;
      else if constexpr ( std::is_pointer_v<T> ) x = nullptr;
      else if constexpr ( std::is_default_constructible_v<T> ) x = T();
      else std::memset( &x , 0 , sizeof(T) );
    }
  }

  template<typename... ids>
  struct FieldTuple
  {
    static constexpr size_t TupleSize = sizeof...(ids);
    using FieldIdsTuple = FlatTuple< FieldId<ids> ... >;
    using TupleType = FlatTuple< typename FieldId<ids>::value_type ... >;
    using StorageT = TupleType;
    template<class id> using HasField = std::integral_constant< bool , find_index_of_id_v<id,ids...> != bad_field_index >;
    template<class id> static constexpr bool has_field( FieldId<id> ) { return HasField<id>::value; }

    TupleType m_values;

    ONIKA_HOST_DEVICE_FUNC inline FieldTuple() { zero(); }
    FieldTuple( const FieldTuple & ) = default;
    FieldTuple & operator = ( const FieldTuple & ) = default;
    ONIKA_HOST_DEVICE_FUNC inline FieldTuple( const typename FieldId<ids>::value_type & ... v ) requires ( sizeof...(ids) > 0 ) : m_values( v... ) {}
    inline FieldTuple( const std::tuple< typename FieldId<ids>::value_type ... > & t ) { assign_from_std( t , std::index_sequence_for<ids...>{} ); }
    // from a tuple over another field set: common fields are copied, the others zeroed
    template<typename... other> requires ( ! std::is_same_v< FieldTuple<other...> , FieldTuple > )
    ONIKA_HOST_DEVICE_FUNC inline FieldTuple( const FieldTuple<other...> & o ) { copy_or_zero_fields( o ); }

    template<class id> ONIKA_HOST_DEVICE_FUNC inline auto & operator [] ( FieldId<id> )
    {
      static_assert( HasField<id>::value , "field not in tuple" );
      return m_values.template get_nth< find_index_of_id_v<id,ids...> >();
    }
    template<class id> ONIKA_HOST_DEVICE_FUNC inline const auto & operator [] ( FieldId<id> ) const
    {
      static_assert( HasField<id>::value , "field not in tuple" );
      return m_values.template get_nth< find_index_of_id_v<id,ids...> >();
    }
    template<size_t I> ONIKA_HOST_DEVICE_FUNC inline auto & get_nth() { return m_values.template get_nth<I>(); }
    template<size_t I> ONIKA_HOST_DEVICE_FUNC inline const auto & get_nth() const { return m_values.template get_nth<I>(); }

    // f( FieldId<id>{} , value ) for every field, in declaration order
    template<class Func> inline void apply_fields( Func f ) const { ( ... , f( FieldId<ids>{} , (*this)[ FieldId<ids>{} ] ) ); }

    ONIKA_HOST_DEVICE_FUNC inline void zero() { ( ... , details::zero_data( (*this)[ FieldId<ids>{} ] ) ); }

    template<typename... other> ONIKA_HOST_DEVICE_FUNC inline void copy_existing_fields( const FieldTuple<other...> & o ) { ( ... , take<ids,false>( o ) ); }
    template<typename... other> ONIKA_HOST_DEVICE_FUNC inline void copy_or_zero_fields( const FieldTuple<other...> & o ) { ( ... , take<ids,true>( o ) ); }

    ONIKA_HOST_DEVICE_FUNC inline bool operator == ( const FieldTuple & o ) const { return ( ... && ( (*this)[ FieldId<ids>{} ] == o[ FieldId<ids>{} ] ) ); }
    ONIKA_HOST_DEVICE_FUNC inline bool operator != ( const FieldTuple & o ) const { return ! ( *this == o ); }

  private:
    template<class id, bool ZeroIfMissing, typename... other> ONIKA_HOST_DEVICE_FUNC inline void take( const FieldTuple<other...> & o )
    {
      if constexpr ( find_index_of_id_v<id,other...> != bad_field_index ) (*this)[ FieldId<id>{} ] = o[ FieldId<id>{} ];
      else if constexpr ( ZeroIfMissing ) details::zero_data( (*this)[ FieldId<id>{} ] );
    }
    template<class Tp, size_t... I> inline void assign_from_std( const Tp & t , std::index_sequence<I...> )
    {
      ( ... , ( m_values.template get_nth<I>() = std::get<I>(t) ) );
    }
  };

  // the empty tuple holds nothing and has no field
  template<> struct FieldTuple<>
  {
    static constexpr size_t TupleSize = 0;
    using FieldIdsTuple = FlatTuple<>;
    using TupleType = FlatTuple<>;
    template<class id> using HasField = std::false_type;
    template<class id> static constexpr bool has_field( FieldId<id> ) { return false; }
  };

  template<class Fids> struct _FieldTupleFromFieldIds;
  template<class... ids> struct _FieldTupleFromFieldIds< FieldIds<ids...> > { using type = FieldTuple<ids...>; };
  template<class Fids> using FieldTupleFromFieldIds = typename _FieldTupleFromFieldIds<Fids>::type;

  template<class TupleT, class id> static inline constexpr bool field_tuple_has_field_v = TupleT::template HasField<id>::value;

  // zero-initialised tuple over the given fields / tuple holding the given values
  template<typename... ids>
  ONIKA_HOST_DEVICE_FUNC inline FieldTuple<ids...> make_field_tuple( const FieldId<ids> & ... ) { return FieldTuple<ids...>(); }
  template<typename... ids>
  ONIKA_HOST_DEVICE_FUNC inline FieldTuple<ids...> make_field_tuple( const FieldId<ids> & ... , const typename FieldId<ids>::value_type & ... v ) requires ( sizeof...(ids) > 0 )
  { return FieldTuple<ids...>( v... ); }

} }
