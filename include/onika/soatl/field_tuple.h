// onika/soatl/field_tuple.h -- one element of a field-array container as a value (AoS view): FieldTuple<ids...>.
#pragma once
#include <tuple>
#include <onika/cuda/cuda.h>
#include <onika/flat_tuple.h>
#include <onika/soatl/field_id.h>

namespace onika { namespace soatl {

  template<typename... ids>
  struct FieldTuple
  {
    static constexpr size_t TupleSize = sizeof...(ids);
    using StorageT = FlatTuple< typename FieldId<ids>::value_type ... >;
    StorageT m_values;

    FieldTuple() = default;
    ONIKA_HOST_DEVICE_FUNC inline FieldTuple( const typename FieldId<ids>::value_type & ... v ) requires ( sizeof...(ids) > 0 ) : m_values( v... ) {}
    inline FieldTuple( const std::tuple< typename FieldId<ids>::value_type ... > & t ) { assign_from_std( t , std::index_sequence_for<ids...>{} ); }

    template<class id> static constexpr bool has_field( FieldId<id> ) { return find_index_of_id_v<id,ids...> != bad_field_index; }

    template<class id> ONIKA_HOST_DEVICE_FUNC inline auto & operator [] ( FieldId<id> )
    {
      static_assert( find_index_of_id_v<id,ids...> != bad_field_index , "field not in tuple" );
      return m_values.template get_nth< find_index_of_id_v<id,ids...> >();
    }
    template<class id> ONIKA_HOST_DEVICE_FUNC inline const auto & operator [] ( FieldId<id> ) const
    {
      static_assert( find_index_of_id_v<id,ids...> != bad_field_index , "field not in tuple" );
      return m_values.template get_nth< find_index_of_id_v<id,ids...> >();
    }
    template<size_t I> ONIKA_HOST_DEVICE_FUNC inline auto & get_nth() { return m_values.template get_nth<I>(); }
    template<size_t I> ONIKA_HOST_DEVICE_FUNC inline const auto & get_nth() const { return m_values.template get_nth<I>(); }

    // copy the fields both tuples have; others are left untouched
    template<typename... other> ONIKA_HOST_DEVICE_FUNC inline void copy_existing_fields( const FieldTuple<other...>& o )
    {
      ( ... , copy_one<other>( o ) );
    }
  private:
    template<class oid, typename... other> ONIKA_HOST_DEVICE_FUNC inline void copy_one( const FieldTuple<other...>& o )
    {
      if constexpr ( find_index_of_id_v<oid,ids...> != bad_field_index ) (*this)[ FieldId<oid>{} ] = o[ FieldId<oid>{} ];
    }
    template<class Tp, size_t... I> inline void assign_from_std( const Tp& t , std::index_sequence<I...> )
    {
      ( ... , ( m_values.template get_nth<I>() = std::get<I>(t) ) );
    }
  };

  template<typename... ids>
  ONIKA_HOST_DEVICE_FUNC inline FieldTuple<ids...> make_field_tuple( const FieldId<ids>& ... , const typename FieldId<ids>::value_type & ... v ) { return FieldTuple<ids...>( v... ); }

} }
