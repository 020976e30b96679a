// onika/soatl/field_pointer_tuple.h -- a bundle of column pointers (one per field) plus an element count, trivially
// copyable into kernels: the "view" of a field-array container (make_field_arrays_view) or of externally owned columns.
#pragma once
#include <onika/cuda/cuda.h>
#include <onika/flat_tuple.h>
#include <onika/soatl/constants.h>
#include <onika/soatl/field_id.h>

namespace onika { namespace soatl {

  template<size_t A, size_t C, typename... ids>
  struct FieldPointerTuple
  {
    static_assert( A > 0 && IsPowerOf2<A>::value , "alignment must be a power of 2" );
    static_assert( C > 0 , "chunk size must be strictly positive" );
    static constexpr size_t Alignment = A;
    static constexpr size_t ChunkSize = C;
    static constexpr size_t TupleSize = sizeof...(ids);
    using FieldIdsTuple = FlatTuple< FieldId<ids> ... >;
    using TupleType = FlatTuple< typename FieldId<ids>::value_type * ... >;
    template<class id> using HasField = std::integral_constant< bool , find_index_of_id_v<id,ids...> != bad_field_index >;
    template<class id> static constexpr bool has_field( FieldId<id> ) { return HasField<id>::value; }

    TupleType m_pointers;                 // all null by default
    size_t m_size = 0;

    ONIKA_HOST_DEVICE_FUNC inline FieldPointerTuple() { zero(); }
    FieldPointerTuple( const FieldPointerTuple & ) = default;
    FieldPointerTuple & operator = ( const FieldPointerTuple & ) = default;
    ONIKA_HOST_DEVICE_FUNC inline FieldPointerTuple( typename FieldId<ids>::value_type * ... p ) requires ( sizeof...(ids) > 0 ) : m_pointers( p... ) {}
    ONIKA_HOST_DEVICE_FUNC inline FieldPointerTuple( size_t n , typename FieldId<ids>::value_type * ... p ) requires ( sizeof...(ids) > 0 ) : m_pointers( p... ) , m_size( n ) {}
    ONIKA_HOST_DEVICE_FUNC inline FieldPointerTuple( const TupleType & t ) : m_pointers( t ) {}
    // from a bundle over another field set: common columns are shared, the others null
    template<typename... other> requires ( ! std::is_same_v< FieldPointerTuple<A,C,other...> , FieldPointerTuple > )
    ONIKA_HOST_DEVICE_FUNC inline FieldPointerTuple( const FieldPointerTuple<A,C,other...> & o ) : m_size( o.size() ) { copy_or_zero_fields( o ); }

    ONIKA_HOST_DEVICE_FUNC inline const TupleType & tuple() const { return m_pointers; }
    ONIKA_HOST_DEVICE_FUNC inline TupleType & tuple() { return m_pointers; }
    ONIKA_HOST_DEVICE_FUNC inline size_t size() const { return m_size; }
    ONIKA_HOST_DEVICE_FUNC static constexpr size_t alignment() { return A; }
    ONIKA_HOST_DEVICE_FUNC static constexpr size_t chunksize() { return C; }

    template<class id> ONIKA_HOST_DEVICE_FUNC inline typename FieldId<id>::value_type * __restrict__ operator [] ( FieldId<id> ) const
    {
      static_assert( HasField<id>::value , "field not in pointer tuple" );
      using T = typename FieldId<id>::value_type;
      static constexpr size_t I = find_index_of_id_v<id,ids...>;
      T * p = m_pointers.template get_nth<I>();
      return (T*) ONIKA_BUILTIN_ASSUME_ALIGNED( p , A );
    }
    template<class id> ONIKA_HOST_DEVICE_FUNC inline void set_pointer( FieldId<id> , typename FieldId<id>::value_type * p )
    {
      if constexpr ( HasField<id>::value ) m_pointers.template get_nth< find_index_of_id_v<id,ids...> >() = p;
    }

    ONIKA_HOST_DEVICE_FUNC inline void zero() { ( ... , set_pointer( FieldId<ids>{} , nullptr ) ); }
    template<typename... other> ONIKA_HOST_DEVICE_FUNC inline void copy_existing_fields( const FieldPointerTuple<A,C,other...> & o ) { ( ... , take<ids,false>( o ) ); }
    template<typename... other> ONIKA_HOST_DEVICE_FUNC inline void copy_or_zero_fields( const FieldPointerTuple<A,C,other...> & o ) { ( ... , take<ids,true>( o ) ); }

    ONIKA_HOST_DEVICE_FUNC inline bool operator == ( const FieldPointerTuple & o ) const { return ( ... && ( (*this)[ FieldId<ids>{} ] == o[ FieldId<ids>{} ] ) ); }
    ONIKA_HOST_DEVICE_FUNC inline bool operator != ( const FieldPointerTuple & o ) const { return ! ( *this == o ); }

  private:
    template<class id, bool NullIfMissing, typename... other> ONIKA_HOST_DEVICE_FUNC inline void take( const FieldPointerTuple<A,C,other...> & o )
    {
      if constexpr ( find_index_of_id_v<id,other...> != bad_field_index ) set_pointer( FieldId<id>{} , o[ FieldId<id>{} ] );
      else if constexpr ( NullIfMissing ) set_pointer( FieldId<id>{} , nullptr );
    }
  };

  // a bundle without fields holds nothing
  template<size_t A, size_t C> struct FieldPointerTuple<A,C> { static constexpr size_t TupleSize = 0; };

  // null bundle over the given fields
  template<size_t A, size_t C, typename... ids>
  inline FieldPointerTuple<A,C,ids...> make_field_pointer_tuple( cst::align<A> , cst::chunk<C> , FieldId<ids>... ) { return FieldPointerTuple<A,C,ids...>(); }

  // the given columns of a field-array container as a bundle (borrows the container's storage)
  template<class ArrayT, typename... ids>
  inline FieldPointerTuple<ArrayT::Alignment,ArrayT::ChunkSize,ids...> make_field_arrays_view( const ArrayT & a , FieldId<ids>... )
  {
    return FieldPointerTuple<ArrayT::Alignment,ArrayT::ChunkSize,ids...>( a.size() , a[ FieldId<ids>{} ] ... );
  }

} }
