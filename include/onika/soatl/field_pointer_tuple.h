// onika/soatl/field_pointer_tuple.h -- a bundle of column pointers (one per field), trivially copyable into kernels.
#pragma once
#include <onika/cuda/cuda.h>
#include <onika/flat_tuple.h>
#include <onika/soatl/field_id.h>

namespace onika { namespace soatl {

  template<size_t A, size_t C, typename... ids>
  struct FieldPointerTuple
  {
    static constexpr size_t Alignment = A;
    static constexpr size_t ChunkSize = C;
    static constexpr size_t TupleSize = sizeof...(ids);
    FlatTuple< typename FieldId<ids>::value_type * ... > m_pointers;

    FieldPointerTuple() = default;
    ONIKA_HOST_DEVICE_FUNC inline FieldPointerTuple( typename FieldId<ids>::value_type * ... p ) requires ( sizeof...(ids) > 0 ) : m_pointers( p... ) {}

    template<class id> ONIKA_HOST_DEVICE_FUNC inline typename FieldId<id>::value_type * __restrict__ operator [] ( FieldId<id> ) const
    {
      static_assert( find_index_of_id_v<id,ids...> != bad_field_index , "field not in pointer tuple" );
      return m_pointers.template get_nth< find_index_of_id_v<id,ids...> >();
    }
    template<class id> ONIKA_HOST_DEVICE_FUNC inline void set_pointer( FieldId<id> , typename FieldId<id>::value_type * p )
    {
      if constexpr ( find_index_of_id_v<id,ids...> != bad_field_index ) m_pointers.template get_nth< find_index_of_id_v<id,ids...> >() = p;
    }
    template<class id> static constexpr bool has_field( FieldId<id> ) { return find_index_of_id_v<id,ids...> != bad_field_index; }
    static constexpr size_t alignment() { return A; }
    static constexpr size_t chunksize() { return C; }
  };

} }
