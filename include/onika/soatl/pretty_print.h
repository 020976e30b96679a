// onika/soatl/pretty_print.h -- human-readable description of a field-array container (layout parameters and, per
// field, name / value type / column address).
#pragma once
#include <tuple>
#include <typeinfo>
#include <onika/soatl/field_id.h>

namespace onika { namespace soatl {

  namespace details
  {
    template<class StreamT, class ArrayT, class id>
    static inline void pretty_print_column( StreamT & out , const ArrayT & arrays , FieldId<id> f )
    {
      out << "  " << FieldId<id>::name() << " : type=" << typeid( typename FieldId<id>::value_type ).name()
          << ", addr=" << static_cast<const void*>( arrays[f] ) << "\n";
    }
  }

  template<class StreamT, class ArrayT, typename... ids>
  static inline StreamT & pretty_print_arrays( StreamT & out , const ArrayT & arrays , const FieldId<ids> & ... f )
  {
    out << "alignment=" << arrays.alignment() << "\n" << "chunksize=" << arrays.chunksize() << "\n"
        << "size=" << arrays.size() << "\n" << "datasize=" << arrays.storage_size() << "\n"
        << "contains " << sizeof...(ids) << " fields :\n";
    ( ... , details::pretty_print_column( out , arrays , f ) );
    return out;
  }

  template<class StreamT, class ArrayT, typename... ids>
  static inline StreamT & pretty_print_arrays( StreamT & out , const ArrayT & arrays , const std::tuple< FieldId<ids> ... > & )
  {
    return pretty_print_arrays( out , arrays , FieldId<ids>{} ... );
  }

  // every field of the container
  template<class StreamT, class ArrayT>
  static inline StreamT & pretty_print_arrays( StreamT & out , const ArrayT & arrays )
  {
    return pretty_print_arrays( out , arrays , typename ArrayT::FieldIdsTuple{} );
  }

} }
