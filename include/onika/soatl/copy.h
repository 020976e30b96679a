// onika/soatl/copy.h -- soatl::copy(dst, src [, first, n] [, fields...]): column-wise copy between two containers that
// may have different alignment / chunk parameters (e.g. <64,8,...> -> the <1,1,...> wire layout).  One memcpy per
// column on host-visible storage, as in the reference; the device-side equivalent is onk_soatl_repack.
#pragma once
#include <cassert>
#include <cstring>
#include <onika/soatl/field_id.h>

namespace onika { namespace soatl {

  template<class DstT, class SrcT, class... ids>
  static inline void copy( DstT& dst , const SrcT& src , size_t first , size_t n , const FieldId<ids>& ... )
  {
    assert( first + n <= dst.size() && first + n <= src.size() );
    ( ... , std::memcpy( dst[FieldId<ids>{}] + first , src[FieldId<ids>{}] + first , n * sizeof( typename FieldId<ids>::value_type ) ) );
  }

  namespace copy_details
  {
    template<class DstT, class SrcT, class... ids>
    static inline void copy_all( DstT& dst , const SrcT& src , size_t first , size_t n , std::tuple< FieldId<ids> ... > ) { copy( dst , src , first , n , FieldId<ids>{} ... ); }
  }

  // every field of the destination
  template<class DstT, class SrcT>
  static inline void copy( DstT& dst , const SrcT& src , size_t first , size_t n ) { copy_details::copy_all( dst , src , first , n , typename DstT::FieldIdsTuple{} ); }
  template<class DstT, class SrcT>
  static inline void copy( DstT& dst , const SrcT& src ) { assert( dst.size() == src.size() ); copy( dst , src , 0 , src.size() ); }
  // whole range, selected fields
  template<class DstT, class SrcT, class id1, class... ids>
  static inline void copy( DstT& dst , const SrcT& src , const FieldId<id1>& f1 , const FieldId<ids>& ... f ) { assert( dst.size() == src.size() ); copy( dst , src , 0 , src.size() , f1 , f ... ); }

} }
