// onika/soatl/field_combiner.h -- a derived field computed on the fly from stored fields: arrays[combiner][i].
#pragma once
#include <onika/cuda/cuda.h>
#include <onika/soatl/field_id.h>

namespace onika { namespace soatl {
  template<class FuncT, class... fids>
  struct FieldCombiner
  {
    FuncT m_func;
    using value_type = decltype( m_func( typename FieldId<fids>::value_type{} ... ) );
    static const char* name() { return "combiner"; }
  };
  template<class FuncT, class... fids>
  inline FieldCombiner<FuncT,fids...> make_field_combiner( const FuncT& f , const FieldId<fids>& ... ) { return { f }; }
} }
