// onika/soatl/field_combiner.h -- a derived field computed on the fly from stored fields: arrays[combiner][i].
//
// FieldCombiner<FuncT,fids...> holds the function object; value_type is what it returns for the fields' value types.  The
// reference only declares the primary template and defines combiners through the two declaration macros below (same macro
// names, arguments and generated members: include/onika/soatl/field_combiner.h:43-68 there); here the primary template is
// usable as it is as well (make_field_combiner), the macros specialise it to give a combiner its own name.
#pragma once
#include <onika/cuda/cuda.h>
#include <onika/macro_utils.h>
#include <onika/soatl/field_id.h>

namespace onika { namespace soatl {
  template<class FuncT, class... fids>
  struct FieldCombiner
  {
    FuncT m_func;
    using value_type = decltype( m_func( typename FieldId<fids>::value_type{} ... ) );
    static const char* short_name() { return "combiner"; }
    static const char* name() { return "combiner"; }
  };
  template<class FuncT, class... fids>
  inline FieldCombiner<FuncT,fids...> make_field_combiner( const FuncT& f , const FieldId<fids>& ... ) { return { f }; }
} }

#define _ONIKA_GET_TYPE_FROM_FIELD_ID(x) typename FieldId<x>::value_type {}

// ns::CombT = a combiner of the listed field tags through FuncT, named `combiner`
#define ONIKA_DECLARE_FIELD_COMBINER(ns,CombT,combiner,FuncT,...) \
namespace onika { namespace soatl { \
  template<> struct FieldCombiner< FuncT OPT_COMMA_VA_ARGS(__VA_ARGS__) > \
  { \
    FuncT m_func; \
    using value_type = decltype( m_func( EXPAND_WITH_FUNC(_ONIKA_GET_TYPE_FROM_FIELD_ID OPT_COMMA_VA_ARGS(__VA_ARGS__)) ) ); \
    static const char* short_name() { return #combiner; } \
    static const char* name() { return #combiner; } \
  }; \
} } \
namespace ns { using CombT = onika::soatl::FieldCombiner< FuncT OPT_COMMA_VA_ARGS(__VA_ARGS__) >; }

// the same with names asked from the function object at run time (m_func.short_name() / m_func.name())
#define ONIKA_DECLARE_DYNAMIC_FIELD_COMBINER(ns,CombT,FuncT,...) \
namespace onika { namespace soatl { \
  template<> struct FieldCombiner< FuncT OPT_COMMA_VA_ARGS(__VA_ARGS__) > \
  { \
    FuncT m_func; \
    using value_type = decltype( m_func( EXPAND_WITH_FUNC(_ONIKA_GET_TYPE_FROM_FIELD_ID OPT_COMMA_VA_ARGS(__VA_ARGS__)) ) ); \
    const char* short_name() const { return m_func.short_name(); } \
    const char* name() const { return m_func.name(); } \
  }; \
} } \
namespace ns { using CombT = onika::soatl::FieldCombiner< FuncT OPT_COMMA_VA_ARGS(__VA_ARGS__) >; }
