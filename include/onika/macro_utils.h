// onika/macro_utils.h -- small preprocessor helpers used by the ONIKA_CU_* macro layer (B200-native front end).
#pragma once

// expands to ", args" when args are given, to nothing otherwise
#define OPT_COMMA_VA_ARGS(...) __VA_OPT__(,) __VA_ARGS__

// marks variables as used without evaluating anything at run time
namespace onika { template<class... T> constexpr inline void fake_use_of_variables(const T&...) {} }
#define FAKE_USE_OF_VARIABLES(...) ::onika::fake_use_of_variables(__VA_ARGS__);

#define ONIKA_STRINGIFY_(x) #x
#define ONIKA_STRINGIFY(x) ONIKA_STRINGIFY_(x)
#define ONIKA_CONCAT_(a,b) a##b
#define ONIKA_CONCAT(a,b) ONIKA_CONCAT_(a,b)
#define ONIKA_STR(x) ONIKA_STRINGIFY_(x)
#define GET_FIRST_ARG(XXX,...) XXX

// ---- list expansion (the reference's EXPAND_WITH_* family, include/onika/macro_utils.h:23-103 there; used by the field
// combiner declaration macros).  First argument = a prefix token sequence or a function-like macro, the others = the list:
//   EXPAND_WITH_PREFIX(p,a,b)              ->   p a , p b            EXPAND_WITH_PREFIX_PREPEND_COMMA(p,a,b)  ->  , p a , p b
//   EXPAND_WITH_FUNC(f,a,b)                ->   f(a) , f(b)          EXPAND_WITH_FUNC_PREPEND_COMMA(f,a,b)    ->  , f(a) , f(b)
//   EXPAND_WITH_FUNC_NOSEP(f,a,b)          ->   f(a) f(b)            an empty list expands to nothing
// Built on recursive __VA_OPT__ re-scanning instead of one macro per list length: lists of up to 256 elements.
#define _ONIKA_PP_PARENS ()
#define _ONIKA_PP_RESCAN(...)  _ONIKA_PP_RESCAN3(_ONIKA_PP_RESCAN3(_ONIKA_PP_RESCAN3(_ONIKA_PP_RESCAN3(__VA_ARGS__))))
#define _ONIKA_PP_RESCAN3(...) _ONIKA_PP_RESCAN2(_ONIKA_PP_RESCAN2(_ONIKA_PP_RESCAN2(_ONIKA_PP_RESCAN2(__VA_ARGS__))))
#define _ONIKA_PP_RESCAN2(...) _ONIKA_PP_RESCAN1(_ONIKA_PP_RESCAN1(_ONIKA_PP_RESCAN1(_ONIKA_PP_RESCAN1(__VA_ARGS__))))
#define _ONIKA_PP_RESCAN1(...) __VA_ARGS__

#define _ONIKA_PP_FUNC_COMMA(f,a,...)     f(a) __VA_OPT__( , _ONIKA_PP_FUNC_COMMA_NEXT _ONIKA_PP_PARENS (f,__VA_ARGS__) )
#define _ONIKA_PP_FUNC_COMMA_NEXT()       _ONIKA_PP_FUNC_COMMA
#define _ONIKA_PP_FUNC_NOSEP(f,a,...)     f(a) __VA_OPT__( _ONIKA_PP_FUNC_NOSEP_NEXT _ONIKA_PP_PARENS (f,__VA_ARGS__) )
#define _ONIKA_PP_FUNC_NOSEP_NEXT()       _ONIKA_PP_FUNC_NOSEP
#define _ONIKA_PP_PREFIX_COMMA(p,a,...)   p a __VA_OPT__( , _ONIKA_PP_PREFIX_COMMA_NEXT _ONIKA_PP_PARENS (p,__VA_ARGS__) )
#define _ONIKA_PP_PREFIX_COMMA_NEXT()     _ONIKA_PP_PREFIX_COMMA

// (one level of indirection: callers write EXPAND_WITH_FUNC( f OPT_COMMA_VA_ARGS(list) ), whose single argument only falls
// apart into `f , list...` once it has been expanded)
#define _ONIKA_PP_EWF(f,...)      __VA_OPT__( _ONIKA_PP_RESCAN( _ONIKA_PP_FUNC_COMMA(f,__VA_ARGS__) ) )
#define _ONIKA_PP_EWF_PC(f,...)   __VA_OPT__( , _ONIKA_PP_RESCAN( _ONIKA_PP_FUNC_COMMA(f,__VA_ARGS__) ) )
#define _ONIKA_PP_EWF_NS(f,...)   __VA_OPT__( _ONIKA_PP_RESCAN( _ONIKA_PP_FUNC_NOSEP(f,__VA_ARGS__) ) )
#define _ONIKA_PP_EWP(p,...)      __VA_OPT__( _ONIKA_PP_RESCAN( _ONIKA_PP_PREFIX_COMMA(p,__VA_ARGS__) ) )
#define _ONIKA_PP_EWP_PC(p,...)   __VA_OPT__( , _ONIKA_PP_RESCAN( _ONIKA_PP_PREFIX_COMMA(p,__VA_ARGS__) ) )
#define EXPAND_WITH_FUNC(...)                  _ONIKA_PP_EWF(__VA_ARGS__)
#define EXPAND_WITH_FUNC_PREPEND_COMMA(...)    _ONIKA_PP_EWF_PC(__VA_ARGS__)
#define EXPAND_WITH_FUNC_NOSEP(...)            _ONIKA_PP_EWF_NS(__VA_ARGS__)
#define EXPAND_WITH_PREFIX(...)                _ONIKA_PP_EWP(__VA_ARGS__)
#define EXPAND_WITH_PREFIX_PREPEND_COMMA(...)  _ONIKA_PP_EWP_PC(__VA_ARGS__)
