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
