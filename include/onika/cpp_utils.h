// onika/cpp_utils.h -- ONIKA_AUTORUN_INIT(name) { ... } : code run when the plugin's shared object is loaded.
#pragma once
#include <onika/macro_utils.h>
#ifndef ONIKA_CURRENT_PACKAGE_NAME
#  if defined(ONIKA_PACKAGE_OVERRIDE)
#    define ONIKA_CURRENT_PACKAGE_NAME ONIKA_PACKAGE_OVERRIDE
#  elif defined(ONIKA_PACKAGE_NAME)
#    define ONIKA_CURRENT_PACKAGE_NAME ONIKA_PACKAGE_NAME
#  else
#    define ONIKA_CURRENT_PACKAGE_NAME onika
#  endif
#endif
// token helpers under the reference's names (include/onika/cpp_utils.h:23-34 there): plugins use them for their own unique names
#define USTAMP_CONCAT(a,b) ONIKA_CONCAT(a,b)
#define USTAMP_STR(x) ONIKA_STRINGIFY(x)
#define MAKE_UNIQUE_NAME(base,sep,I,J) ONIKA_UNIQUE_NAME_(base,sep,I,J)
#define ONIKA_UNIQUE_NAME_(base,sep,I,J) base##sep##I##sep##J
#define CONSTRUCTOR_ATTRIB __attribute__((constructor))
#define ONIKA_AUTORUN_NAME_(base,line,pkg) base##_##line##_##pkg
#define ONIKA_AUTORUN_NAME(base,line,pkg) ONIKA_AUTORUN_NAME_(base,line,pkg)
#define ONIKA_AUTORUN_INIT(name) __attribute__((constructor)) void ONIKA_AUTORUN_NAME(__onika_autorun_##name,__LINE__,ONIKA_CURRENT_PACKAGE_NAME) ()
