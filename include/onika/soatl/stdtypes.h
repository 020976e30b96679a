// onika/soatl/stdtypes.h
#pragma once
#include <cstdint>
#ifndef SOATL_SIZE_TYPE_64BITS
#ifndef SOATL_SIZE_TYPE_32BITS
#define SOATL_SIZE_TYPE_32BITS 1     // the reference's default build (CMakeLists option SOATL_SIZE_TYPE_32BITS=ON)
#endif
#endif
namespace onika { namespace soatl {
# ifdef SOATL_SIZE_TYPE_64BITS
  using array_size_t = uint64_t;
# else
  using array_size_t = uint32_t;
# endif
} }
