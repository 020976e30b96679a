// onika/cuda/cuda_math.h -- small math helpers usable on host and device.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <onika/cuda/cuda.h>

#if defined(__CUDA_ARCH__)
#  define onika_fast_sincosf(t,s,c) __sincosf(t,s,c)
#else
#  define onika_fast_sincosf(t,s,c) do{ *(s) = std::sin(t); *(c) = std::cos(t); }while(false)
#endif

namespace onika { namespace cuda {

  template<class T> ONIKA_HOST_DEVICE_FUNC inline const T& max( const T& a, const T& b ) { return a > b ? a : b; }
  template<class T> ONIKA_HOST_DEVICE_FUNC inline const T& min( const T& a, const T& b ) { return a < b ? a : b; }
  template<class T> ONIKA_HOST_DEVICE_FUNC inline const T& clamp( const T& x, const T& lo , const T& hi ) { return x < lo ? lo : ( x > hi ? hi : x ); }

  // limits as static constexpr members (readable in device code without calling host constexpr functions)
  template<class T> struct numeric_limits;
  template<> struct numeric_limits<uint16_t> { static inline constexpr uint16_t max = 0xFFFFu; };
  template<> struct numeric_limits<uint32_t> { static inline constexpr uint32_t max = 0xFFFFFFFFu; };
  template<> struct numeric_limits<uint64_t> { static inline constexpr uint64_t max = 0xFFFFFFFFFFFFFFFFull; };
  template<> struct numeric_limits<float>  { static inline constexpr float infinity = std::numeric_limits<float>::infinity();  static inline constexpr float quiet_NaN = std::numeric_limits<float>::quiet_NaN(); };
  template<> struct numeric_limits<double> { static inline constexpr double infinity = std::numeric_limits<double>::infinity(); static inline constexpr double quiet_NaN = std::numeric_limits<double>::quiet_NaN(); };

  // x^p for a non-negative integer p by repeated squaring (constexpr, no libm)
  ONIKA_HOST_DEVICE_FUNC inline constexpr double pow_ce_uint( double x , unsigned int p )
  {
    double r = 1.0;
    while( p != 0 ) { if( p & 1u ) r *= x; x *= x; p >>= 1; }
    return r;
  }
  // x^p for an integral-valued double p (negative allowed); NaN when p is not integral
  ONIKA_HOST_DEVICE_FUNC inline constexpr double pow_ce_dint( double x , double p )
  {
    if( p < 0.0 ) return 1.0 / pow_ce_dint( x , -p );
    const unsigned int ip = static_cast<unsigned int>( p );
    if( p != double(ip) ) return numeric_limits<double>::quiet_NaN;
    return pow_ce_uint( x , ip );
  }

} }
