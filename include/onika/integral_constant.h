// onika/integral_constant.h -- compile-time integer tags usable in host and device code.
#pragma once
#include <onika/cuda/cuda.h>

namespace onika
{
  template<class T, T V>
  struct IntegralConst
  {
    using value_type = T;
    static constexpr T value = V;
    ONIKA_HOST_DEVICE_FUNC constexpr operator T () const { return V; }
    ONIKA_HOST_DEVICE_FUNC constexpr T operator () () const { return V; }
    ONIKA_HOST_DEVICE_FUNC constexpr bool operator == ( T other ) const { return V == other; }
    ONIKA_HOST_DEVICE_FUNC constexpr bool operator != ( T other ) const { return V != other; }
    template<class U, U W> ONIKA_HOST_DEVICE_FUNC constexpr auto operator + (IntegralConst<U,W>) const { return IntegralConst<decltype(V+W),(V+W)>{}; }
    template<class U, U W> ONIKA_HOST_DEVICE_FUNC constexpr auto operator - (IntegralConst<U,W>) const { return IntegralConst<decltype(V-W),(V-W)>{}; }
    template<class U, U W> ONIKA_HOST_DEVICE_FUNC constexpr auto operator * (IntegralConst<U,W>) const { return IntegralConst<decltype(V*W),(V*W)>{}; }
    template<class U, U W> ONIKA_HOST_DEVICE_FUNC constexpr auto operator / (IntegralConst<U,W>) const { return IntegralConst<decltype(V/W),(V/W)>{}; }
  };

  template<bool B>         using BoolConst = IntegralConst<bool,B>;
  template<int I>          using IntConst  = IntegralConst<int,I>;
  template<unsigned int I> using UIntConst = IntegralConst<unsigned int,I>;
  using TrueType  = BoolConst<true>;
  using FalseType = BoolConst<false>;

  // ::value is the constant's tag object (true) or its plain value (false): lets one template serve both call styles
  template<class IntegralConstT, bool ConstTypeOrValue> struct GetAsConstTypeOrValue { static constexpr IntegralConstT value = {}; };
  template<class IntegralConstT> struct GetAsConstTypeOrValue<IntegralConstT,false> { static constexpr typename IntegralConstT::value_type value = IntegralConstT::value; };
}
