// onika/lambda_tools.h -- compile-time detection of the call signatures a functor supports.
// block_parallel_for uses it to find prolog/epilog/single-task overloads and the element dimensionality.
#pragma once
#include <type_traits>
#include <utility>

namespace onika
{
  namespace lambda_details
  {
    template<class F, class R, class = void, class... A> struct ReturnsExactly : std::false_type {};
    template<class F, class R, class... A>
    struct ReturnsExactly<F,R,std::void_t<decltype(std::declval<F>()(std::declval<A>()...))>,A...>
      : std::bool_constant< std::is_same_v<decltype(std::declval<F>()(std::declval<A>()...)),R> > {};
    template<class F, class = void, class... A> struct Callable : std::false_type {};
    template<class F, class... A>
    struct Callable<F,std::void_t<decltype(std::declval<F>()(std::declval<A>()...))>,A...> : std::true_type {};
  }

  // true when F can be called with Args... and returns exactly R
  template<class F, class R, class... Args>
  static inline constexpr bool lambda_is_compatible_with_v = lambda_details::ReturnsExactly<F,R,void,Args...>::value;

  // true when F can be called with Args..., whatever it returns
  template<class F, class... Args>
  static inline constexpr bool lambda_is_callable_with_args_v = lambda_details::Callable<F,void,Args...>::value;
}
