// onika/lambda_tools.h -- compile-time detection of the call signatures a functor supports.
// block_parallel_for uses it to find prolog/epilog/single-task overloads and the element dimensionality.
#pragma once
#include <type_traits>
#include <utility>
#include <onika/flat_tuple.h>

namespace onika
{
  namespace lambda_details
  {
    template<class F, class R, class = void, class... A> struct ReturnsExactly : std::false_type {};
    template<class F, class R, class... A>
    struct ReturnsExactly<F,R,std::void_t<decltype(std::declval<F>()(std::declval<A>()...))>,A...>
      : std::bool_constant< std::is_same_v<decltype(std::declval<F>()(std::declval<A>()...)),R> > {};
    template<class F, class = void, class... A> struct Callable : std::false_type {};
    // as in the reference (include/onika/lambda_tools.h:47-48 there: `sizeof(decltype(call)) >= 0`), the call must return an
    // object type: a call that returns void does not count
    template<class F, class... A>
    struct Callable<F,std::enable_if_t< ( sizeof(decltype(std::declval<F>()(std::declval<A>()...))) > 0 ) >,A...> : std::true_type {};
  }

  // true when F can be called with Args... and returns exactly R
  template<class F, class R, class... Args>
  static inline constexpr bool lambda_is_compatible_with_v = lambda_details::ReturnsExactly<F,R,void,Args...>::value;

  // true when F can be called with Args... and returns a value, whatever its type (false for void results, see above)
  template<class F, class... Args>
  static inline constexpr bool lambda_is_callable_with_args_v = lambda_details::Callable<F,void,Args...>::value;

  // the same two tests as class templates over an argument list packed in a FlatTuple (the reference's spelling)
  template<class R, class F, class ArgsTp> struct FunctorSupportsCallSignature : std::false_type {};
  template<class R, class F, class... Args>
  struct FunctorSupportsCallSignature< R , F , FlatTuple<Args...> > : std::bool_constant< lambda_is_compatible_with_v<F,R,Args...> > {};
  template<class F, class ArgsTp> struct FunctorSupportsCallArgs : std::false_type {};
  template<class F, class... Args>
  struct FunctorSupportsCallArgs< F , FlatTuple<Args...> > : std::bool_constant< lambda_is_callable_with_args_v<F,Args...> > {};
}
