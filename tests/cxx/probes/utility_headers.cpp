// one program, two header trees (tests/test_frontend_cpu.py): call-signature detection, integral constants, flat tuples, autorun / unique-name macros
#include <onika/memory/allocator.h>
#include <onika/lambda_tools.h>
#include <onika/integral_constant.h>
#include <onika/flat_tuple.h>
#include <onika/cpp_utils.h>
#include <cstdio>
using namespace onika;
struct F { int operator () (double) const { return 1; } void operator () (int,int) const {} };
static int g_ran = 0;
ONIKA_AUTORUN_INIT(probe4) { g_ran = 7; }
int main()
{
  std::printf("%d %d %d %d | %d %d %d\n", int(lambda_is_compatible_with_v<F,int,double>), int(lambda_is_compatible_with_v<F,void,double>), int(lambda_is_compatible_with_v<F,void,int,int>), int(lambda_is_callable_with_args_v<F,const char*>),
    int(FunctorSupportsCallSignature<int,F,FlatTuple<double>>::value), int(FunctorSupportsCallSignature<void,F,FlatTuple<double>>::value), int(FunctorSupportsCallArgs<F,FlatTuple<int,int>>::value));
  constexpr auto c = IntConst<6>{} * IntConst<7>{} - IntConst<2>{};
  std::printf("%d %d %d %d %d\n", int(c), int(c == 40), int(c != 40), int(GetAsConstTypeOrValue<IntConst<3>,false>::value), int(GetAsConstTypeOrValue<IntConst<3>,true>::value));
  auto t = make_flat_tuple( 1 , 2.5 , 'x' , "str" );
  auto s = flat_tuple_subset( t , std::index_sequence<2,0>{} );
  std::printf("%zu %zu %d %g %c %s | %c %d | %g %d\n", tuple_size_const_v<decltype(t)>, decltype(s)::size(), t.get(tuple_index<0>), t.get_nth_const<1>(), t.get_copy(tuple_index_t<2>{}), t.get_nth<3>(),
              s.get(tuple_index<0>), s.get_nth<1>(), t.get_copy(tuple_index<1>), int(std::is_same_v<flat_tuple_element_t<decltype(t),3>,const char*>));
  std::printf("%d %s %d\n", g_ran, USTAMP_STR(USTAMP_CONCAT(ab,cd)), MAKE_UNIQUE_NAME(1,0,2,3));
  return 0;
}
