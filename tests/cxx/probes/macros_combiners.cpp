// one program, two header trees: compiled against this repo's include/ and against the reference's own include/ it must print the same
// lines (tests/test_frontend_cpu.py).  Preprocessor list helpers, field combiner declaration macros, FieldIdsIncludingSequence.
#include <onika/macro_utils.h>
#include <onika/soatl/field_id.h>
#include <onika/soatl/field_combiner.h>
#include <cstdio>
#include <cstring>
#include <type_traits>
struct t_rx {}; struct t_ry {}; struct t_m {}; struct t_q {};
namespace onika { namespace soatl {
  template<> struct FieldId<t_rx> { using value_type = double; using Id = t_rx; static const char* name() { return "rx"; } };
  template<> struct FieldId<t_ry> { using value_type = double; using Id = t_ry; static const char* name() { return "ry"; } };
  template<> struct FieldId<t_m>  { using value_type = float;  using Id = t_m;  static const char* name() { return "m"; } };
  template<> struct FieldId<t_q>  { using value_type = int;    using Id = t_q;  static const char* name() { return "q"; } };
} }
struct NormFunc { inline double operator () (double x, double y) const { return x*x + y*y; } };
struct ChargeFunc { const char* short_name() const { return "qm"; } const char* name() const { return "charge_by_mass"; } inline float operator () (int q, float m) const { return q / m; } };
struct ConstFunc { inline int operator () () const { return 42; } };
ONIKA_DECLARE_FIELD_COMBINER( probe , Norm2Combiner , norm2 , NormFunc , t_rx , t_ry )
ONIKA_DECLARE_DYNAMIC_FIELD_COMBINER( probe , ChargeCombiner , ChargeFunc , t_q , t_m )
ONIKA_DECLARE_FIELD_COMBINER( probe , ConstCombiner , fortytwo , ConstFunc )
#define SQ(x) ((x)*(x))
#define DECL(x) int x = 1;
static int sum() { return 0; }
template<class... T> static int sum(int a, T... b) { return a + sum(b...); }
int main()
{
  using namespace onika::soatl;
  const int a[] = { EXPAND_WITH_PREFIX(10+,1,2,3) , -1 EXPAND_WITH_PREFIX_PREPEND_COMMA(100+,4,5) EXPAND_WITH_PREFIX_PREPEND_COMMA(7+) };
  for(int x : a) std::printf("%d ", x);
  std::printf("| %d %d %d %d\n", sum( EXPAND_WITH_FUNC(SQ,1,2,3,4,5,6,7,8,9,10,11,12) ), sum( 5 EXPAND_WITH_FUNC_PREPEND_COMMA(SQ,2,3) ), sum( EXPAND_WITH_FUNC(SQ) ), sum( 1 EXPAND_WITH_FUNC_PREPEND_COMMA(SQ) ));
  EXPAND_WITH_FUNC_NOSEP(DECL,u,v,w) EXPAND_WITH_FUNC_NOSEP(DECL)
  std::printf("%d %s %d\n", u+v+w, ONIKA_STR(GET_FIRST_ARG(alpha,beta,gamma)), ONIKA_CONCAT(1,2));
  probe::Norm2Combiner n2 = { NormFunc{} }; probe::ChargeCombiner qm = { ChargeFunc{} }; probe::ConstCombiner c = {};
  std::printf("%s %s %g %d | %s %s %g %d | %s %d %d\n", n2.short_name(), probe::Norm2Combiner::name(), n2.m_func(3.0,4.0), int(std::is_same_v<probe::Norm2Combiner::value_type,double>),
              qm.short_name(), qm.name(), double(qm.m_func(3,2.0f)), int(std::is_same_v<probe::ChargeCombiner::value_type,float>), c.name(), c.m_func(), int(std::is_same_v<probe::ConstCombiner::value_type,int>));
  using Ref = FieldIds<t_rx,t_ry,t_m,t_q>;
  std::printf("%zu %zu %zu %zu %d %d\n", FieldIdsIncludingSequence<Ref,FieldIds<t_ry>>::n_fields(), FieldIdsIncludingSequence<Ref,FieldIds<t_q,t_rx>>::n_fields(),
              FieldIdsIncludingSequence<Ref,FieldIds<>>::n_fields(), FieldIdsIncludingSequence<Ref,FieldIds<t_m,t_ry>>::n_fields(),
              int(std::is_same_v< FieldIdsIncludingSequence<Ref,FieldIds<t_ry>>::type , FieldIds<t_rx,t_ry> >), int(std::is_same_v< FieldIdsIncludingSequence<Ref,FieldIds<t_m>>::type , FieldIds<t_rx,t_ry,t_m> >));
  return 0;
}
