"""CPU: the boundary compiles where a maintainer would use it -- include/onika_b200.h as plain C11 and as C++, a C
program linked against libonika_b200.so, and the C++ front end (include/onika/**) in a host-only g++ translation unit."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")


def _env():
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    return env


def test_c_abi_header_is_valid_c11_and_cxx(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include <onika_b200.h>\nint main(void){ return onk_abi_version() == ONK_ABI_VERSION ? 0 : 1; }\n')
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-fsyntax-only", f"-I{INC}", str(src)], env=_env())
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", f"-I{INC}", str(src)], env=_env())


def test_plain_c_program_links_and_calls_the_library(tmp_path):
    from onika_b200 import _build
    lib = _build.build()
    src = tmp_path / "link.c"
    src.write_text(r'''
#include <onika_b200.h>
#include <stdio.h>
int main(void)
{
  const uint32_t fb[6] = {8,8,8,1,8,4};
  uint64_t off[6];
  const uint64_t cap = onk_soatl_update_capacity(17, 0, 16);
  const uint64_t total = onk_soatl_layout(64, fb, 6, cap, off);
  printf("%d %llu %llu %llu %llu\n", onk_abi_version(), (unsigned long long)cap, (unsigned long long)total,
         (unsigned long long)off[3], (unsigned long long)off[5]);
  int n = -1;
  const int rc = onk_device_count(&n);            /* no GPU here: a clean error code, not a crash */
  printf("rc=%d err=%s\n", rc, rc ? onk_last_error() : "none");
  return 0;
}
''')
    exe = tmp_path / "link"
    subprocess.check_call(["gcc", "-std=c11", f"-I{INC}", str(src), "-o", str(exe), f"-L{os.path.dirname(lib)}", "-lonika_b200",
                           f"-Wl,-rpath,{os.path.dirname(lib)}"], env=_env())
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    assert out.splitlines()[0] == "2 32 1216 768 1088"          # SURVEY 8a row a23 layout probe


def test_cxx_front_end_compiles_in_a_host_only_translation_unit(tmp_path):
    """plugins also contain host-only .cpp files that include the onika headers (no nvcc): they must parse with g++"""
    src = tmp_path / "host_tu.cpp"
    src.write_text(r'''
#include <onika/parallel/block_parallel_for.h>
#include <onika/parallel/parallel_for.h>
#include <onika/parallel/memset.h>
#include <onika/parallel/parallel_execution_context_allocator.h>
#include <onika/parallel/parallel_execution_queue.h>
#include <onika/parallel/parallel_execution_operators.h>
#include <onika/extras/array2d.h>
#include <onika/extras/block_parallel_value_add_functor.h>
#include <onika/cuda/cuda_reduction.h>
#include <onika/cuda/memcpy.h>
#include <onika/soatl/field_arrays.h>
#include <onika/soatl/compute.h>
#include <onika/soatl/copy.h>
#include <onika/soatl/field_tuple.h>
#include <onika/soatl/field_pointer_tuple.h>
#include <onika/soatl/field_combiner.h>
#include <onika/soatl/pretty_print.h>
#include <onika/soatl/field_id_tuple_utils.h>
#include <onika/parallel/parallel_execution_graph.h>
#include <onika/parallel/reduction.h>
#include <onika/cuda/stl_adaptors.h>
#include <onika/scg/operator.h>
#include <onika/app/api.h>
struct tag_x {}; struct tag_i {};
namespace onika { namespace soatl {
  template<> struct FieldId<tag_x> { using value_type = double; using Id = tag_x; static const char* name() { return "x"; } };
  template<> struct FieldId<tag_i> { using value_type = int; using Id = tag_i; static const char* name() { return "i"; } };
} }
using namespace onika;
static_assert( sizeof(parallel::ParallelDataAccess) == 64 );
static_assert( sizeof(parallel::HostKernelExecutionScratch) == 1024 && sizeof(parallel::GPUKernelExecutionScratch) == 1024 );
static_assert( ! parallel::functor_gpu_support_v< extras::BlockParallelValueAddFunctor<true> > , "host-only TU: no device target" );
static_assert( soatl::find_index_of_id_v<tag_i,tag_x,tag_i> == 1 );
// value tuples are plain host objects: zero by default, field-wise conversion, equality
int tuple_probe()
{
  soatl::FieldTuple<tag_x,tag_i> t;                       // zeroed
  soatl::FieldTuple<tag_i> only_i( 7 );
  t.copy_existing_fields( only_i );
  soatl::FieldTuple<tag_i,tag_x> swapped( t );             // conversion by field, not by position
  const double sorted[4] = { 1.0 , 2.0 , 4.0 , 8.0 };
  return ( t[soatl::FieldId<tag_x>{}] == 0.0 && swapped[soatl::FieldId<tag_i>{}] == 7 && swapped == swapped
           && onika::cuda::lower_bound( sorted , sorted+4 , 3.0 ) == sorted+2 ) ? 0 : 1;
}
int layout_probe()
{
  // pure layout arithmetic of FieldArrays<64,16,x,i>: column offsets for a capacity (host only, no allocation)
  return int( soatl::pfa_pointer_offset<64,16,tag_i,tag_x,tag_i>( 32 ) );    // 32*8 = 256
}
''')
    subprocess.check_call(["g++", "-std=c++20", "-fopenmp", "-Wall", "-c", f"-I{INC}", str(src), "-o", str(tmp_path / "host_tu.o")], env=_env())


def test_host_side_logic_of_the_front_end_runs_without_a_gpu(tmp_path):
    """pure host logic of the C++ front end, executed here: data-access conflict tests and conflict maps, the split of an
    execution space into dependent / independent elements, coordinate-range helpers, value tuples"""
    from onika_b200 import _build
    lib = _build.build()
    src = tmp_path / "host_logic.cpp"
    src.write_text(r'''
#include <onika/parallel/parallel_data_access.h>
#include <onika/parallel/parallel_execution_space.h>
#include <onika/soatl/field_tuple.h>
#include <cstdio>
#include <vector>
using namespace onika; using namespace onika::parallel;
struct tag_a {}; struct tag_b {};
namespace onika { namespace soatl {
  template<> struct FieldId<tag_a> { using value_type = double; using Id = tag_a; static const char* name() { return "a"; } };
  template<> struct FieldId<tag_b> { using value_type = int; using Id = tag_b; static const char* name() { return "b"; } };
} }
int main()
{
  double x[4], y[4];
  ParallelDataAccessVector rw_x, ro_x, rw_y, cross_x;
  rw_x.push_back( local_access( x , 2 , AccessStencilElement::RW , "x" ) );
  ro_x.push_back( local_access( x , 2 , AccessStencilElement::RO , "x" ) );
  rw_y.push_back( local_access( y , 2 , AccessStencilElement::RW , "y" ) );
  cross_x.push_back( access_cross_2d( x , AccessStencilElement::RW , AccessStencilElement::RO , "xc" ) );
  AccessConflictMap cm;
  concurrent_data_access_conflict_map( rw_x , cross_x , cm );
  std::printf("%d %d %d %d\n", int(concurrent_data_access(rw_x,ro_x)), int(concurrent_data_access(rw_x,rw_y)), int(concurrent_data_access(cross_x,rw_x)), int(cm.size()));
  using C2 = oarray_t<ssize_t,2>;
  const ParallelExecutionSpace<2> first = { {0,0} , {4,4} } , second = { {3,0} , {7,2} };
  std::vector<C2> stencil = { C2{0,0} , C2{-1,0} , C2{1,0} , C2{0,-1} , C2{0,1} } , dep , rem;
  concurrent_execution_space_elements( first , second , std::span<C2>( stencil ) , dep , rem );
  std::printf("%zu %zu %zu %zu\n", dep.size(), rem.size(), coord_range_size( C2{0,0} , C2{4,4} ), coord_range_size( ssize_t(3) , ssize_t(1) ));
  soatl::FieldTuple<tag_a,tag_b> t;  t[soatl::FieldId<tag_b>{}] = 5;
  soatl::FieldTuple<tag_b,tag_a> u( t );
  std::printf("%g %d %d\n", u[soatl::FieldId<tag_a>{}], u[soatl::FieldId<tag_b>{}], int( u == u ));
  return 0;
}
''')
    exe = tmp_path / "host_logic"
    subprocess.check_call(["g++", "-std=c++20", "-fopenmp", f"-I{INC}", str(src), "-o", str(exe), f"-L{os.path.dirname(lib)}", "-lonika_b200",
                           f"-Wl,-rpath,{os.path.dirname(lib)}"], env=_env())
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines()
    assert out[0].split()[:3] == ["1", "0", "1"] and int(out[0].split()[3]) == 5     # x vs x conflicts, x vs y does not; RW centre against the 5-point cross
    assert out[1] == "4 4 16 0"
    assert out[2] == "0 5 1"


def test_lane_dag_placement_is_pure_host_logic(tmp_path):
    """pre_process_queue (any_lane + declared accesses -> lanes and cross-lane edges) needs no GPU: run it here on queued,
    never scheduled operations and check the placement the GPU test relies on"""
    from onika_b200 import _build
    lib = _build.build()
    src = tmp_path / "lane_dag.cpp"
    src.write_text(r'''
#include <onika/parallel/block_parallel_for.h>
#include <onika/parallel/parallel_execution_context_allocator.h>
#include <onika/parallel/parallel_execution_queue.h>
#include <onika/parallel/parallel_execution_operators.h>
#include <cstdio>
using namespace onika; using namespace onika::parallel;
struct HostOnly { inline void operator () ( size_t ) const {} };
int main()
{
  onika::cuda::CudaContext::set_global_gpu_enable( false );          // no device in this process
  ParallelExecutionQueue owner;                                       // default queue of the contexts; never flushed
  ParallelExecutionContextAllocator peca;
  auto ctx = [&](const char* tag) { return peca.create( tag , "" , &owner , nullptr ); };
  double A[1], B[1], C[1];
  const auto a_rw = local_access( A , 2 , AccessStencilElement::RW , "a" ), a_ro = local_access( A , 2 , AccessStencilElement::RO , "a" );
  const auto b_rw = local_access( B , 2 , AccessStencilElement::RW , "b" ), c_rw = local_access( C , 2 , AccessStencilElement::RW , "c" );
  ParallelExecutionContext* pec[7] = { ctx("a1") , ctx("b1") , ctx("a2") , ctx("ab") , ctx("a3") , ctx("c1") , ctx("free") };
  ParallelExecutionQueueBase q;
  q << any_lane(4)
    << a_rw << block_parallel_for( 8 , HostOnly{} , pec[0] )
    << b_rw << block_parallel_for( 8 , HostOnly{} , pec[1] )
    << a_rw << block_parallel_for( 8 , HostOnly{} , pec[2] )
    << a_ro << b_rw << block_parallel_for( 8 , HostOnly{} , pec[3] )
    << a_rw << block_parallel_for( 8 , HostOnly{} , pec[4] )
    << c_rw << block_parallel_for( 8 , HostOnly{} , pec[5] )
    << block_parallel_for( 8 , HostOnly{} , pec[6] );                 // no declared access: round robin like the reference
  q.pre_process_queue( q.m_queue_list );
  for(int k=0;k<7;k++)
  {
    int nw = 0; for(int l=0;l<256;l++) if( pec[k]->depends_on_lane(l) ) ++nw;
    std::printf("%d:%d ", int(pec[k]->m_lane), nw);
  }
  std::printf("\n");
  // hand the never-scheduled operations back: destroy the type-erased functor, run the finalize callback (returns the context)
  while( q.m_queue_list != nullptr )
  {
    auto* p = q.m_queue_list; q.m_queue_list = p->m_next; p->m_next = nullptr;
    auto fin = p->m_finalize;
    reinterpret_cast<BlockParallelForHostFunctor*>( p->m_host_scratch.functor_data ) -> ~BlockParallelForHostFunctor();
    if( fin.m_func != nullptr ) ( * fin.m_func )( p , fin.m_data );
  }
  return 0;
}
''')
    exe = tmp_path / "lane_dag"
    subprocess.check_call(["g++", "-std=c++20", "-fopenmp", f"-I{INC}", str(src), "-o", str(exe), f"-L{os.path.dirname(lib)}", "-lonika_b200",
                           f"-Wl,-rpath,{os.path.dirname(lib)}"], env=_env())
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    lanes = [int(x.split(":")[0]) for x in out]
    waits = [int(x.split(":")[1]) for x in out]
    assert lanes[6] == 0                                             # phase 1: the undeclared operation takes the first lane of the cycle
    assert lanes[0] == lanes[2] == lanes[3] == lanes[4]              # the chain on A (reads of A included) stays on one lane
    assert lanes[1] != lanes[0] and lanes[5] not in (lanes[0], lanes[1])
    assert waits == [0, 0, 0, 1, 0, 0, 0]                            # only "B += A" waits for another lane


def test_msp_flat_batch_reader(tmp_path):
    """the .msp subset reader of tests/cxx/run_operator.cu (comments, flow maps, bare operator names, task_group / name keys,
    errors with file:line) on the committed batch files and on malformed input"""
    here = os.path.dirname(os.path.abspath(__file__))
    src = tmp_path / "msp_probe.cpp"
    src.write_text(r'''
#include "msp_reader.h"
int main(int argc, char* argv[])
{
  std::vector<msp::BatchSpec> b;
  if( ! msp::parse_msp( argv[1] , b ) ) return 3;
  for(const auto& s : b)
  {
    std::printf("%s|%s|%d|", s.key.c_str(), s.name.c_str(), int(s.task_group));
    for(const auto& it : s.body) { std::printf("%s(", it.op.c_str()); for(const auto& kv : it.values) std::printf("%s=%s;", kv.first.c_str(), kv.second.c_str()); std::printf(")"); }
    std::printf("\n");
  }
  return 0;
}
''')
    exe = tmp_path / "msp_probe"
    subprocess.check_call(["g++", "-std=c++17", "-Wall", f"-I{os.path.join(here, 'cxx')}", str(src), "-o", str(exe)], env=_env())

    def parse(path):
        r = subprocess.run([str(exe), str(path)], capture_output=True, text=True)
        return r.returncode, r.stdout.splitlines(), r.stderr

    rc, out, _ = parse(os.path.join(here, "cxx", "msp", "tutorial_chain.msp"))
    assert rc == 0 and out == ["tutorial_chain|tutorial_chain_test|1|synchronous_block_parallel(my_value=1.1;)async_block_parallel(my_value=1.2;)"
                               "synchronous_block_parallel(my_value=1.3;)"]
    rc, out, _ = parse(os.path.join(here, "cxx", "msp", "concurrent_lanes.msp"))
    assert rc == 0 and out == ["lanes_then_auto|lane_operators_back_to_back|1|concurrent_block_parallel(rows=4;columns=256;value=2.3;iter_count=2;)"
                               "automatic_concurrent_parallel(iter_count=1;value=0.5;)"]
    ok = tmp_path / "ok.msp"
    ok.write_text("a:\n  body:\n    - nop\n    - op2: { s: \"quoted text\" }   # trailing comment\nb:\n  task_group: false\n  body:\n    - x: { k: 1 }\n")
    rc, out, _ = parse(ok)
    assert rc == 0 and out == ["a|a|0|nop()op2(s=quoted text;)", "b|b|0|x(k=1;)"]
    for bad, where in (("a:\n  body:\n    - op: { k 1 }\n", ":3:"), ("- op\n", ":1:"), ("a:\n  - op\n", ":2:"), ("a:\n  colour: red\n", ":2:")):
        f = tmp_path / "bad.msp"
        f.write_text(bad)
        rc, out, err = parse(f)
        assert rc == 3 and where in err, (bad, err)


def test_oarray_order_and_pfa_size_calculator_match_the_reference_headers(tmp_path):
    """oarray_t::operator<=> (coordinate 0 most significant, include/onika/oarray.h:49-58 in the reference) and the SOATL size
    calculator names (pfa_size_calculator_t, pfa_details::check_size_for_capacity, pfa_storage_size, pfa_pointer_offset:
    include/onika/soatl/packed_field_arrays.h:36-178): ONE program compiled against this repo's headers and, where the
    reference tree is present, against the reference's own headers -- the two outputs must be identical, and the sizes must
    equal the oracle's calculator (pinned against oracle/_ref in test_oracle_vs_reference.py)."""
    import oracle
    src = tmp_path / "probe.cpp"
    src.write_text(r'''
#include <onika/oarray.h>
#include <onika/soatl/field_id.h>
#include <onika/soatl/packed_field_arrays.h>
#include <algorithm>
#include <cstdio>
#include <vector>
using namespace onika;
struct t_rx {}; struct t_at {}; struct t_e {}; struct t_mid {}; struct t_h {};
namespace onika { namespace soatl {
  template<> struct FieldId<t_rx>  { using value_type = double;        using Id = t_rx;  static const char* name() { return "rx"; } };
  template<> struct FieldId<t_at>  { using value_type = unsigned char; using Id = t_at;  static const char* name() { return "at"; } };
  template<> struct FieldId<t_e>   { using value_type = double;        using Id = t_e;   static const char* name() { return "e"; } };
  template<> struct FieldId<t_mid> { using value_type = int;           using Id = t_mid; static const char* name() { return "mid"; } };
  template<> struct FieldId<t_h>   { using value_type = short;         using Id = t_h;   static const char* name() { return "h"; } };
} }
template<size_t A, size_t C> void sizes()
{
  using Fids = soatl::FieldIds<t_rx,t_at,t_e,t_mid,t_h>;
  for(size_t cap : { size_t(0) , C , 2*C , 7*C , 1000*C })
  {
    using IA = std::integral_constant<size_t,A>; using IC = std::integral_constant<size_t,C>;
    std::printf("%zu %zu %zu %zu %zu %zu %zu\n", A, C, cap, soatl::pfa_size_calculator_t<A,Fids,C>::storage_size(cap),
                soatl::pfa_details::check_size_for_capacity( IA{} , IC{} , Fids{} , cap ),
                soatl::pfa_storage_size<A,C,Fids>( cap ), soatl::pfa_pointer_offset<A,C,t_mid,t_rx,t_at,t_e,t_mid,t_h>( cap ) );
  }
}
int main()
{
  using C3 = oarray_t<long,3>;
  std::vector<C3> v = { C3{{2,0,0}} , C3{{0,0,5}} , C3{{0,3,0}} , C3{{1,9,9}} , C3{{0,3,-1}} , C3{{2,0,0}} };
  std::sort( v.begin() , v.end() );
  for(const auto& c : v) std::printf("%ld,%ld,%ld ", c[0], c[1], c[2]);
  std::printf("| %d %d %d\n", int( C3{{0,9,9}} < C3{{1,0,0}} ), int( C3{{1,0,0}} < C3{{0,9,9}} ), int( C3{{1,2,3}} == C3{{1,2,3}} ));
  sizes<64,16>(); sizes<64,8>(); sizes<256,16>(); sizes<1,1>(); sizes<32,4>();
  return 0;
}
''')
    stub = os.path.join(os.path.dirname(INC), "oracle", "stub")

    def build_and_run(include, name):
        exe = tmp_path / name
        subprocess.check_call(["g++", "-std=c++20", "-fopenmp", "-DSOATL_SIZE_TYPE_32BITS=1", f"-I{include}", f"-I{stub}", str(src), "-o", str(exe)]
                              + ([f"-L{os.path.dirname(lib)}", "-lonika_b200", f"-Wl,-rpath,{os.path.dirname(lib)}"] if include == INC else []), env=_env())
        return subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines()

    from onika_b200 import _build
    lib = _build.build()
    ours = build_and_run(INC, "ours")
    assert ours[0] == "0,0,5 0,3,-1 0,3,0 1,9,9 2,0,0 2,0,0 | 1 0 1"          # coordinate 0 first
    sizes = {8: "rx", 1: "at", 4: "mid", 2: "h"}
    for line in ours[1:]:
        a, c, cap, calc, chk, stor, off = map(int, line.split())
        assert calc == chk == stor == oracle.orc.storage_size_calc(a, c, [8, 1, 8, 4, 2], cap), line
        assert off == oracle.orc.storage_size_calc(a, c, [8, 1, 8], cap), line
    ref_inc = "/root/reference/include"
    if os.path.isdir(ref_inc):
        theirs = build_and_run(ref_inc, "theirs")
        assert theirs == ours


def test_host_atomics_and_stl_adaptors_match_the_reference_headers(tmp_path):
    """onika/atomic.h (capture_atomic_add / _sub / _min / _max, inplace_atomic_add / _sub: include/onika/atomic.h:14-62 in the
    reference), the host meaning of the ONIKA_CU_ATOMIC_* macros (cuda.h:200-207 there) and onika/cuda/stl_adaptors.h (vector_data
    / vector_size, lower_bound as a first-not-less scan, span, pair ordering, forward memmove, the printf-backed cout): ONE
    program compiled against this repo's headers and against the reference's own headers prints the same lines."""
    src = tmp_path / "probe2.cpp"
    src.write_text(r'''
#include <onika/atomic.h>
#include <onika/cuda/cuda.h>
#include <onika/cuda/stl_adaptors.h>
#include <onika/type_utils.h>
#include <cstdio>
#include <span>
#include <vector>
int main()
{
  long n = 0, n2 = 0; double s = 0.0, mn = 1e300, mx = -1e300; unsigned int u = 100000u; int lo = 1 << 30, hi = -(1 << 30);
# pragma omp parallel for num_threads(8)
  for(int i=0;i<20000;i++)
  {
    const double v = double( ( i * 7919 ) % 20011 ) - 3000.0;
    onika::capture_atomic_add( n , 1L ); onika::inplace_atomic_add( s , 0.5 ); onika::capture_atomic_sub( u , 2u ); onika::inplace_atomic_sub( n2 , 3L );
    onika::capture_atomic_min( mn , v ); onika::capture_atomic_max( mx , v );
    ONIKA_CU_ATOMIC_ADD( n2 , 5 ); ONIKA_CU_ATOMIC_MIN( lo , int(v) ); ONIKA_CU_ATOMIC_MAX( hi , int(v) ); ONIKA_CU_ATOMIC_SUB( s , 0.25 );
  }
  std::printf("%ld %ld %.3f %.1f %.1f %u %d %d\n", n, n2, s, mn, mx, u, lo, hi);
  double x = 3.0; long k = 7; float f = 2.0f;
  const double a0 = onika::capture_atomic_add( x , 1.5 ), a1 = onika::capture_atomic_sub( x , 0.25 ), a2 = onika::capture_atomic_min( x , 9.0 ), a3 = onika::capture_atomic_min( x , 1.0 ), a4 = onika::capture_atomic_max( x , 0.5 ), a5 = onika::capture_atomic_max( x , 8.0 );
  const long k0 = onika::capture_atomic_add( k , 3L ); const float f0 = onika::capture_atomic_max( f , 5.0f );
  std::printf("%g %g %g %g %g %g -> %g | %ld %ld | %g %g\n", a0, a1, a2, a3, a4, a5, x, k0, k, double(f0), double(f));
  long r = ONIKA_CU_ATOMIC_ADD( k , 10 ); std::printf("%ld %ld\n", r, k);

  std::vector<double> v = { 1.0 , 2.0 , 2.0 , 5.0 , 8.0 };
  const std::vector<double>& cv = v;
  std::printf("%d %d %zu\n", int( onika::cuda::vector_data(v) == v.data() ), int( onika::cuda::vector_data(cv) == v.data() ), onika::cuda::vector_size(cv));
  std::vector<double> empty; std::printf("%zu\n", onika::cuda::vector_size(empty));
  const double unsorted[6] = { 1.0 , 4.0 , 2.0 , 9.0 , 3.0 , 0.5 };
  for(double q : { 0.0 , 2.0 , 2.5 , 8.0 , 9.0 , 100.0 })
    std::printf("%td %td ", onika::cuda::lower_bound( v.data() , v.data()+5 , q ) - v.data() , onika::cuda::lower_bound( unsorted , unsorted+6 , q ) - unsorted );
  std::printf("\n");
  auto sp = onika::cuda::make_span( v ); auto csp = onika::cuda::make_const_span( cv );
  sp[1] = 2.5; double sum = 0; for(double e : csp) sum += e;
  onika::cuda::span<double> agg = { v.data()+1 , 3 };
  std::printf("%zu %zu %d %g %g %g %d %d %d\n", sp.size(), csp.size(), int(sp.empty()), sum, agg[0], *(agg.end()-1), int( onika::is_span_v< onika::cuda::span<double> > ), int( onika::is_span_v< std::span<const long> > ), int( onika::is_span_v< std::vector<double> > ));
  using P = onika::cuda::pair<int,double>;
  const P p1 = {1,2.0} , p2 = {1,3.0} , p3 = {2,0.0} , p4 = {1,2.0};
  std::printf("%d %d %d %d %d %d %d %d\n", int(p1<p2), int(p2<p1), int(p1<p3), int(p3>=p2), int(p1>=p4), int(p1!=p4), int(p1!=p2), int(p1==p4));
  unsigned char buf[40]; for(int i=0;i<40;i++) buf[i] = (unsigned char) i;
  onika::cuda::memmove( buf , buf+8 , 24 ); onika::cuda::memmove( buf+1 , buf+4 , 7 ); onika::cuda::memmove( buf+30 , buf+30 , 5 );
  for(int i=0;i<40;i++) std::printf("%d ", int(buf[i]));
  std::printf("\n");
  unsigned int grid = 0, idx_sum = 0, blk = 0, tid = 0;
# pragma omp parallel num_threads(4) reduction(+:idx_sum) reduction(max:grid,blk,tid)
  { grid = ONIKA_CU_GRID_SIZE; idx_sum += ONIKA_CU_BLOCK_IDX; blk = ONIKA_CU_BLOCK_SIZE; tid = ONIKA_CU_THREAD_IDX; }
  std::printf("%u %u %u %u %u %u\n", grid, idx_sum, blk, tid, (unsigned int) ONIKA_CU_GRID_SIZE, (unsigned int) ONIKA_CU_BLOCK_IDX);
  onika::cuda::cout << "cout " << uint64_t(7) << " " << int32_t(-3) << " " << 2.5 << " " << 1.5f << " " << int16_t(-2) << "\n";
  return 0;
}
''')
    stub = os.path.join(os.path.dirname(INC), "oracle", "stub")

    def build_and_run(include, name):
        exe = tmp_path / name
        subprocess.check_call(["g++", "-std=c++20", "-fopenmp", "-O1", f"-I{include}", f"-I{stub}", str(src), "-o", str(exe)]
                              + ([f"-L{os.path.dirname(lib)}", "-lonika_b200", f"-Wl,-rpath,{os.path.dirname(lib)}"] if include == INC else []), env=_env())
        return subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines()

    from onika_b200 import _build
    lib = _build.build()
    ours = build_and_run(INC, "ours")
    assert ours[0] == "20000 40000 5000.000 -3000.0 17010.0 60000 -3000 17010"
    assert ours[1] == "3 4.5 4.25 4.25 1 1 -> 8 | 7 10 | 2 5"
    assert ours[2] == "10 20"
    assert ours[5] == "0 0 1 1 3 1 4 3 5 3 5 6 "          # sorted: std::lower_bound; unsorted: the first element that is not less
    assert ours[8].split()[:12] == ["8", "12", "13", "14", "15", "16", "17", "18", "16", "17", "18", "19"]
    assert ours[9] == "4 6 1 0 1 0"          # host items: the OpenMP team is the grid, the thread's rank the block index
    assert ours[10] == "cout 7 -3  2.500000000e+00  1.500000000e+00 -2"
    ref_inc = "/root/reference/include"
    if os.path.isdir(ref_inc):
        theirs = build_and_run(ref_inc, "theirs")
        assert theirs == ours


def _probe_on_both_header_trees(tmp_path, name, extra=()):
    """compiles tests/cxx/probes/<name>.cpp with g++ against this repo's headers and, where the reference tree is present,
    against the reference's own headers; returns (our lines, their lines or None)"""
    from onika_b200 import _build
    lib = _build.build()
    src = os.path.join(ROOT, "tests", "cxx", "probes", name + ".cpp")
    stub = os.path.join(ROOT, "oracle", "stub")

    def build_and_run(include, tag):
        exe = tmp_path / (name + "_" + tag)
        subprocess.check_call(["g++", "-std=c++20", "-fopenmp", "-O1", *extra, f"-I{include}", f"-I{stub}", src, "-o", str(exe)]
                              + ([f"-L{os.path.dirname(lib)}", "-lonika_b200", f"-Wl,-rpath,{os.path.dirname(lib)}"] if include == INC else []), env=_env())
        return subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines()

    ours = build_and_run(INC, "ours")
    ref_inc = "/root/reference/include"
    theirs = build_and_run(ref_inc, "theirs") if os.path.isdir(ref_inc) else None
    return ours, theirs


def test_list_macros_field_combiner_declarations_and_including_sequences_match_the_reference_headers(tmp_path):
    """EXPAND_WITH_PREFIX / EXPAND_WITH_FUNC (+ _PREPEND_COMMA, _NOSEP), GET_FIRST_ARG, ONIKA_STR, ONIKA_CONCAT
    (include/onika/macro_utils.h:23-119 in the reference), ONIKA_DECLARE_FIELD_COMBINER / ONIKA_DECLARE_DYNAMIC_FIELD_COMBINER
    (soatl/field_combiner.h:43-68) with 0, 2 and mixed-type field lists, FieldIdsIncludingSequence (soatl/field_id.h:110-121)"""
    ours, theirs = _probe_on_both_header_trees(tmp_path, "macros_combiners")
    assert ours == ["11 12 13 -1 104 105 | 650 18 0 1", "3 alpha 12", "norm2 norm2 25 1 | qm charge_by_mass 1.5 1 | fortytwo 42 1", "2 4 0 3 1 1"]
    assert theirs is None or theirs == ours


def test_utility_headers_match_the_reference_headers(tmp_path):
    """lambda_is_compatible_with_v / lambda_is_callable_with_args_v and their class forms (include/onika/lambda_tools.h:33-49 in
    the reference; a void-returning call is not "callable with args" there, nor here), IntegralConst arithmetic, comparisons and
    GetAsConstTypeOrValue (integral_constant.h:27-66), FlatTuple get / get_copy / get_nth_const / flat_tuple_subset /
    tuple_size_const_v / make_flat_tuple (flat_tuple.h:27-112), ONIKA_AUTORUN_INIT and the USTAMP / MAKE_UNIQUE_NAME macros
    (cpp_utils.h:23-44)"""
    ours, theirs = _probe_on_both_header_trees(tmp_path, "utility_headers")
    assert ours == ["1 0 1 0 | 1 0 0", "40 1 0 3 3", "4 2 1 2.5 x str | x 1 | 2.5 1", "7 abcd 10203"]
    assert theirs is None or theirs == ours


def test_device_memmove_both_directions_against_std_memmove(tmp_path):
    """onika::cuda::memmove copies forwards (the only direction the reference implements, stl_adaptors.h:46-70 there: it aborts
    when dst > src) and backwards: every (dst offset, src offset, length) of a 96-byte window against std::memmove"""
    src = tmp_path / "mm.cpp"
    src.write_text(r'''
#include <onika/cuda/stl_adaptors.h>
#include <cstdio>
#include <cstring>
int main()
{
  alignas(8) unsigned char a[96], b[96];
  long bad = 0, n = 0;
  for(unsigned d=0; d<40; d++) for(unsigned s=0; s<40; s++) for(unsigned len=0; len<=48; len++)
  {
    for(int i=0;i<96;i++) a[i] = b[i] = (unsigned char)( i * 7 + 3 );
    onika::cuda::memmove( a + d , a + s , len );
    std::memmove( b + d , b + s , len );
    bad += ( std::memcmp( a , b , 96 ) != 0 ); n++;
  }
  std::printf("%ld %ld\n", n, bad);
  return bad != 0;
}
''')
    exe = tmp_path / "mm"
    subprocess.check_call(["g++", "-std=c++20", "-O1", f"-I{INC}", str(src), "-o", str(exe)], env=_env())
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.split() == ["78400", "0"], out.stdout


def test_file_rendezvous_all_gather_between_processes(tmp_path):
    """the handle all-gather MultiGpuExchange falls back to when the application has none (include/onika/parallel/multi_gpu.h:
    file_allgather): 4 processes, 64 bytes each, every process ends with everybody's bytes in rank order; a missing rank times out"""
    from onika_b200 import _build
    lib = _build.build()
    src = tmp_path / "fag.cpp"
    src.write_text(r'''
#include <onika/parallel/multi_gpu.h>
#include <cstdio>
#include <cstdlib>
int main(int argc, char** argv)
{
  const int rank = std::atoi(argv[1]) , world = std::atoi(argv[2]);
  const double timeout = std::atof(argv[4]);
  unsigned char mine[64] , all[64*8];
  for(int i=0;i<64;i++) mine[i] = (unsigned char)( rank * 64 + i );
  const bool ok = onika::parallel::file_allgather( argv[3] , "probe" , rank , world , mine , all , 64 , timeout );
  if( ! ok ) { std::printf("timeout\n"); return 3; }
  for(int r=0;r<world;r++) for(int i=0;i<64;i++) if( all[r*64+i] != (unsigned char)( r * 64 + i ) ) return 4;
  std::printf("ok %d\n", rank);
  return 0;
}
''')
    exe = tmp_path / "fag"
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-fopenmp", f"-I{INC}", str(src), "-o", str(exe), f"-L{os.path.dirname(lib)}", "-lonika_b200",
                           f"-Wl,-rpath,{os.path.dirname(lib)}"], env=_env())
    d1 = tmp_path / "rdv1"
    d1.mkdir()
    procs = [subprocess.Popen([str(exe), str(r), "4", str(d1), "60"], stdout=subprocess.PIPE, text=True) for r in (2, 0, 3, 1)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert sorted(outs) == [f"ok {r}\n" for r in range(4)]
    d2 = tmp_path / "rdv2"
    d2.mkdir()
    lone = subprocess.run([str(exe), "0", "2", str(d2), "0.3"], capture_output=True, text=True, timeout=60)      # rank 1 never comes
    assert lone.returncode == 3 and "timeout" in lone.stdout


def test_msp_yaml_reader_on_the_reference_files(tmp_path):
    """tests/cxx/msp_yaml.h (what `run_operator --app` reads): block maps / sequences, `- key:` items continuing as maps, nested
    flow maps, quoted scalars with escapes, comments, document layering -- on the reference's own data/config/main-config.msp
    and data/exemples/concurrent_block_parallel.msp (read in place; skipped parts when the tree is absent) and on edge cases"""
    here = os.path.dirname(os.path.abspath(__file__))
    src = tmp_path / "yaml_probe.cpp"
    src.write_text(r'''
#include "msp_yaml.h"
int main(int argc, char* argv[])
{
  msp::Node doc; doc.kind = msp::Node::Map;
  for(int i=1;i<argc;i++)
  {
    msp::Node d; std::string err;
    if( ! msp::parse_yaml_file( argv[i] , d , err ) ) { std::fprintf(stderr,"%s\n",err.c_str()); return 3; }
    doc = msp::merge_nodes( doc , d );
  }
  for(const auto& kv : doc.map)
  {
    std::string v = kv.second.str();
    for(auto& c : v) if( c == '\n' ) c = '$';
    std::printf("%s => %s\n", kv.first.c_str(), v.c_str());
  }
  return 0;
}
''')
    exe = tmp_path / "yaml_probe"
    subprocess.check_call(["g++", "-std=c++17", "-Wall", f"-I{os.path.join(here, 'cxx')}", str(src), "-o", str(exe)], env=_env())

    def parse(*paths):
        r = subprocess.run([str(exe), *map(str, paths)], capture_output=True, text=True)
        return r.returncode, dict(ln.split(" => ", 1) for ln in r.stdout.splitlines()), r.stderr

    a = tmp_path / "a.msp"
    a.write_text('top:\n  sub:\n    deep: 1   # comment\n    text: "a # not a comment: still text"\nrun: nop\nlist:\n- x\n- y: { k: [1, 2, {z: 3}] }\n'
                 'sim:\n  - msg: "line\\nbreak"\n  - plain\n  - unit:\n      verbose: true\n      map: { a: b , c: d }\n  - op: { s: 1 }\n    extra: 2\n  -\n    nested: block\n'
                 'multi: { a: 1 ,\n          b: 2 }\n')
    rc, d, err = parse(a)
    assert rc == 0, err
    assert d["top"] == "{sub:{deep:1,text:a # not a comment: still text}}" and d["run"] == "nop"
    assert d["list"] == "[x,{y:{k:[1,2,{z:3}]}}]"
    assert d["sim"] == "[{msg:line$break},plain,{unit:{verbose:true,map:{a:b,c:d}}},{op:{s:1},extra:2},{nested:block}]"
    assert d["multi"] == "{a:1,b:2}"
    b = tmp_path / "b.msp"
    b.write_text("run: run_test\ntop:\n  sub:\n    deep: 2\n  other: 3\nlist: [z]\n")
    rc, d, _ = parse(a, b)                      # layering: maps merge key by key, scalars and sequences are replaced
    assert rc == 0 and d["run"] == "run_test" and d["top"] == "{sub:{deep:2,text:a # not a comment: still text},other:3}" and d["list"] == "[z]"
    for bad, where in (("a:\n    b: 1\n  c: 2\n", ":3:"), ("a: { b: 1\n", ":1:"), ("a:\n  - x\n  y: 1\n", ":3:"), ("a: [1, , 2]\n", ":1:"), ("  a: 1\n", ":1:")):
        f = tmp_path / "bad.msp"
        f.write_text(bad)
        rc, _, err = parse(f)
        assert rc == 3 and where in err, (bad, err)
    ref = "/root/reference/data"
    if os.path.isdir(ref):
        rc, d, err = parse(os.path.join(ref, "config", "main-config.msp"), os.path.join(ref, "exemples", "concurrent_block_parallel.msp"))
        assert rc == 0, err
        assert d["configuration"] == "{logging:{debug:false}}" and d["run"] == "nop" and d["init"] == "nop"
        assert d["run_test"] == "{name:concurrent_block_parallel_test,task_group:true,body:[{concurrent_block_parallel:{value:2.3,rows:4,columns:65536,iter_count:65536}}]}"
        sim = d["simulation"]
        assert sim.startswith("[{message:") and sim.endswith(",mpi_comm_world,init_cuda,global,{unit_system:{verbose:true,unit_system:{length:meter,mass:kilogram,time:second,"
                                                                "charge:coulomb,temperature:kelvin,amount:mol,luminosity:candela,angle:radian,energy:joule}}},init,run,finalize_cuda]")
