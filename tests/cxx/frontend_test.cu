// tests/cxx/frontend_test.cu -- scenarios written against the onika API exactly as reference plugins / tests use it
// (onikaParallelTutorial plugins, tests/parallel_queue_lane.cu, tests/benchmark.cpp, tests/soatlserializetest.cpp,
// tests/soatupletest.cpp), compiled against THIS repo's include/onika front end and libonika_b200.so.
// usage: frontend_test <scenario> <output file> [args...] ; raw little-endian doubles (or bytes) are written to the
// output file and compared with the oracle by tests/test_gpu_cxx_frontend.py.
#include <onika/parallel/block_parallel_for.h>
#include <onika/parallel/parallel_for.h>
#include <onika/parallel/memset.h>
#include <onika/parallel/reduction.h>
#include <onika/parallel/parallel_execution_context_allocator.h>
#include <onika/parallel/parallel_execution_queue.h>
#include <onika/parallel/parallel_execution_operators.h>
#include <onika/extras/array2d.h>
#include <onika/extras/block_parallel_value_add_functor.h>
#include <onika/cuda/cuda_reduction.h>
#include <onika/soatl/field_arrays.h>
#include <onika/soatl/pretty_print.h>
#include <onika/soatl/field_id_tuple_utils.h>
#include <onika/soatl/compute.h>
#include <onika/soatl/copy.h>
#include <onika/memory/allocator.h>
#include <onika/cuda/memcpy.h>
#include <onika/cuda/input_array_span.h>
#include <onika/cuda/ro_shallow_copy.h>
#include <onika/cuda/cuda_math.h>
#include "declare_fields.h"

#include <cfloat>
#include <chrono>
#include <sstream>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

using namespace onika;
using namespace onika::parallel;
using onika::memory::CudaMMVector;

// same hash as tests/_inputs.py:lcg_uniform
static double lcg(uint64_t i, uint64_t seed, double lo, double hi)
{
  uint64_t h = ( i * 2654435761ull + seed * 40503ull + 12345ull ) & 0xFFFFFFFFull;
  h = ( ( h ^ ( h >> 15 ) ) * 2246822519ull ) & 0xFFFFFFFFull;
  h = ( h ^ ( h >> 13 ) ) & 0xFFFFFFFFull;
  return lo + ( hi - lo ) * ( double(h) / 4294967296.0 );
}
template<class V> static void fill(V& v, uint64_t seed, double lo, double hi) { for(size_t i=0;i<v.size();i++) v[i] = lcg(i,seed,lo,hi); }

struct Out
{
  FILE* f;
  explicit Out(const char* path) : f( std::fopen(path,"wb") ) { if(!f) { std::perror(path); std::exit(2); } }
  ~Out() { std::fclose(f); }
  void doubles(const double* p, size_t n) { std::fwrite( p , sizeof(double) , n , f ); }
  void bytes(const void* p, size_t n) { std::fwrite( p , 1 , n , f ); }
  void u64(uint64_t v) { std::fwrite( &v , 8 , 1 , f ); }
};

static ParallelExecutionContextAllocator& peca() { static auto* a = new ParallelExecutionContextAllocator(); return *a; }
static ParallelExecutionContext* parallel_execution_context(const char* tag, const char* sub = "")
{
  return peca().create( tag , sub , & ParallelExecutionQueue::default_queue() , onika::cuda::CudaContext::default_cuda_ctx() );
}

// ---------------------------------------------------------------- user functors (what a plugin author writes)
struct AxpyFunctor
{
  double * __restrict__ y; const double * __restrict__ x; double a;
  ONIKA_HOST_DEVICE_FUNC inline void operator () ( size_t i ) const { y[i] = y[i] + a * x[i]; }
};
namespace onika { namespace parallel { template<> struct ParallelForFunctorTraits<AxpyFunctor> { static inline constexpr bool CudaCompatible = true; }; } }

struct Grid3DFunctor   // plugins/onikaParallelTutorial/parallel_for_3d_benchmark.cu: Grid3DBenchmarkFunctor
{
  double * const __restrict__ m_array = nullptr;
  const long m_size = 0;
  ONIKA_HOST_DEVICE_FUNC inline void operator () ( onikaInt3_t c ) const
  {
    const unsigned long N = m_size, i = c.x, j = c.y, k = c.z;
    ONIKA_CU_ATOMIC_ADD( m_array[ (k*N+j)*N+i ] , 1.0 );
  }
};
// block-level variant: the whole block is called once per cell, one thread counts the visit (on the host a block is one
// thread, so both targets count 1 per cell; the unguarded tutorial functor counts blockDim per cell on any GPU)
struct Grid3DBlockFunctor
{
  double * const __restrict__ m_array = nullptr;
  const long m_size = 0;
  ONIKA_HOST_DEVICE_FUNC inline void operator () ( onikaInt3_t c ) const
  {
    const unsigned long N = m_size, i = c.x, j = c.y, k = c.z;
    if( ONIKA_CU_THREAD_IDX == 0 ) ONIKA_CU_ATOMIC_ADD( m_array[ (k*N+j)*N+i ] , 1.0 );
  }
};
namespace onika { namespace parallel {
  template<> struct BlockParallelForFunctorTraits<Grid3DBlockFunctor> { static inline constexpr bool CudaCompatible = true; };
  template<> struct BlockParallelForFunctorTraits<Grid3DFunctor> { static inline constexpr bool CudaCompatible = true; };
  template<> struct ParallelForFunctorTraits<Grid3DFunctor> { static inline constexpr bool CudaCompatible = true; };
} }

// reduction through the return-data channel: block_reduce_* then one atomic per block
struct ReduceResult { double vmin; double vmax; double vsum; unsigned long long count; };
struct RowReduceFunctor
{
  const double * __restrict__ data; size_t cols; ReduceResult* result;   // result: device return-data pointer
  ONIKA_HOST_DEVICE_FUNC inline void operator () ( size_t row ) const
  {
    double mn = DBL_MAX, mx = -DBL_MAX, sm = 0.0; unsigned long long cnt = 0;
    ONIKA_CU_BLOCK_SIMD_FOR(size_t, j, 0, cols)
    {
      const double v = data[ row*cols + j ];
      if( v < mn ) mn = v;
      if( v > mx ) mx = v;
      sm += v; ++cnt;
    }
    mn = onika::cuda::block_reduce_min( mn , onika::IntConst<ONIKA_CU_MAX_THREADS_PER_BLOCK>{} );
    mx = onika::cuda::block_reduce_max( mx , onika::IntConst<ONIKA_CU_MAX_THREADS_PER_BLOCK>{} );
    sm = onika::cuda::block_reduce_add( sm , onika::IntConst<ONIKA_CU_MAX_THREADS_PER_BLOCK>{} );
    cnt = onika::cuda::block_reduce_add( cnt , onika::IntConst<ONIKA_CU_MAX_THREADS_PER_BLOCK>{} );
    if( ONIKA_CU_THREAD_IDX == 0 )
    {
      ONIKA_CU_ATOMIC_MIN( result->vmin , mn );
      ONIKA_CU_ATOMIC_MAX( result->vmax , mx );
      ONIKA_CU_ATOMIC_ADD( result->vsum , sm );
      ONIKA_CU_ATOMIC_ADD( result->count , cnt );
    }
  }
};
namespace onika { namespace parallel { template<> struct BlockParallelForFunctorTraits<RowReduceFunctor> { static inline constexpr bool CudaCompatible = true; }; } }

// cell grid the way an exaNBody operator sees it: an array of per-cell FieldArrays
using CellArrays = soatl::FieldArrays<64,16,_tag_particle_rx,_tag_particle_ry,_tag_particle_rz,_tag_particle_e>;
struct CellSumFunctor
{
  const CellArrays * cells; double * cell_sum; double * total;
  ONIKA_HOST_DEVICE_FUNC inline void operator () ( size_t c ) const
  {
    const double * __restrict__ e = cells[c][particle_e];
    const size_t n = cells[c].size();
    double s = 0.0;
    if( ONIKA_CU_THREAD_IDX == 0 ) { for(size_t p=0;p<n;p++) s += e[p]; cell_sum[c] = s; ONIKA_CU_ATOMIC_ADD( *total , s ); }   // particle order
  }
};
namespace onika { namespace parallel { template<> struct BlockParallelForFunctorTraits<CellSumFunctor> { static inline constexpr bool CudaCompatible = true; }; } }

// prolog / epilog protocol + element list
struct ScaleListedFunctor
{
  double * __restrict__ data; double s; int * prolog_calls; int * epilog_calls;
  ONIKA_HOST_DEVICE_FUNC inline void operator () ( ssize_t i ) const { if( ONIKA_CU_THREAD_IDX == 0 ) data[i] *= s; }
  inline void operator () ( block_parallel_for_gpu_prolog_t ) const { ++ *prolog_calls; }
  inline void operator () ( block_parallel_for_gpu_epilog_t ) const { ++ *epilog_calls; }
};
namespace onika { namespace parallel { template<> struct BlockParallelForFunctorTraits<ScaleListedFunctor> { static inline constexpr bool CudaCompatible = true; }; } }

// helpers of include/onika/cuda: cu_block_memcpy / cu_block_memmove, InputArraySpan (small array inside the functor),
// VectorShallowCopy, cuda_math
struct HelpersFunctor
{
  const uint8_t* src; uint8_t* dst; uint8_t* mv; size_t row_bytes;
  onika::cuda::InputArraySpan<double,8> coeffs;            // <= 8 elements: travels inside the kernel parameters
  onika::cuda::VectorShallowCopy<double> table;
  double* out;
  ONIKA_HOST_DEVICE_FUNC inline void operator () ( size_t row ) const
  {
    // misaligned byte ranges on purpose: row r starts at r*row_bytes + (r % 7)
    onika::cuda::cu_block_memcpy( dst + row*row_bytes + (row%7) , src + row*row_bytes + (row%5) , row_bytes - 8 );
    // overlapping move inside mv: shift the row's payload by 3 bytes to the right
    onika::cuda::cu_block_memmove( mv + row*row_bytes + 3 , mv + row*row_bytes , row_bytes - 8 );
    if( ONIKA_CU_THREAD_IDX == 0 )
    {
      double acc = 0.0;
      for(size_t k=0;k<coeffs.size();k++) acc += coeffs[k] * table[ ( row + k ) % table.size() ];
      out[row] = onika::cuda::clamp( acc , -1.0 , 1.0 ) + onika::cuda::pow_ce_dint( 2.0 , double(row % 5) );
    }
  }
};
namespace onika { namespace parallel { template<> struct BlockParallelForFunctorTraits<HelpersFunctor> { static inline constexpr bool CudaCompatible = true; }; } }

// a kernel of the user's own, launched with ONIKA_CU_LAUNCH_KERNEL on a lane's stream (as tests/gpu_uvm_benchmark.cu does)
ONIKA_DEVICE_KERNEL_FUNC void user_double_kernel( double * data , size_t n )
{
  const size_t i = size_t(ONIKA_CU_BLOCK_IDX) * ONIKA_CU_BLOCK_SIZE + ONIKA_CU_THREAD_IDX;
  if( i < n ) data[i] = data[i] * 2.0;
}

static int scenario_helpers(Out& out, size_t rows)
{
  const size_t row_bytes = 1000;
  CudaMMVector<uint8_t> src( rows*row_bytes ), dst( rows*row_bytes , uint8_t(0xEE) ), mv( rows*row_bytes );
  for(size_t i=0;i<src.size();i++) { src[i] = uint8_t( ( i * 131u + 7u ) >> 3 ); mv[i] = uint8_t( ( i * 29u + 1u ) >> 2 ); }
  std::vector<double> c = { 0.5 , -0.25 , 0.125 , 1.5 , -2.0 };
  CudaMMVector<double> table( 64 ); fill( table , 4 , -1.0 , 1.0 );
  CudaMMVector<double> res( rows , 0.0 );
  HelpersFunctor f = { src.data() , dst.data() , mv.data() , row_bytes , onika::cuda::InputArraySpan<double,8>( c ) , onika::cuda::VectorShallowCopy<double>( table ) , res.data() };
  block_parallel_for( rows , f , parallel_execution_context("helpers") );
  {
    auto stream = onika::cuda::CudaContext::default_cuda_ctx()->getThreadStream( 2 );
    ONIKA_CU_LAUNCH_KERNEL( (rows+127)/128 , 128 , 0 , stream , user_double_kernel , res.data() , rows );
    ONIKA_CU_CHECK_ERRORS( ONIKA_CU_STREAM_SYNCHRONIZE( stream ) );
  }
  out.bytes( dst.data() , dst.size() );
  out.bytes( mv.data() , mv.size() );
  out.doubles( res.data() , rows );
  return 0;
}

// the block vocabulary inside a CTA of sub-blocks (and, with the same source, inside ordinary CTAs): every item records what
// its "block" looks like through the ONIKA_CU_* macros and the block_reduce_* helpers
template<int BS>
struct BlockVocabularyFunctor
{
  const double* values; double* out;      // out: 8 doubles per item
  ONIKA_HOST_DEVICE_FUNC inline void operator () ( size_t i ) const
  {
    const unsigned int tid = ONIKA_CU_THREAD_IDX , bs = ONIKA_CU_BLOCK_SIZE;
    const double v = values[ i * BS + ( tid % BS ) ];
    const double sum_ids = onika::cuda::block_reduce_add( double(tid+1) , onika::IntConst<BS>{} );       // bs*(bs+1)/2
    const double vmin = onika::cuda::block_reduce_min( v , onika::IntConst<BS>{} );
    const double vmax = onika::cuda::block_reduce_max( v , onika::IntConst<BS>{} );
    const int any = ONIKA_CU_BLOCK_SYNC_OR( tid == ( i % BS ) );               // exactly one thread of the block says yes
    const int all = ONIKA_CU_BLOCK_SYNC_AND( tid < bs );
    const int none = ONIKA_CU_BLOCK_SYNC_OR( tid >= bs );
    double cnt = 0.0;
    ONIKA_CU_BLOCK_SIMD_FOR(unsigned int, k, 0, 3u*BS+1u) { cnt += 1.0; }       // iterations shared by the block's threads
    cnt = onika::cuda::block_reduce_add( cnt , onika::IntConst<BS>{} );         // 3*BS+1 in total
    ONIKA_CU_BLOCK_SYNC();
    if( tid == 0 )
    {
      double* o = out + i * 8;
      o[0] = double(bs); o[1] = sum_ids; o[2] = vmin; o[3] = vmax; o[4] = double( ( any != 0 ) + 2 * ( all != 0 ) + 4 * ( none != 0 ) );
      o[5] = cnt; o[6] = double( ONIKA_CU_BLOCK_IDX < ONIKA_CU_GRID_SIZE ); o[7] = double( ONIKA_CU_BLOCK_DIMS.x * ONIKA_CU_BLOCK_DIMS.y * ONIKA_CU_BLOCK_DIMS.z );
    }
  }
};
template<int BS> struct PackedBlockVocabularyFunctor : public BlockVocabularyFunctor<BS> {};
namespace onika { namespace parallel {
  template<int BS> struct BlockParallelForFunctorTraits< BlockVocabularyFunctor<BS> > { static inline constexpr bool CudaCompatible = true; };
  template<int BS> struct BlockParallelForFunctorTraits< PackedBlockVocabularyFunctor<BS> > { static inline constexpr bool CudaCompatible = true; static inline constexpr bool SubBlockItems = true; };
} }

static int scenario_block_vocabulary(Out& out, size_t nitems)
{
  CudaMMVector<double> values( nitems * 128 ); fill( values , 21 , -1.0 , 1.0 );
  auto run = [&](auto functor_of, unsigned int bs)
  {
    CudaMMVector<double> res( nitems * 8 , -1.0 );
    BlockParallelForOptions o; o.max_block_size = bs;
    block_parallel_for( nitems , functor_of( res.data() ) , parallel_execution_context("vocabulary") , o );
    out.doubles( res.data() , res.size() );
  };
  // ordinary CTAs (one block per item at a time) and CTAs of sub-blocks, block sizes 8 / 16 / 32; then a 128-thread block
  run( [&](double* r) { return BlockVocabularyFunctor<8>{ values.data() , r }; } , 8 );
  run( [&](double* r) { return PackedBlockVocabularyFunctor<8>{ { values.data() , r } }; } , 8 );
  run( [&](double* r) { return BlockVocabularyFunctor<16>{ values.data() , r }; } , 16 );
  run( [&](double* r) { return PackedBlockVocabularyFunctor<16>{ { values.data() , r } }; } , 16 );
  run( [&](double* r) { return BlockVocabularyFunctor<32>{ values.data() , r }; } , 32 );
  run( [&](double* r) { return PackedBlockVocabularyFunctor<32>{ { values.data() , r } }; } , 32 );
  run( [&](double* r) { return PackedBlockVocabularyFunctor<128>{ { values.data() , r } }; } , 128 );   // too large to pack: ordinary CTAs
  return 0;
}

// ---------------------------------------------------------------- scenarios
static int scenario_axpy(Out& out, size_t n)
{
  CudaMMVector<double> x(n), y(n);
  fill(x,1,-1,1); fill(y,2,-1,1);
  AxpyFunctor f{ y.data(), x.data(), 0.5 };
  parallel_for( n , f , parallel_execution_context("axpy") );     // synchronous: the temporary wrapper dies here
  out.doubles( y.data() , n );
  // parallel_memset on a second buffer
  CudaMMVector<double> z(n, -1.0);
  parallel_memset( z.data() , n , 3.25 , parallel_execution_context("memset") );
  out.doubles( z.data() , n );
  return 0;
}

static int scenario_value_add(Out& out, size_t rows, size_t cols)
{
  extras::Array2D a; a.resize(rows,cols);
  fill(a.m_data,7,1e-3,1.0);
  extras::BlockParallelValueAddFunctor<true> f = { a , 1.1 };
  block_parallel_for( rows , f , parallel_execution_context("vadd","1d") );
  out.doubles( a.data() , rows*cols );
  const ParallelExecutionSpace<2> r2 = { {0,0} , { ssize_t(rows) , ssize_t(cols) } };
  block_parallel_for( r2 , f , parallel_execution_context("vadd","2d") );
  out.doubles( a.data() , rows*cols );
  // work-stealing grid (fixed_gpu_grid_size), 1D
  BlockParallelForOptions o; o.fixed_gpu_grid_size = true;
  block_parallel_for( rows , f , parallel_execution_context("vadd","ws") , o );
  out.doubles( a.data() , rows*cols );
  return 0;
}

// tests/parallel_queue_lane.cu run_test(): kernel1 on the GPU, kernel2 declared host-only (hybrid), 4 queue modes
static int scenario_lanes(Out& out, const std::string& mode, size_t rows, size_t cols, bool kernel_1d)
{
  extras::Array2D array1, array2;
  array1.resize(rows,cols); array2.resize(rows,cols);
  fill(array1.m_data,8,1e-3,1.0); fill(array2.m_data,9,1e-3,1.0);
  const ParallelExecutionSpace<2> range2d = { {0,0} , { ssize_t(rows) , ssize_t(cols) } };
  extras::BlockParallelValueAddFunctor<true>  array1_kernel1 = { array1 , 1.1 };
  extras::BlockParallelValueAddFunctor<false> array1_kernel2 = { array1 , 1.2 };
  extras::BlockParallelValueAddFunctor<true>  array2_kernel1 = { array2 , 1.3 };
  extras::BlockParallelValueAddFunctor<false> array2_kernel2 = { array2 , 1.4 };
  auto & pq = ParallelExecutionQueue::default_queue();
  auto op11 = kernel_1d ? block_parallel_for( rows , array1_kernel1 , parallel_execution_context("Array1","Kernel1") )
                        : block_parallel_for( range2d , array1_kernel1 , parallel_execution_context("Array1","Kernel1") );
  auto op12 = kernel_1d ? block_parallel_for( rows , array1_kernel2 , parallel_execution_context("Array1","Kernel2") )
                        : block_parallel_for( range2d , array1_kernel2 , parallel_execution_context("Array1","Kernel2") );
  auto op21 = kernel_1d ? block_parallel_for( rows , array2_kernel1 , parallel_execution_context("Array2","Kernel1") )
                        : block_parallel_for( range2d , array2_kernel1 , parallel_execution_context("Array2","Kernel1") );
  auto op22 = kernel_1d ? block_parallel_for( rows , array2_kernel2 , parallel_execution_context("Array2","Kernel2") )
                        : block_parallel_for( range2d , array2_kernel2 , parallel_execution_context("Array2","Kernel2") );
  std::shared_ptr<std::mutex> start_sync;
  if( mode == "hybrid-delayed-lane" )
  {
    ParallelExecutionQueueBase qa, qb;
    qa << set_lane(0) << std::move(op11) << set_lane(1) << std::move(op21);
    qb << set_lane(0) << std::move(op12) << set_lane(1) << std::move(op22);
    pq << std::move(qa) << std::move(qb);
  }
  else if( mode == "hybrid-lane" )
  {
    pq << set_lane(0) << std::move(op11) << set_lane(1) << std::move(op21)
       << set_lane(0) << std::move(op12) << set_lane(1) << std::move(op22);
  }
  else if( mode == "hybrid-sequence" )
  {
    pq << std::move(op11) << std::move(op12) << std::move(op21) << std::move(op22);
  }
  else if( mode == "singletask-dependency" )
  {
    const auto a1_loc = local_access( array1.m_data.data() , 2 , AccessStencilElement::RW , "a1_loc" );
    const auto a1_nbh = access_cross_2d( array1.m_data.data() , AccessStencilElement::RW , AccessStencilElement::RO , "a1_nbh" );
    const ParallelExecutionSpace<2> single_task_space = { {64,64} , {980,980} };
    start_sync = std::make_shared<std::mutex>();
    start_sync->lock();
    auto single_task = make_single_task_block_parallel_functor( [start_sync]() { start_sync->lock(); start_sync->unlock(); } );
    pq << any_lane() << a1_loc << block_parallel_for( single_task_space , single_task , parallel_execution_context("Array1","UnlockTask1") )
       << any_lane() << a1_nbh << std::move(op11)
       << any_lane() << a1_nbh << std::move(op12);
    // array2 operations run synchronously afterwards (their wrappers die at the end of this function otherwise)
  }
  else { std::fprintf(stderr,"unknown test mode %s\n",mode.c_str()); return 2; }
  if( start_sync != nullptr ) start_sync->unlock();
  pq << flush;
  const bool idle_probe = pq.query_status();   // may be true or false; must not deadlock
  (void) idle_probe;
  pq << synchronize;
  if( mode == "singletask-dependency" ) { { auto w = std::move(op21); } { auto w = std::move(op22); } }
  out.doubles( array1.data() , rows*cols );
  out.doubles( array2.data() , rows*cols );
  return pq.empty() ? 0 : 3;
}

// automatic lanes from declared data accesses (plugins/onikaParallelTutorial/automatic_concurrent_parallel.cu pattern,
// extended with an operation that depends on two lanes): results must equal serial FIFO execution
struct RepeatedAddFunctor   // a[i][j] += v, `iters` times, one rounding per add: slow on purpose, exact by construction
{
  extras::Array2DReference a; double v; int iters;
  ONIKA_HOST_DEVICE_FUNC inline void operator () ( size_t i ) const
  {
    ONIKA_CU_BLOCK_SIMD_FOR(size_t, j, 0, a.columns()) { double x = a[i][j]; for(int k=0;k<iters;k++) x = x + v; a[i][j] = x; }
  }
};
struct AddArrayFunctor      // b[i][j] += a[i][j]
{
  extras::Array2DReference a; extras::Array2DReference b;
  ONIKA_HOST_DEVICE_FUNC inline void operator () ( size_t i ) const
  {
    ONIKA_CU_BLOCK_SIMD_FOR(size_t, j, 0, a.columns()) b[i][j] += a[i][j];
  }
};
namespace onika { namespace parallel {
  template<> struct BlockParallelForFunctorTraits<RepeatedAddFunctor> { static inline constexpr bool CudaCompatible = true; };
  template<> struct BlockParallelForFunctorTraits<AddArrayFunctor> { static inline constexpr bool CudaCompatible = true; };
} }

static int scenario_autolanes(Out& out, size_t rows, size_t cols, int iters)
{
  extras::Array2D A, B, C;
  A.resize(rows,cols); B.resize(rows,cols); C.resize(rows,cols);
  fill(A.m_data,8,1e-3,1.0); fill(B.m_data,9,1e-3,1.0); fill(C.m_data,10,1e-3,1.0);
  const auto a_rw = local_access( A.m_data.data() , 2 , AccessStencilElement::RW , "a_rw" );
  const auto a_ro = local_access( A.m_data.data() , 2 , AccessStencilElement::RO , "a_ro" );
  const auto b_rw = local_access( B.m_data.data() , 2 , AccessStencilElement::RW , "b_rw" );
  const auto c_rw = local_access( C.m_data.data() , 2 , AccessStencilElement::RW , "c_rw" );
  ParallelExecutionContext* pec[6] = { parallel_execution_context("A","add1") , parallel_execution_context("B","slow") , parallel_execution_context("A","add2")
                                     , parallel_execution_context("B","+=A") , parallel_execution_context("A","add3") , parallel_execution_context("C","add") };
  auto & pq = ParallelExecutionQueue::default_queue();
  pq << any_lane(4)
     << a_rw << block_parallel_for( rows , extras::BlockParallelValueAddFunctor<true>{ A , 1.1 } , pec[0] )
     << b_rw << block_parallel_for( rows , RepeatedAddFunctor{ B , 1.3 , iters } , pec[1] )              // long: op 3 must wait for it across lanes
     << a_rw << block_parallel_for( rows , extras::BlockParallelValueAddFunctor<true>{ A , 1.2 } , pec[2] )
     << a_ro << b_rw << block_parallel_for( rows , AddArrayFunctor{ A , B } , pec[3] )                   // reads A (lane of 0,2), writes B (lane of 1)
     << a_rw << block_parallel_for( rows , extras::BlockParallelValueAddFunctor<true>{ A , 1.4 } , pec[4] ) // write-after-read on A
     << c_rw << block_parallel_for( rows , RepeatedAddFunctor{ C , 0.7 , 3 } , pec[5] );                  // independent of everything
  pq << flush;
  double lanes[6], waits[6];
  for(int k=0;k<6;k++)
  {
    lanes[k] = pec[k]->m_lane;
    int nw = 0; for(int l=0;l<256;l++) if( pec[k]->depends_on_lane(l) ) ++nw;
    waits[k] = nw;
  }
  pq << synchronize;
  out.doubles( A.data() , rows*cols ); out.doubles( B.data() , rows*cols ); out.doubles( C.data() , rows*cols );
  out.doubles( lanes , 6 ); out.doubles( waits , 6 );
  return pq.empty() ? 0 : 3;
}

// ordering rule of the lane placement for one-sided declarations (any_lane + accesses): a WO producer followed by an RO
// consumer (read-after-write) and an RO consumer followed by a WO producer (write-after-read) must both order, two readers
// need not.  The slow kernels make a missing edge visible in the results.
static int scenario_autolanes_raw_war(Out& out, size_t rows, size_t cols, int iters)
{
  extras::Array2D A, B, C, D;
  A.resize(rows,cols); B.resize(rows,cols); C.resize(rows,cols); D.resize(rows,cols);
  fill(A.m_data,8,1e-3,1.0); fill(B.m_data,9,1e-3,1.0); fill(C.m_data,10,1e-3,1.0); fill(D.m_data,11,1e-3,1.0);
  const auto a_wo = local_access( A.m_data.data() , 2 , AccessStencilElement::WO , "a_wo" );
  const auto a_ro = local_access( A.m_data.data() , 2 , AccessStencilElement::RO , "a_ro" );
  const auto b_rw = local_access( B.m_data.data() , 2 , AccessStencilElement::RW , "b_rw" );
  const auto c_rw = local_access( C.m_data.data() , 2 , AccessStencilElement::RW , "c_rw" );
  const auto d_rw = local_access( D.m_data.data() , 2 , AccessStencilElement::RW , "d_rw" );
  ParallelExecutionContext* pec[6] = { parallel_execution_context("A","produce") , parallel_execution_context("B","+=A") , parallel_execution_context("C","slow")
                                     , parallel_execution_context("C","+=A") , parallel_execution_context("A","overwrite") , parallel_execution_context("D","+=A") };
  auto & pq = ParallelExecutionQueue::default_queue();
  pq << any_lane(4)
     << a_wo << block_parallel_for( rows , RepeatedAddFunctor{ A , 1.1 , iters } , pec[0] )      // slow producer of A, declared write-only
     << a_ro << b_rw << block_parallel_for( rows , AddArrayFunctor{ A , B } , pec[1] )             // RAW: must see the produced A
     << c_rw << block_parallel_for( rows , RepeatedAddFunctor{ C , 1.3 , iters } , pec[2] )      // slow, independent of A
     << a_ro << c_rw << block_parallel_for( rows , AddArrayFunctor{ A , C } , pec[3] )             // reads A behind the slow kernel on C
     << a_wo << block_parallel_for( rows , extras::BlockParallelValueAddFunctor<true>{ A , 1.4 } , pec[4] )   // WAR: must wait for BOTH readers of A
     << a_ro << d_rw << block_parallel_for( rows , AddArrayFunctor{ A , D } , pec[5] );            // RAW again, after the overwrite
  pq << flush;
  double lanes[6], waits[6];
  for(int k=0;k<6;k++)
  {
    lanes[k] = pec[k]->m_lane;
    int nw = 0; for(int l=0;l<256;l++) if( pec[k]->depends_on_lane(l) ) ++nw;
    waits[k] = nw;
  }
  pq << synchronize;
  out.doubles( A.data() , rows*cols ); out.doubles( B.data() , rows*cols ); out.doubles( C.data() , rows*cols ); out.doubles( D.data() , rows*cols );
  out.doubles( lanes , 6 ); out.doubles( waits , 6 );
  // the printed test of the reference keeps its (different) meaning: writer vs reader -> no "conflict", reader vs reader -> "conflict"
  ParallelDataAccessVector vw, vr; vw.push_back( a_wo ); vr.push_back( a_ro );
  const double ref_rule[4] = { double( concurrent_data_access( vw , vr ) ) , double( concurrent_data_access( vr , vr ) ) ,
                               double( data_access_dependency( vw , vr ) ) , double( data_access_dependency( vr , vr ) ) };
  out.doubles( ref_rule , 4 );
  return pq.empty() ? 0 : 3;
}

// what the lane DAG buys: the tutorial pattern (2 arrays x 2 kernels, few blocks each) serialised on one lane, on two
// explicit lanes, and under any_lane() with declared accesses.  Prints wall milliseconds of flush..synchronize.
static int scenario_autolanes_perf(Out& out, size_t rows, size_t cols, int iters)
{
  extras::Array2D A, B;
  A.resize(rows,cols); B.resize(rows,cols);
  fill(A.m_data,8,1e-3,1.0); fill(B.m_data,9,1e-3,1.0);
  const auto a_rw = local_access( A.m_data.data() , 2 , AccessStencilElement::RW , "a_rw" );
  const auto b_rw = local_access( B.m_data.data() , 2 , AccessStencilElement::RW , "b_rw" );
  auto & pq = ParallelExecutionQueue::default_queue();
  double ms[3] = { 0 , 0 , 0 };
  for(int rep=0; rep<3; rep++) for(int mode=0; mode<3; mode++)
  {
    auto a1 = block_parallel_for( rows , RepeatedAddFunctor{ A , 1.1 , iters } , parallel_execution_context("A","k1") );
    auto a2 = block_parallel_for( rows , RepeatedAddFunctor{ A , 1.2 , iters } , parallel_execution_context("A","k2") );
    auto b1 = block_parallel_for( rows , RepeatedAddFunctor{ B , 1.3 , iters } , parallel_execution_context("B","k1") );
    auto b2 = block_parallel_for( rows , RepeatedAddFunctor{ B , 1.4 , iters } , parallel_execution_context("B","k2") );
    if( mode == 0 )      pq << set_lane(0) << std::move(a1) << std::move(a2) << std::move(b1) << std::move(b2);
    else if( mode == 1 ) pq << set_lane(0) << std::move(a1) << set_lane(1) << std::move(b1) << set_lane(0) << std::move(a2) << set_lane(1) << std::move(b2);
    else                 pq << any_lane() << a_rw << std::move(a1) << a_rw << std::move(a2) << b_rw << std::move(b1) << b_rw << std::move(b2);
    const auto t0 = std::chrono::steady_clock::now();
    pq << flush << synchronize;
    const double t = std::chrono::duration<double,std::milli>( std::chrono::steady_clock::now() - t0 ).count();
    if( rep == 0 || t < ms[mode] ) ms[mode] = t;      // best of 3 (first round also warms up the managed arrays)
  }
  std::printf("autolanes_perf rows=%zu cols=%zu iters=%d : one_lane %.3f ms , explicit_lanes %.3f ms , any_lane+accesses %.3f ms\n", rows, cols, iters, ms[0], ms[1], ms[2]);
  out.doubles( ms , 3 );
  out.doubles( A.data() , rows*cols ); out.doubles( B.data() , rows*cols );
  return pq.empty() ? 0 : 3;
}

// capture(g) : the tutorial lane pattern (GPU kernel then host-only kernel per array, any_lane + accesses, plus a
// return-data reduction) recorded once and replayed `replays` times; also times N small operations direct vs replayed
static int scenario_graph(Out& out, size_t rows, size_t cols, int replays, int nsmall)
{
  extras::Array2D A, B;
  A.resize(rows,cols); B.resize(rows,cols);
  fill(A.m_data,8,1e-3,1.0); fill(B.m_data,9,1e-3,1.0);
  const auto a_rw = local_access( A.m_data.data() , 2 , AccessStencilElement::RW , "a_rw" );
  const auto b_rw = local_access( B.m_data.data() , 2 , AccessStencilElement::RW , "b_rw" );
  auto & pq = ParallelExecutionQueue::default_queue();
  ReduceResult* red = nullptr;                                  // pinned: the D2H copy of every replay lands here
  ONIKA_B200_CHECK( onk_alloc( sizeof(ReduceResult) , 64 , ONK_MEM_PINNED , (void**) &red ) );
  *red = ReduceResult{ DBL_MAX , -DBL_MAX , 0.0 , 0 };
  double n_ops = 0, n_lanes = 0;
  {
    ParallelExecutionGraph g;
    auto* red_pec = parallel_execution_context("A","reduce");
    red_pec->init_device_scratch();
    BlockParallelForOptions ropts; ropts.return_data = red; ropts.return_data_size = sizeof(ReduceResult);
    pq << capture(g) << any_lane(4)
       << a_rw << block_parallel_for( rows , extras::BlockParallelValueAddFunctor<true>{ A , 1.1 } , parallel_execution_context("A","k1") )
       << b_rw << block_parallel_for( rows , extras::BlockParallelValueAddFunctor<true>{ B , 1.3 } , parallel_execution_context("B","k1") )
       << a_rw << block_parallel_for( rows , extras::BlockParallelValueAddFunctor<false>{ A , 1.2 } , parallel_execution_context("A","k2host") )
       << b_rw << block_parallel_for( rows , extras::BlockParallelValueAddFunctor<true>{ B , 1.4 } , parallel_execution_context("B","k2") )
       << a_rw << block_parallel_for( rows , RowReduceFunctor{ A.data() , cols , (ReduceResult*) red_pec->get_device_return_data_ptr() } , red_pec , ropts )
       << flush;
    if( ! pq.empty() ) return 4;                              // recorded operations belong to the graph, not to the queue
    n_ops = g.number_of_operations(); n_lanes = g.number_of_lanes();
    for(int r=0;r<replays;r++) g.launch();
    g.wait();
  }
  out.doubles( A.data() , rows*cols ); out.doubles( B.data() , rows*cols );
  const double redv[4] = { red->vmin , red->vmax , red->vsum , double(red->count) };
  out.doubles( redv , 4 ); out.doubles( &n_ops , 1 ); out.doubles( &n_lanes , 1 );
  onk_free( red , sizeof(ReduceResult) );

  // launch overhead: nsmall dependent small operations, scheduled one by one vs replayed from a graph
  extras::Array2D S; S.resize(8,256);
  for(size_t k=0;k<8*256;k++) S.m_data[k] = 0.0;
  double ms[2] = { 0 , 0 }, submit_ms = 0;
  for(int rep=0;rep<3;rep++)
  {
    for(int k=0;k<nsmall;k++) pq << set_lane(0) << block_parallel_for( 8 , RepeatedAddFunctor{ S , 1.0 , 1 } , parallel_execution_context("S","direct") );
    auto t0 = std::chrono::steady_clock::now();
    pq << flush;
    const double tf = std::chrono::duration<double,std::milli>( std::chrono::steady_clock::now() - t0 ).count();
    pq << synchronize;
    const double t = std::chrono::duration<double,std::milli>( std::chrono::steady_clock::now() - t0 ).count();
    if( rep == 0 || t < ms[0] ) { ms[0] = t; submit_ms = tf; }
  }
  double untimed_ms = 0;
  ParallelExecutionContext::s_gpu_timing = false;              // same operations without the per-operation timing events
  for(int rep=0;rep<3;rep++)
  {
    for(int k=0;k<nsmall;k++) pq << set_lane(0) << block_parallel_for( 8 , RepeatedAddFunctor{ S , 1.0 , 1 } , parallel_execution_context("S","untimed") );
    auto t0 = std::chrono::steady_clock::now();
    pq << flush << synchronize;
    const double t = std::chrono::duration<double,std::milli>( std::chrono::steady_clock::now() - t0 ).count();
    if( rep == 0 || t < untimed_ms ) untimed_ms = t;
  }
  ParallelExecutionContext::s_gpu_timing = true;
  {
    ParallelExecutionGraph g;
    pq << capture(g);
    for(int k=0;k<nsmall;k++) pq << set_lane(0) << block_parallel_for( 8 , RepeatedAddFunctor{ S , 1.0 , 1 } , parallel_execution_context("S","graph") );
    pq << flush;
    for(int rep=0;rep<3;rep++)
    {
      auto t0 = std::chrono::steady_clock::now();
      g.launch(); g.wait();
      const double t = std::chrono::duration<double,std::milli>( std::chrono::steady_clock::now() - t0 ).count();
      if( rep == 0 || t < ms[1] ) ms[1] = t;
    }
  }
  std::printf("graph: %d small operations : scheduled one by one %.3f ms (flush alone %.3f ms) , without timing events %.3f ms , one graph replay %.3f ms\n", nsmall, ms[0], submit_ms, untimed_ms, ms[1]);
  out.doubles( ms , 2 );
  out.doubles( S.data() , 8*256 );                             // every element was incremented 9*nsmall times
  return pq.empty() ? 0 : 3;
}

// config C4 through the GENERIC functor path, the way an application writes it: one block per cell, ONIKA_CU_BLOCK_SIMD_FOR over
// the cell's particles, block reduction, thread 0 stores.  Times block sizes 128 (reference default: 12.5 % of the lanes busy
// with 16 particles per cell) and 32 (opts.max_block_size) on a field-major column; prints microseconds per sweep.
template<int BS>
struct GenericCellSumFunctor
{
  const double * __restrict__ values; const uint64_t * __restrict__ cell_begin; double * __restrict__ cell_sum;
  ONIKA_HOST_DEVICE_FUNC inline void operator () ( size_t c ) const
  {
    const uint64_t b = cell_begin[c], e = cell_begin[c+1];
    double s = 0.0;
    ONIKA_CU_BLOCK_SIMD_FOR(uint64_t, p, b, e) { s += values[p]; }
    s = onika::cuda::block_reduce_add( s , onika::IntConst<BS>{} );
    if( ONIKA_CU_THREAD_IDX == 0 ) cell_sum[c] = s;
  }
};
namespace onika { namespace parallel { template<int BS> struct BlockParallelForFunctorTraits< GenericCellSumFunctor<BS> > { static inline constexpr bool CudaCompatible = true; }; } }

// the same functor source, declared safe for sub-block packing (no function-scope shared state): several cells per CTA
template<int BS> struct PackedCellSumFunctor : public GenericCellSumFunctor<BS> {};
namespace onika { namespace parallel { template<int BS> struct BlockParallelForFunctorTraits< PackedCellSumFunctor<BS> >
  { static inline constexpr bool CudaCompatible = true; static inline constexpr bool SubBlockItems = true; }; } }

static int scenario_cells_generic_perf(Out& out, size_t ncells)
{
  std::vector<uint64_t> begin(ncells+1,0);
  for(size_t c=0;c<ncells;c++) begin[c+1] = begin[c] + uint64_t( lcg(c,42,8.0,25.0) );     // ~16 particles per cell
  const size_t np = begin[ncells];
  void *pv=nullptr, *pb=nullptr, *ps=nullptr;
  ONIKA_B200_CHECK( onk_alloc( np*8 , 256 , ONK_MEM_DEVICE , &pv ) ); ONIKA_B200_CHECK( onk_alloc( (ncells+1)*8 , 256 , ONK_MEM_DEVICE , &pb ) );
  ONIKA_B200_CHECK( onk_alloc( ncells*8 , 256 , ONK_MEM_DEVICE , &ps ) );
  { std::vector<double> v(np); for(size_t i=0;i<np;i++) v[i] = lcg(i,40,-1.0,1.0);
    ONIKA_B200_CHECK( onk_memcpy( pv , v.data() , np*8 , 0 ) ); ONIKA_B200_CHECK( onk_memcpy( pb , begin.data() , (ncells+1)*8 , 0 ) ); ONIKA_B200_CHECK( onk_lane_sync(0) ); }
  auto & pq = ParallelExecutionQueue::default_queue();
  void* stream = nullptr; ONIKA_B200_CHECK( onk_lane_stream( 0 , &stream ) );
  cudaEvent_t e0, e1; ONIKA_CU_CHECK_ERRORS( cudaEventCreate(&e0) ); ONIKA_CU_CHECK_ERRORS( cudaEventCreate(&e1) );
  double us[8] = {0,0,0,0,0,0,0,0};
  auto timeit = [&](int slot, auto make_op)
  {
    for(int i=0;i<2;i++) pq << set_lane(0) << make_op();
    pq << flush << synchronize;
    ONIKA_CU_CHECK_ERRORS( cudaEventRecord( e0 , (cudaStream_t) stream ) );
    for(int i=0;i<5;i++) pq << set_lane(0) << make_op();
    pq << flush;
    ONIKA_CU_CHECK_ERRORS( cudaEventRecord( e1 , (cudaStream_t) stream ) );
    pq << synchronize;
    float ms = 0; ONIKA_CU_CHECK_ERRORS( cudaEventElapsedTime( &ms , e0 , e1 ) );
    us[slot] = ms * 1e3 / 5;
  };
  ParallelExecutionContext::s_gpu_timing = false;
  GenericCellSumFunctor<128> f128 = { (const double*)pv , (const uint64_t*)pb , (double*)ps };
  GenericCellSumFunctor<32>  f32  = { (const double*)pv , (const uint64_t*)pb , (double*)ps };
  BlockParallelForOptions o32; o32.max_block_size = 32;
  timeit( 0 , [&]{ return block_parallel_for( ncells , f128 , parallel_execution_context("cells","bs128") ); } );
  std::vector<double> r128(ncells); ONIKA_B200_CHECK( onk_memcpy( r128.data() , ps , ncells*8 , 0 ) ); ONIKA_B200_CHECK( onk_lane_sync(0) );
  timeit( 1 , [&]{ return block_parallel_for( ncells , f32 , parallel_execution_context("cells","bs32") , o32 ); } );
  std::vector<double> r32(ncells); ONIKA_B200_CHECK( onk_memcpy( r32.data() , ps , ncells*8 , 0 ) ); ONIKA_B200_CHECK( onk_lane_sync(0) );
  { // typed field-major kernel on the same data
    for(int i=0;i<2;i++) ONIKA_B200_CHECK( onk_cell_sum_packed_f64( (const double*)pv , np , (const uint64_t*)pb , ncells , (double*)ps , nullptr , 0 ) );
    ONIKA_CU_CHECK_ERRORS( cudaEventRecord( e0 , (cudaStream_t) stream ) );
    for(int i=0;i<5;i++) ONIKA_B200_CHECK( onk_cell_sum_packed_f64( (const double*)pv , np , (const uint64_t*)pb , ncells , (double*)ps , nullptr , 0 ) );
    ONIKA_CU_CHECK_ERRORS( cudaEventRecord( e1 , (cudaStream_t) stream ) );
    ONIKA_B200_CHECK( onk_lane_sync(0) );
    float ms = 0; ONIKA_CU_CHECK_ERRORS( cudaEventElapsedTime( &ms , e0 , e1 ) ); us[2] = ms * 1e3 / 5;
  }
  std::vector<double> rt(ncells); ONIKA_B200_CHECK( onk_memcpy( rt.data() , ps , ncells*8 , 0 ) ); ONIKA_B200_CHECK( onk_lane_sync(0) );
  // sub-block items: the same operator() in CTAs of 256/bs sub-blocks of bs threads
  std::vector<double> rsub[5]; for(auto& r : rsub) r.resize( ncells );
  auto sub = [&](int slot, auto functor, unsigned int bs)
  {
    ONIKA_B200_CHECK( onk_memset_bytes( ps , 0xff , ncells*8 , 0 ) );
    BlockParallelForOptions o; o.max_block_size = bs;
    timeit( 3+slot , [&]{ return block_parallel_for( ncells , functor , parallel_execution_context("cells","sub") , o ); } );
    ONIKA_B200_CHECK( onk_memcpy( rsub[slot].data() , ps , ncells*8 , 0 ) ); ONIKA_B200_CHECK( onk_lane_sync(0) );
  };
  sub( 0 , PackedCellSumFunctor<32>{ { (const double*)pv , (const uint64_t*)pb , (double*)ps } } , 32 );
  sub( 1 , PackedCellSumFunctor<16>{ { (const double*)pv , (const uint64_t*)pb , (double*)ps } } , 16 );
  sub( 2 , PackedCellSumFunctor<8>{ { (const double*)pv , (const uint64_t*)pb , (double*)ps } } , 8 );
  // even smaller sub-blocks: more cells in flight per SM (each item is a chain of dependent loads: the path is latency-bound)
  sub( 3 , PackedCellSumFunctor<4>{ { (const double*)pv , (const uint64_t*)pb , (double*)ps } } , 4 );
  sub( 4 , PackedCellSumFunctor<2>{ { (const double*)pv , (const uint64_t*)pb , (double*)ps } } , 2 );
  ParallelExecutionContext::s_gpu_timing = true;
  std::printf("cells_generic_perf ncells=%zu particles=%zu : generic block 128 %.1f us , generic block 32 %.1f us , typed field-major %.1f us , sub-block items 32/16/8/4/2 threads %.1f / %.1f / %.1f / %.1f / %.1f us\n",
              ncells, np, us[0], us[1], us[2], us[3], us[4], us[5], us[6], us[7]);
  out.doubles( us , 6 );
  out.doubles( rt.data() , ncells );            // in-order sums (typed kernel): the reference result
  out.doubles( r128.data() , ncells ); out.doubles( r32.data() , ncells );   // tree sums of the generic functor: equal within rounding
  for(int k=0;k<3;k++) out.doubles( rsub[k].data() , ncells );
  out.doubles( us+6 , 2 );                      // appended in round 2: sub-blocks of 4 and 2 threads
  for(int k=3;k<5;k++) out.doubles( rsub[k].data() , ncells );
  onk_free( pv , np*8 ); onk_free( pb , (ncells+1)*8 ); onk_free( ps , ncells*8 );
  return pq.empty() ? 0 : 3;
}

static int scenario_grid3d(Out& out)
{
  { // which cells of a second 2D sweep must wait for a first one, given a cross-shaped conflict stencil (host logic)
    using C2 = onika::oarray_t<ssize_t,2>;
    const ParallelExecutionSpace<2> first = { {0,0} , {4,4} } , second = { {3,0} , {7,2} };
    std::vector<C2> stencil = { C2{0,0} , C2{-1,0} , C2{1,0} , C2{0,-1} , C2{0,1} } , dep , rem;
    concurrent_execution_space_elements( first , second , std::span<C2>( stencil ) , dep , rem );
    // column i=3 lies inside `first`, column i=4 touches it through the (-1,0) offset, i=5,6 are independent
    if( dep.size() != 4 || rem.size() != 4 ) return 7;
    for(const auto& c : dep) if( c[0] > 4 ) return 8;
    for(const auto& c : rem) if( c[0] < 5 ) return 9;
  }
  {
    const ssize_t N = 4;
    CudaMMVector<double> m( N*N*N , 0.0 );
    Grid3DBlockFunctor f = { m.data() , N };
    ParallelExecutionSpace<3> r = { {0,0,0} , {N,N,N} };
    block_parallel_for( r , f , parallel_execution_context("grid3d","block") );
    out.doubles( m.data() , m.size() );
  }
  {
    const ssize_t N = 16;
    CudaMMVector<double> m( N*N*N , 0.0 );
    Grid3DFunctor f = { m.data() , N };
    ParallelExecutionSpace<3> r = { {2,3,3} , {N-1,N-3,N-3} };
    parallel_for( r , f , parallel_execution_context("grid3d","pfor") );
    out.doubles( m.data() , m.size() );
  }
  return 0;
}

static int scenario_reduce(Out& out, size_t rows, size_t cols)
{
  CudaMMVector<double> d( rows*cols );
  fill(d,17,-1.0,1.0);
  ReduceResult res = { DBL_MAX , -DBL_MAX , 0.0 , 0ull };        // initial value travels to the device, result travels back
  auto* pec = parallel_execution_context("reduce");
  pec->init_device_scratch();                                     // the functor needs the device return-data pointer
  RowReduceFunctor f = { d.data() , cols , (ReduceResult*) pec->get_device_return_data_ptr() };
  BlockParallelForOptions o; o.return_data = &res; o.return_data_size = sizeof(res);
  int cb_fired = 0;
  o.user_cb = { [](void* p){ ++ *static_cast<int*>(p); } , &cb_fired };
  block_parallel_for( rows , f , pec , o );
  const double r[5] = { res.vmin , res.vmax , res.vsum , double(res.count) , double(cb_fired) };
  out.doubles( r , 5 );
  return 0;
}

static int scenario_cells(Out& out, size_t ncells)
{
  // sizes: deterministic pseudo-Poisson-ish in [0,64]
  std::vector<uint32_t> counts(ncells);
  size_t total = 0;
  for(size_t c=0;c<ncells;c++) { counts[c] = uint32_t( lcg(c,42,0.0,33.0) ); if( c % 97 == 0 ) counts[c] = 64; if( c % 89 == 5 ) counts[c] = 0; total += counts[c]; }
  onika::memory::CudaMMArray<CellArrays> cells; cells.resize(ncells);
  for(size_t c=0;c<ncells;c++) new( &cells[c] ) CellArrays();
  size_t k = 0, bytes = 0;
  for(size_t c=0;c<ncells;c++)
  {
    cells[c].resize( counts[c] );
    for(size_t p=0;p<counts[c];p++,k++) cells[c][particle_e][p] = lcg(k,40,-1.0,1.0);
    bytes += cells[c].storage_size();
  }
  CudaMMVector<double> sums( ncells , 0.0 );
  CudaMMVector<double> tot( 1 , 0.0 );
  CellSumFunctor f = { cells.m_data_pointer , sums.data() , tot.data() };
  block_parallel_for( ncells , f , parallel_execution_context("cells") );
  out.u64( ncells ); out.u64( total ); out.u64( bytes );
  std::vector<double> cnt( counts.begin() , counts.end() );
  out.doubles( cnt.data() , ncells );
  out.doubles( sums.data() , ncells );
  out.doubles( tot.data() , 1 );
  for(size_t c=0;c<ncells;c++) cells[c].clear();
  return 0;
}

static int scenario_listed(Out& out)
{
  const size_t n = 1000;
  CudaMMVector<double> d(n); fill(d,3,1e-3,1.0);
  CudaMMVector<ssize_t> idx; for(size_t i=0;i<n;i+=3) idx.push_back( ssize_t(i) );
  int prolog = 0, epilog = 0;
  ScaleListedFunctor f = { d.data() , 2.0 , &prolog , &epilog };
  ParallelExecutionSpace<1,1> space = { {0} , { ssize_t(idx.size()) } , std::span<const ssize_t>( idx.data() , idx.size() ) };
  block_parallel_for( space , f , parallel_execution_context("listed") );
  out.doubles( d.data() , n );
  const double pe[2] = { double(prolog) , double(epilog) };
  out.doubles( pe , 2 );
  return 0;
}

// SOATL: layout, resize policy, tuple access, serialisation, device parallel_apply_simd
struct DistFunctor
{
  double ax, ay, az;
  ONIKA_HOST_DEVICE_FUNC inline void operator () ( double& d, double x, double y, double z ) const
  {
    x = x - ax; y = y - ay; z = z - az;
    d = sqrt( x*x + y*y + z*z );
    x /= d; y /= d; z /= d;
    d += x/(y*z);
  }
};
namespace onika { namespace soatl { template<> struct ApplyFunctorTraits<DistFunctor> { static inline constexpr bool CudaCompatible = true; }; } }

static int scenario_soatl(Out& out, size_t n)
{
  using namespace onika::soatl;
  // --- tests/soatupletest.cpp style static checks ---
  static_assert( find_index_of_id<_tag_particle_rx,_tag_particle_rx,_tag_particle_ry>::index == 0 );
  static_assert( find_index_of_id<_tag_particle_ry,_tag_particle_rx,_tag_particle_ry>::index == 1 );
  static_assert( find_index_of_id<_tag_particle_e,_tag_particle_rx,_tag_particle_ry>::index == bad_field_index );
  // --- layout probe: <64,16,rx,ry,rz,atype,e,mid>, resize sequence ---
  {
    auto fa = make_hybrid_field_arrays( cst::align<64>() , cst::chunk<16>() , particle_rx , particle_ry , particle_rz , particle_atype , particle_e , particle_mid );
    const size_t seq[] = { 17 , 33 , 10 , 0 , 100 , 40 };
    for(size_t s : seq)
    {
      fa.resize(s);
      out.u64( fa.capacity() ); out.u64( fa.storage_size() );
      const uint8_t* base = (const uint8_t*) fa.storage_ptr();
      const uint8_t* p[6] = { (uint8_t*) fa[particle_rx] , (uint8_t*) fa[particle_ry] , (uint8_t*) fa[particle_rz] , (uint8_t*) fa[particle_atype] , (uint8_t*) fa[particle_e] , (uint8_t*) fa[particle_mid] };
      for(int k=0;k<6;k++) out.u64( base ? uint64_t( p[k] - base ) : 0 );
    }
    fa.clear();
    if( fa.storage_ptr() != nullptr || fa.capacity() != 0 ) return 4;
  }
  // --- tuple access, growth keeps contents ---
  {
    auto fa = make_hybrid_field_arrays( cst::align<64>() , cst::chunk<8>() , particle_e , particle_atype , particle_rx , particle_mid , particle_ry , particle_rz );
    for(size_t i=0;i<n;i++) fa.push_back( fa.make_tuple( lcg(i,30,1e-3,1), (unsigned char)( lcg(i,34,0,50) ), lcg(i,31,1e-3,1), int32_t( lcg(i,35,0,500) ), lcg(i,32,1e-3,1), lcg(i,33,1e-3,1) ) );
    if( fa.size() != n ) return 5;
    auto t = fa[ n/2 ];
    if( t[particle_rx] != lcg(n/2,31,1e-3,1) || t[particle_mid] != int32_t( lcg(n/2,35,0,500) ) ) return 6;
    // serialise into the <1,1> wire layout (tests/soatlserializetest.cpp:86-104)
    auto wire = make_hybrid_field_arrays( cst::align<1>() , cst::chunk<1>() , particle_e , particle_atype , particle_rx , particle_mid , particle_ry , particle_rz );
    wire.resize(n);
    copy( wire , fa , 0 , n/2 );
    copy( wire , fa , n/2 , n - n/2 , particle_rx , particle_ry );
    copy( wire , fa , particle_rz , particle_e , particle_atype , particle_mid );
    copy( wire , fa );
    out.u64( wire.storage_size() );
    out.bytes( wire.storage_ptr() , wire.storage_size() );
    // value tuples and pointer bundles (tests/soatupletest.cpp idioms): zero by default, field-wise conversion, views
    {
      FieldTuple<_tag_particle_rx,_tag_particle_mid,_tag_particle_atype> z;
      if( z[particle_rx] != 0.0 || z[particle_mid] != 0 || z[particle_atype] != 0 ) return 7;
      FieldTuple<_tag_particle_rx,_tag_particle_e,_tag_particle_mid> part( t );            // rx, e, mid from the 6-field element
      if( part[particle_rx] != t[particle_rx] || part[particle_e] != t[particle_e] || part[particle_mid] != t[particle_mid] ) return 8;
      FieldTuple<_tag_particle_rx,_tag_particle_rz> two( 7.0 , 9.0 );
      FieldTuple<_tag_particle_rz,_tag_particle_ry> conv( two );                           // rz copied, ry absent in the source -> 0
      if( conv[particle_rz] != 9.0 || conv[particle_ry] != 0.0 ) return 9;
      conv[particle_ry] = 5.0; conv.copy_existing_fields( FieldTuple<_tag_particle_rz>( 1.0 ) );
      if( conv[particle_rz] != 1.0 || conv[particle_ry] != 5.0 || ! ( conv == conv ) || conv != conv ) return 10;
      int nfields = 0; double acc = 0.0;
      two.apply_fields( [&](auto , double v) { ++nfields; acc += v; } );
      if( nfields != 2 || acc != 16.0 ) return 11;
      static_assert( field_tuple_has_field_v< decltype(two) , _tag_particle_rz > && ! field_tuple_has_field_v< decltype(two) , _tag_particle_e > );
      static_assert( std::is_same_v< FieldTupleFromFieldIds< FieldIds<_tag_particle_rx,_tag_particle_rz> > , decltype(two) > );

      auto view = make_field_arrays_view( fa , particle_rz , particle_e );
      if( view.size() != n || view[particle_rz] != fa[particle_rz] || view[particle_e] != fa[particle_e] ) return 12;
      auto nul = make_field_pointer_tuple( cst::align<64>() , cst::chunk<8>() , particle_rz , particle_ry );
      if( nul[particle_rz] != nullptr || nul[particle_ry] != nullptr ) return 13;
      nul.copy_or_zero_fields( view );
      if( nul[particle_rz] != fa[particle_rz] || nul[particle_ry] != nullptr || ! ( nul == nul ) ) return 14;
      std::ostringstream desc;
      pretty_print_arrays( desc , fa );
      if( desc.str().find("contains 6 fields") == std::string::npos || desc.str().find("Particle position Z") == std::string::npos ) return 16;
      const auto fids = onika::make_flat_tuple( particle_rx , particle_atype , particle_mid );       // 8 + 1 + 4 bytes, each padded to 8
      if( field_id_tuple_size_bytes( fids ) != 24 || field_id_size_bytes( particle_atype , 4 ) != 4 ) return 17;
      onika::cuda::pair<int,double> pa{ 1 , 2.0 }, pb{ 1 , 3.0 };
      const double sorted[5] = { 1.0 , 2.0 , 2.0 , 4.0 , 8.0 };
      if( ! ( pa < pb ) || pa == pb || onika::cuda::lower_bound( sorted , sorted+5 , 2.0 ) != sorted+1 || onika::cuda::lower_bound( sorted , sorted+5 , 9.0 ) != sorted+5 ) return 15;
    }
    wire.clear(); fa.clear();
  }
  // --- tests/benchmark.cpp kernel through parallel_apply_simd with a device lambda ---
  {
    auto arrays = make_hybrid_field_arrays( cst::align<64>() , cst::chunk<16>() , particle_rx , particle_ry , particle_rz , particle_e );
    arrays.resize(n);
    for(size_t i=0;i<arrays.capacity();i++) { arrays[particle_rx][i]=2.0; arrays[particle_ry][i]=3.0; arrays[particle_rz][i]=4.0; arrays[particle_e][i]=0.0; }
    for(size_t i=0;i<n;i++) { arrays[particle_rx][i]=lcg(i,10,0,1); arrays[particle_ry][i]=lcg(i,11,0,1); arrays[particle_rz][i]=lcg(i,12,0,1); }
    const double ax = arrays[particle_rx][0], ay = arrays[particle_ry][0], az = arrays[particle_rz][0];
    parallel_apply_simd( [ax,ay,az] ONIKA_HOST_DEVICE_FUNC (double& d, double x, double y, double z)
      {
        x = x - ax; y = y - ay; z = z - az;
        d = sqrt( x*x + y*y + z*z );
        x /= d; y /= d; z /= d;
        d += x/(y*z);
      } , arrays , particle_e , particle_rx , particle_ry , particle_rz );
    out.doubles( arrays[particle_e] , n );
    // the same sweep by a named functor type that declares itself device-compatible (ApplyFunctorTraits)
    for(size_t i=0;i<arrays.capacity();i++) arrays[particle_e][i] = 0.0;
    parallel_apply_simd( DistFunctor{ ax , ay , az } , arrays , particle_e , particle_rx , particle_ry , particle_rz );
    out.doubles( arrays[particle_e] , n );
    arrays.clear();
  }
  return 0;
}

// soatl::parallel_apply_simd on the device, the two C3 kernels: the 8-field read-modify-write sweep and the
// tests/benchmark.cpp operator, through the GENERIC vectorised kernel of soatl/compute.h, next to the typed entry points
// of the library on the same arrays.  Small n: every element goes to the output (parity); timed = 1: prints/returns times.
static int scenario_soatl_apply(Out& out, size_t n, bool timed)
{
  using namespace onika::soatl;
  void* stream = nullptr; ONIKA_B200_CHECK( onk_lane_stream( 0 , &stream ) );
  cudaEvent_t e0, e1; ONIKA_CU_CHECK_ERRORS( cudaEventCreate(&e0) ); ONIKA_CU_CHECK_ERRORS( cudaEventCreate(&e1) );
  auto time_us = [&](auto launch, int reps) -> double
  {
    for(int i=0;i<2;i++) launch();
    ONIKA_CU_CHECK_ERRORS( cudaEventRecord( e0 , (cudaStream_t) stream ) );
    for(int i=0;i<reps;i++) launch();
    ONIKA_CU_CHECK_ERRORS( cudaEventRecord( e1 , (cudaStream_t) stream ) );
    ONIKA_B200_CHECK( onk_lane_sync(0) );
    float ms = 0; ONIKA_CU_CHECK_ERRORS( cudaEventElapsedTime( &ms , e0 , e1 ) );
    return ms * 1e3 / reps;
  };
  double us[4] = {0,0,0,0};
  const double s = 1.0000001; const double c[8] = { 0.125 , 0.25 , 0.375 , 0.5 , 0.625 , 0.75 , 0.875 , 1.0 };
  {
    auto fa = make_hybrid_field_arrays( cst::align<256>() , cst::chunk<16>() , bench_f0 , bench_f1 , bench_f2 , bench_f3 , bench_f4 , bench_f5 , bench_f6 , bench_f7 );
    fa.resize(n);
    double* col[8] = { fa[bench_f0] , fa[bench_f1] , fa[bench_f2] , fa[bench_f3] , fa[bench_f4] , fa[bench_f5] , fa[bench_f6] , fa[bench_f7] };
    if( ! timed ) { for(int k=0;k<8;k++) for(size_t i=0;i<fa.capacity();i++) col[k][i] = ( i < n ) ? lcg(i,20+k,0.0,1.0) : 0.5; }
    else ONIKA_B200_CHECK( onk_memset_bytes( fa.storage_ptr() , 0 , fa.storage_size() , 0 ) );
    ONIKA_B200_CHECK( onk_prefetch( fa.storage_ptr() , fa.storage_size() , 1 , 0 ) ); ONIKA_B200_CHECK( onk_lane_sync(0) );
    const double c0=c[0],c1=c[1],c2=c[2],c3=c[3],c4=c[4],c5=c[5],c6=c[6],c7=c[7];
    auto rmw = [=] ONIKA_HOST_DEVICE_FUNC (double& a0, double& a1, double& a2, double& a3, double& a4, double& a5, double& a6, double& a7)
      { a0 = a0*s + c0; a1 = a1*s + c1; a2 = a2*s + c2; a3 = a3*s + c3; a4 = a4*s + c4; a5 = a5*s + c5; a6 = a6*s + c6; a7 = a7*s + c7; };
    auto sweep = [&]{ parallel_apply_simd( rmw , fa , bench_f0 , bench_f1 , bench_f2 , bench_f3 , bench_f4 , bench_f5 , bench_f6 , bench_f7 ); };
    if( ! timed )
    {
      sweep();                                                   // default target: lane 0, synchronous like the reference's loop
      for(int k=0;k<8;k++) out.doubles( col[k] , n );
    }
    else
    {
      DeviceApplyScope on_lane( 0 , false );                    // caller's lane, no host synchronisation between sweeps
      us[0] = time_us( sweep , 5 );
      us[1] = time_us( [&]{ ONIKA_B200_CHECK( onk_soatl_rmw_f64( fa.storage_ptr() , 256 , 16 , 8 , n , fa.capacity() , s , c , 0 ) ); } , 5 );
    }
    fa.clear();
  }
  {
    auto arrays = make_hybrid_field_arrays( cst::align<256>() , cst::chunk<16>() , particle_rx , particle_ry , particle_rz , particle_e );
    arrays.resize(n);
    if( ! timed )
      for(size_t i=0;i<arrays.capacity();i++)
      { const bool in = i < n; arrays[particle_rx][i] = in ? lcg(i,10,0,1) : 2.0; arrays[particle_ry][i] = in ? lcg(i,11,0,1) : 3.0; arrays[particle_rz][i] = in ? lcg(i,12,0,1) : 4.0; arrays[particle_e][i] = 0.0; }
    else
    {
      ONIKA_B200_CHECK( onk_parallel_memset( arrays[particle_rx] , arrays.capacity() , 0x3fe0000000000000ull , 8 , 0 ) );   // 0.5
      ONIKA_B200_CHECK( onk_parallel_memset( arrays[particle_ry] , arrays.capacity() , 0x3fd0000000000000ull , 8 , 0 ) );   // 0.25
      ONIKA_B200_CHECK( onk_parallel_memset( arrays[particle_rz] , arrays.capacity() , 0x3fe8000000000000ull , 8 , 0 ) );   // 0.75
    }
    ONIKA_B200_CHECK( onk_prefetch( arrays.storage_ptr() , arrays.storage_size() , 1 , 0 ) ); ONIKA_B200_CHECK( onk_lane_sync(0) );
    const double ax = timed ? 0.125 : arrays[particle_rx][0], ay = timed ? 0.0625 : arrays[particle_ry][0], az = timed ? 0.03125 : arrays[particle_rz][0];
    auto dist = [ax,ay,az] ONIKA_HOST_DEVICE_FUNC (double& d, double x, double y, double z)
      {
        x = x - ax; y = y - ay; z = z - az;
        d = sqrt( x*x + y*y + z*z );
        x /= d; y /= d; z /= d;
        d += x/(y*z);
      };
    auto sweep = [&]{ parallel_apply_simd( dist , arrays , particle_e , particle_rx , particle_ry , particle_rz ); };
    if( ! timed ) { sweep(); out.doubles( arrays[particle_e] , n ); }
    else
    {
      DeviceApplyScope on_lane( 0 , false );
      us[2] = time_us( sweep , 5 );
      const size_t npad = ( ( n + 15 ) / 16 ) * 16;
      us[3] = time_us( [&]{ ONIKA_B200_CHECK( onk_soatl_dist_f64( arrays[particle_e] , arrays[particle_rx] , arrays[particle_ry] , arrays[particle_rz] , npad , ax , ay , az , 0 ) ); } , 5 );
    }
    arrays.clear();
  }
  if( timed )
  {
    std::printf("soatl_apply n=%zu : 8-field rmw generic %.1f us (%.0f GB/s) , typed %.1f us (%.0f GB/s) ; benchmark.cpp operator generic %.1f us (%.0f GB/s) , typed %.1f us (%.0f GB/s)\n",
                n, us[0], 128.0*n/us[0]/1e3, us[1], 128.0*n/us[1]/1e3, us[2], 32.0*n/us[2]/1e3, us[3], 32.0*n/us[3]/1e3);
    out.doubles( us , 4 );
  }
  return 0;
}

// mixed column types through the vectorised apply kernel (E = 4 elements per thread: 32-byte, 16-byte, 4-byte accesses side by
// side), read-only columns by value and by const reference, sub-ranges that start off the vector alignment (scalar kernel),
// sizes below one tile; every device result next to the same operator run by the host loop
static int scenario_soatl_apply_mixed(Out& out, size_t n)
{
  using namespace onika::soatl;
  auto make = [&](auto& fa)
  {
    fa.resize(n);
    for(size_t i=0;i<fa.capacity();i++)
    { fa[particle_rx][i] = lcg(i,50,-1.0,1.0); fa[particle_atype][i] = (unsigned char)( lcg(i,51,0.0,200.0) ); fa[particle_mid][i] = int32_t( lcg(i,52,-1000.0,1000.0) ); fa[particle_dist][i] = float( lcg(i,53,0.0,4.0) ); }
  };
  auto dev = make_hybrid_field_arrays( cst::align<64>() , cst::chunk<16>() , particle_rx , particle_atype , particle_mid , particle_dist );
  auto host = make_hybrid_field_arrays( cst::align<64>() , cst::chunk<16>() , particle_rx , particle_atype , particle_mid , particle_dist );
  make(dev); make(host);
  const double k = 0.25;
  auto op = [k] ONIKA_HOST_DEVICE_FUNC (double& x, unsigned char t, int32_t& m, const float& d)
    { x = x * k + double(t); m = m * 2 + int32_t(t) - int32_t(d); };                    // writes x and m, reads t (by value) and d (const reference)
  parallel_apply( op , dev , particle_rx , particle_atype , particle_mid , particle_dist );           // device: vector kernel + scalar tail
  apply( op , host , particle_rx , particle_atype , particle_mid , particle_dist );                   // host loop: the reference's sequential semantics
  // a sub-range that starts at an odd element: columns are no longer vector-aligned -> one element per thread
  const size_t first = ( n > 7 ) ? 3 : 0 , len = ( n > 7 ) ? n - 5 : n;
  auto op2 = [] ONIKA_HOST_DEVICE_FUNC (float& d, double x) { d = d + float(x); };
  parallel_apply( op2 , first , len , dev , particle_dist , particle_rx );
  apply( op2 , first , len , host , particle_dist , particle_rx );
  out.doubles( dev[particle_rx] , n ); out.doubles( host[particle_rx] , n );
  out.bytes( dev[particle_atype] , n ); out.bytes( host[particle_atype] , n );
  out.bytes( dev[particle_mid] , n*4 ); out.bytes( host[particle_mid] , n*4 );
  out.bytes( dev[particle_dist] , n*4 ); out.bytes( host[particle_dist] , n*4 );
  dev.clear(); host.clear();
  return 0;
}

// sub-block items over an element LIST (indexed 1-D space) and with fewer items than one CTA holds sub-blocks
static int scenario_sub_block_listed(Out& out, size_t ncells)
{
  std::vector<uint64_t> begin(ncells+1,0);
  for(size_t c=0;c<ncells;c++) begin[c+1] = begin[c] + uint64_t( lcg(c,42,0.0,25.0) );          // 0..24 particles per cell, empty cells included
  const size_t np = begin[ncells];
  CudaMMVector<double> v( np > 0 ? np : 1 ); for(size_t i=0;i<np;i++) v[i] = lcg(i,40,-1.0,1.0);
  CudaMMVector<uint64_t> b( ncells+1 ); for(size_t c=0;c<=ncells;c++) b[c] = begin[c];
  CudaMMVector<double> sums( ncells , -7.0 );
  CudaMMVector<ssize_t> idx; for(size_t c=0;c<ncells;c+=2) idx.push_back( ssize_t(c) );         // every other cell
  PackedCellSumFunctor<16> f = { { v.data() , b.data() , sums.data() } };
  BlockParallelForOptions o; o.max_block_size = 16;
  ParallelExecutionSpace<1,1> space = { {0} , { ssize_t(idx.size()) } , std::span<const ssize_t>( idx.data() , idx.size() ) };
  block_parallel_for( space , f , parallel_execution_context("cells","listed") , o );
  out.doubles( sums.data() , ncells );
  // 5 items only: less than the 16 sub-blocks of one CTA
  CudaMMVector<double> few( ncells , -7.0 );
  PackedCellSumFunctor<16> g = { { v.data() , b.data() , few.data() } };
  block_parallel_for( std::min<size_t>( 5 , ncells ) , g , parallel_execution_context("cells","few") , o );
  out.doubles( few.data() , ncells );
  block_parallel_for( 0 , g , parallel_execution_context("cells","none") , o );                   // empty space: nothing runs
  // fixed grid + work stealing with sub-blocks: every sub-block draws its own tickets
  CudaMMVector<double> ws( ncells , -7.0 );
  PackedCellSumFunctor<16> h = { { v.data() , b.data() , ws.data() } };
  BlockParallelForOptions ow; ow.max_block_size = 16; ow.fixed_gpu_grid_size = true;
  block_parallel_for( ncells , h , parallel_execution_context("cells","worksteal") , ow );
  out.doubles( ws.data() , ncells );
  return 0;
}

// sub-block items over 3-D and 2-D boxes of cells (sub-box with a non-zero start): the same functor source once as ordinary
// 2D/3D blocks of the configured shape (8x8x4 / 8x8), once packed into CTAs of 16-thread sub-blocks
template<int BS>
struct GridCellSumFunctor
{
  const double * __restrict__ values; const uint64_t * __restrict__ cell_begin; double * __restrict__ cell_sum; ssize_t nx , ny;
  ONIKA_HOST_DEVICE_FUNC inline void operator () ( onikaInt3_t c ) const
  {
    const size_t cell = size_t( ( c.z * ny + c.y ) * nx + c.x );
    const uint64_t b = cell_begin[cell], e = cell_begin[cell+1];
    double s = 0.0;
    ONIKA_CU_BLOCK_SIMD_FOR(uint64_t, p, b, e) { s += values[p]; }
    s = onika::cuda::block_reduce_add( s , onika::IntConst<BS>{} );
    if( ONIKA_CU_THREAD_IDX == 0 ) cell_sum[cell] = s + double( ONIKA_CU_BLOCK_SIZE );       // the block size the functor saw is part of the result
  }
};
template<int BS> struct PackedGridCellSumFunctor : public GridCellSumFunctor<BS> {};
namespace onika { namespace parallel {
  template<int BS> struct BlockParallelForFunctorTraits< GridCellSumFunctor<BS> > { static inline constexpr bool CudaCompatible = true; };
  template<int BS> struct BlockParallelForFunctorTraits< PackedGridCellSumFunctor<BS> > { static inline constexpr bool CudaCompatible = true; static inline constexpr bool SubBlockItems = true; };
} }

static int scenario_sub_block_box(Out& out, ssize_t nx, ssize_t ny, ssize_t nz)
{
  const size_t ncells = size_t( nx * ny * nz );
  std::vector<uint64_t> begin(ncells+1,0);
  for(size_t c=0;c<ncells;c++) begin[c+1] = begin[c] + uint64_t( lcg(c,42,0.0,25.0) );
  const size_t np = begin[ncells];
  CudaMMVector<double> v( np > 0 ? np : 1 ); for(size_t i=0;i<np;i++) v[i] = lcg(i,40,-1.0,1.0);
  CudaMMVector<uint64_t> b( ncells+1 ); for(size_t c=0;c<=ncells;c++) b[c] = begin[c];
  BlockParallelForOptions small; small.max_block_size = 16;
  const ParallelExecutionSpace<3> box3 = { {1,0,2} , {nx,ny-1,nz} };
  const ParallelExecutionSpace<2> box2 = { {0,1} , {nx-1,ny} };
  {
    CudaMMVector<double> sums( ncells , -7.0 );
    GridCellSumFunctor<256> f = { v.data() , b.data() , sums.data() , nx , ny };
    block_parallel_for( box3 , f , parallel_execution_context("box3","blocks") );
    out.doubles( sums.data() , ncells );
  }
  {
    CudaMMVector<double> sums( ncells , -7.0 );
    PackedGridCellSumFunctor<16> f = { { v.data() , b.data() , sums.data() , nx , ny } };
    block_parallel_for( box3 , f , parallel_execution_context("box3","sub-blocks") , small );
    out.doubles( sums.data() , ncells );
  }
  {
    CudaMMVector<double> sums( ncells , -7.0 );
    GridCellSumFunctor<64> f = { v.data() , b.data() , sums.data() , nx , ny };
    block_parallel_for( box2 , f , parallel_execution_context("box2","blocks") );
    out.doubles( sums.data() , ncells );
  }
  {
    CudaMMVector<double> sums( ncells , -7.0 );
    PackedGridCellSumFunctor<16> f = { { v.data() , b.data() , sums.data() , nx , ny } };
    block_parallel_for( box2 , f , parallel_execution_context("box2","sub-blocks") , small );
    out.doubles( sums.data() , ncells );
  }
  {
    // the trait alone does not pack: without a small max_block_size the configured 3-D tile (256 threads) stays
    CudaMMVector<double> sums( ncells , -7.0 );
    PackedGridCellSumFunctor<256> f = { { v.data() , b.data() , sums.data() , nx , ny } };
    block_parallel_for( box3 , f , parallel_execution_context("box3","trait-only") );
    out.doubles( sums.data() , ncells );
  }
  return 0;
}

// a HOST-only functor (CudaCompatible = false) that keeps one accumulator per "block", indexed by ONIKA_CU_BLOCK_IDX and sized by
// the grid: on the host the grid is the OpenMP team and the block index the thread's rank (the reference's meaning,
// cuda.h:221-224 there), so the plain `+=` below is race-free only if every thread really sees its own index
struct HostBlockScratchFunctor
{
  const double * values; double * per_block; unsigned int * max_grid; unsigned int * bad;
  inline void operator () ( size_t i ) const
  {
    const unsigned int b = ONIKA_CU_BLOCK_IDX , g = ONIKA_CU_GRID_SIZE;
    if( b >= g || b >= 1024u || ONIKA_CU_BLOCK_SIZE != 1 || ONIKA_CU_THREAD_IDX != 0 ) { ONIKA_CU_ATOMIC_ADD( *bad , 1u ); return; }
    per_block[ b * 8 ] += values[i];                      // 64-byte stride: one cache line per block
    ONIKA_CU_ATOMIC_MAX( *max_grid , g );
  }
};
namespace onika { namespace parallel { template<> struct BlockParallelForFunctorTraits<HostBlockScratchFunctor> { static inline constexpr bool CudaCompatible = false; }; } }

struct HostOwnerFunctor { double * owner; inline void operator () ( size_t i ) const { owner[i] = double( ONIKA_CU_BLOCK_IDX ); } };
namespace onika { namespace parallel { template<> struct BlockParallelForFunctorTraits<HostOwnerFunctor> { static inline constexpr bool CudaCompatible = false; }; } }

static int scenario_host_block_scratch(Out& out, size_t n)
{
  std::vector<double> v(n); for(size_t i=0;i<n;i++) v[i] = double( int( lcg(i,77,0.0,1000.0) ) );      // integers: any summation order is exact
  std::vector<double> per_block( 1024 * 8 , 0.0 );
  unsigned int max_grid = 0 , bad = 0;
  HostBlockScratchFunctor f = { v.data() , per_block.data() , &max_grid , &bad };
  block_parallel_for( n , f , parallel_execution_context("host","scratch") );      // synchronous: runs as a host node of lane 0
  double total = 0.0 , expect = 0.0; unsigned int used = 0;
  for(size_t b=0;b<1024;b++) { total += per_block[b*8]; used += ( per_block[b*8] != 0.0 ); }
  for(size_t i=0;i<n;i++) expect += v[i];
  const double r[5] = { total , expect , double(max_grid) , double(used) , double(bad) };
  std::printf("host_block_scratch: total %.0f expect %.0f team %u blocks used %u bad %u\n", total, expect, max_grid, used, bad);
  out.doubles( r , 5 );
  // opts.omp_scheduling reaches the host loop: with OMP_SCHED_STATIC thread t owns one contiguous run of items
  std::vector<double> owner( 4096 , -1.0 );
  HostOwnerFunctor g = { owner.data() };
  BlockParallelForOptions o; o.omp_scheduling = OMP_SCHED_STATIC;
  block_parallel_for( owner.size() , g , parallel_execution_context("host","static") , o );
  out.doubles( owner.data() , owner.size() );
  return 0;
}

// reduction helpers queued on lanes together with the kernels that produce their input
static int scenario_reduce_helpers(Out& out, size_t rows, size_t cols)
{
  extras::Array2D a1, a2;
  a1.resize(rows,cols); a2.resize(rows,cols);
  fill(a1.m_data,8,-1.0,1.0); fill(a2.m_data,9,-1.0,1.0);
  extras::BlockParallelValueAddFunctor<true> k1 = { a1 , 1.1 } , k2 = { a2 , 1.3 };
  double r[6] = { 0,0,0,0,0,0 };
  auto & pq = ParallelExecutionQueue::default_queue();
  pq << set_lane(0) << block_parallel_for( rows , k1 , parallel_execution_context("a1","add") )
                    << parallel_reduce_min( a1.data() , rows*cols , &r[0] , parallel_execution_context("a1","min") )
                    << parallel_reduce_max( a1.data() , rows*cols , &r[1] , parallel_execution_context("a1","max") )
                    << parallel_reduce_sum( a1.data() , rows*cols , &r[2] , parallel_execution_context("a1","sum") )
     << set_lane(1) << block_parallel_for( rows , k2 , parallel_execution_context("a2","add") )
                    << parallel_reduce_min( a2.data() , rows*cols , &r[3] , parallel_execution_context("a2","min") )
                    << parallel_reduce_max( a2.data() , rows*cols , &r[4] , parallel_execution_context("a2","max") )
                    << parallel_reduce_sum( a2.data() , rows*cols , &r[5] , parallel_execution_context("a2","sum") )
     << flush << synchronize;
  out.doubles( r , 6 );
  // synchronous form on an empty range: identities
  double e[3] = { 1.0 , 1.0 , 1.0 };
  parallel_reduce_min( a1.data() , 0 , &e[0] , parallel_execution_context("empty","min") );
  parallel_reduce_max( a1.data() , 0 , &e[1] , parallel_execution_context("empty","max") );
  parallel_reduce_sum( a1.data() , 0 , &e[2] , parallel_execution_context("empty","sum") );
  out.doubles( e , 3 );
  return 0;
}

// device-timed throughput of the generic functor path and of the typed fast paths, through the public API
static int scenario_perf(Out&, size_t n)
{
  void* stream = nullptr;
  ONIKA_B200_CHECK( onk_lane_stream( 0 , &stream ) );
  cudaEvent_t e0, e1;
  ONIKA_CU_CHECK_ERRORS( cudaEventCreate(&e0) ); ONIKA_CU_CHECK_ERRORS( cudaEventCreate(&e1) );
  auto & pq = ParallelExecutionQueue::default_queue();
  auto bench = [&](const char* name, double bytes, auto make_op)
  {
    const int reps = 10;
    for(int i=0;i<3;i++) { pq << set_lane(0) << make_op(); } pq << flush << synchronize;
    ONIKA_CU_CHECK_ERRORS( cudaEventRecord( e0 , (cudaStream_t) stream ) );
    for(int i=0;i<reps;i++) { pq << set_lane(0) << make_op(); }
    pq << flush;
    ONIKA_CU_CHECK_ERRORS( cudaEventRecord( e1 , (cudaStream_t) stream ) );
    pq << synchronize;
    float ms = 0; ONIKA_CU_CHECK_ERRORS( cudaEventElapsedTime( &ms , e0 , e1 ) );
    std::printf("%-44s %9.1f us/op  %8.1f GB/s\n", name , ms*1e3/reps , bytes*reps/(ms*1e-3)/1e9 );
  };
  {
    void *px = nullptr, *py = nullptr;                       // device memory: no managed-memory migration in the timings
    ONIKA_B200_CHECK( onk_alloc( n*8 , 256 , ONK_MEM_DEVICE , &px ) ); ONIKA_B200_CHECK( onk_alloc( n*8 , 256 , ONK_MEM_DEVICE , &py ) );
    ONIKA_B200_CHECK( onk_memset_bytes( px , 0 , n*8 , 0 ) ); ONIKA_B200_CHECK( onk_memset_bytes( py , 0 , n*8 , 0 ) );
    AxpyFunctor f{ (double*)py , (const double*)px , 0.5 };
    bench( "parallel_for axpy (generic functor path)" , 24.0*n , [&]{ return parallel_for( n , f , parallel_execution_context("axpy") ); } );
    bench( "parallel_memset (typed or generic)" , 8.0*n , [&]{ return parallel_memset( (double*)py , n , 1.5 , parallel_execution_context("memset") ); } );
    ONIKA_B200_CHECK( onk_free( px , n*8 ) ); ONIKA_B200_CHECK( onk_free( py , n*8 ) );
  }
  {
    const size_t rows = 8192, cols = n / 8192;
    extras::Array2D a; a.resize( rows , cols );
    ONIKA_CU_CHECK_ERRORS( cudaMemPrefetchAsync( a.data() , rows*cols*8 , 0 , (cudaStream_t) stream ) );
    extras::BlockParallelValueAddFunctor<true> f = { a , 1.1 };
    bench( "block_parallel_for value-add rows" , 16.0*rows*cols , [&]{ return block_parallel_for( rows , f , parallel_execution_context("vadd") ); } );
    BlockParallelForOptions ws; ws.fixed_gpu_grid_size = true;
    bench( "block_parallel_for value-add rows, work-stealing (generic)" , 16.0*rows*cols , [&]{ return block_parallel_for( rows , f , parallel_execution_context("vadd") , ws ); } );
    for(int mult : { 8 , 16 , 32 })
    {
      ParallelExecutionContext::s_gpu_sm_mult = mult;
      char name[96]; std::snprintf( name , sizeof(name) , "  same, gpu_sm_mult = %d (grid = %d CTAs)" , mult , 148*mult );
      bench( name , 16.0*rows*cols , [&]{ return block_parallel_for( rows , f , parallel_execution_context("vadd") , ws ); } );
    }
    ParallelExecutionContext::s_gpu_sm_mult = -1;
  }
  return 0;
}

int main(int argc, char* argv[])
{
  if( argc < 3 ) { std::fprintf(stderr,"usage: %s <scenario> <output> [args]\n",argv[0]); return 2; }
  const std::string sc = argv[1];
  Out out( argv[2] );
  auto arg = [&](int i, long def) { return ( argc > i ) ? std::atol(argv[i]) : def; };
  if( onika::cuda::CudaContext::default_cuda_ctx() == nullptr ) return 9;
  int rc = 2;
  if( sc == "axpy" ) rc = scenario_axpy( out , arg(3,100003) );
  else if( sc == "value_add" ) rc = scenario_value_add( out , arg(3,257) , arg(4,129) );
  else if( sc == "lanes" ) rc = scenario_lanes( out , argc>3 ? argv[3] : "hybrid-lane" , arg(4,1024) , arg(5,1024) , arg(6,1) != 0 );
  else if( sc == "autolanes" ) rc = scenario_autolanes( out , arg(3,64) , arg(4,4096) , int(arg(5,2000)) );
  else if( sc == "autolanes_raw_war" ) rc = scenario_autolanes_raw_war( out , arg(3,64) , arg(4,4096) , int(arg(5,2000)) );
  else if( sc == "autolanes_perf" ) rc = scenario_autolanes_perf( out , arg(3,4) , arg(4,8192) , int(arg(5,20000)) );
  else if( sc == "graph" ) rc = scenario_graph( out , arg(3,128) , arg(4,512) , int(arg(5,3)) , int(arg(6,200)) );
  else if( sc == "cells_generic_perf" ) rc = scenario_cells_generic_perf( out , arg(3,1<<18) );
  else if( sc == "grid3d" ) rc = scenario_grid3d( out );
  else if( sc == "reduce" ) rc = scenario_reduce( out , arg(3,777) , arg(4,1001) );
  else if( sc == "cells" ) rc = scenario_cells( out , arg(3,4096) );
  else if( sc == "listed" ) rc = scenario_listed( out );
  else if( sc == "soatl" ) rc = scenario_soatl( out , arg(3,53) );
  else if( sc == "soatl_apply_mixed" ) rc = scenario_soatl_apply_mixed( out , arg(3,5000) );
  else if( sc == "sub_block_listed" ) rc = scenario_sub_block_listed( out , arg(3,1001) );
  else if( sc == "host_block_scratch" ) rc = scenario_host_block_scratch( out , arg(3,200000) );
  else if( sc == "sub_block_box" ) rc = scenario_sub_block_box( out , ssize_t(arg(3,13)) , ssize_t(arg(4,11)) , ssize_t(arg(5,7)) );
  else if( sc == "soatl_apply" ) rc = scenario_soatl_apply( out , arg(3,100003) , arg(4,0) != 0 );
  else if( sc == "perf" ) rc = scenario_perf( out , arg(3,1<<27) );
  else if( sc == "block_vocabulary" ) rc = scenario_block_vocabulary( out , arg(3,1000) );
  else if( sc == "helpers" ) rc = scenario_helpers( out , arg(3,333) );
  else if( sc == "reduce_helpers" ) rc = scenario_reduce_helpers( out , arg(3,513) , arg(4,1027) );
  ParallelExecutionQueue::default_queue() << flush << synchronize;
  return rc;
}
