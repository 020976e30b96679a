// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Thin extern "C" driver around the *unmodified reference* (Collab4exaNBody/onika,
// read from /root/reference at build time, never copied into this repo).  It is compiled
// together with nine reference .cpp files (see oracle/Makefile) into
// oracle/_ref/libonika_ref.so and exposes the reference's own OpenMP execution path
// (parallel_for / block_parallel_for / ParallelExecutionQueue lanes / SOATL FieldArrays /
// soatl::parallel_apply_simd) behind plain-pointer entry points, so that
//   * tests/ can pin the C restatement in oracle/oracle.c against the real reference, and
//   * bench.py can time the reference CPU path on the GPU box's host cores.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library.  The product (onika_b200/) never does.
//
// Every entry point returns the wall time (seconds) spent inside the reference call itself
// (std::chrono around the call, as the reference does: block_parallel_for_adapter.h:265,370),
// excluding the copies this driver performs to move data in/out of reference containers.

#include <onika/parallel/parallel_for.h>
#include <onika/parallel/block_parallel_for.h>
#include <onika/parallel/memset.h>
#include <onika/parallel/parallel_execution_context_allocator.h>
#include <onika/parallel/parallel_execution_queue.h>
#include <onika/parallel/parallel_execution_operators.h>
#include <onika/extras/array2d.h>
#include <onika/extras/block_parallel_value_add_functor.h>
#include <onika/soatl/field_id.h>
#include <onika/soatl/field_arrays.h>
#include <onika/soatl/compute.h>
#include <onika/soatl/copy.h>

#include <chrono>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cmath>
#include <omp.h>

// ---- SOATL field declarations: same pattern as reference tests/declare_fields.h:23-30 ----
#define REF_DECLARE_FIELD(__type,__name,__desc) \
struct __##__name {}; \
namespace onika { namespace soatl { template<> struct FieldId<__##__name> { \
    using value_type = __type; \
    using Id = __##__name; \
    static const char* name() { return __desc ; } \
  }; } } \
static ::onika::soatl::FieldId<__##__name> __name;

REF_DECLARE_FIELD(double        ,f_rx    ,"rx")
REF_DECLARE_FIELD(double        ,f_ry    ,"ry")
REF_DECLARE_FIELD(double        ,f_rz    ,"rz")
REF_DECLARE_FIELD(unsigned char ,f_atype ,"atype")
REF_DECLARE_FIELD(double        ,f_e     ,"e")
REF_DECLARE_FIELD(int32_t       ,f_mid   ,"mid")
REF_DECLARE_FIELD(float         ,f_dist  ,"dist")
REF_DECLARE_FIELD(int16_t       ,f_tmp1  ,"tmp1")
REF_DECLARE_FIELD(int8_t        ,f_tmp2  ,"tmp2")
REF_DECLARE_FIELD(double        ,f_0     ,"f0")
REF_DECLARE_FIELD(double        ,f_1     ,"f1")
REF_DECLARE_FIELD(double        ,f_2     ,"f2")
REF_DECLARE_FIELD(double        ,f_3     ,"f3")
REF_DECLARE_FIELD(double        ,f_4     ,"f4")
REF_DECLARE_FIELD(double        ,f_5     ,"f5")
REF_DECLARE_FIELD(double        ,f_6     ,"f6")
REF_DECLARE_FIELD(double        ,f_7     ,"f7")

// cpu_init / cpu_compute of the reference test tests/gpu_reduce_min_fp64.cu:30-55 (N fixed to 1024*1024
// there).  That TU is compiled as-is (with -Dmain=...) by oracle/Makefile; both are external symbols.
void   cpu_init( double * __restrict__ data );
double cpu_compute( double * __restrict__ data );

namespace
{
  using clk = std::chrono::high_resolution_clock;
  inline double secs(clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b-a).count(); }

  using namespace onika;
  using namespace onika::parallel;

  ParallelExecutionContextAllocator & peca()
  {
    static ParallelExecutionContextAllocator * a = new ParallelExecutionContextAllocator(); // leaked on purpose (static dtor order)
    return *a;
  }

  // same recipe as SURVEY.md 8c step 3: host target, fork-join OpenMP (omp_num_tasks = 0)
  ParallelExecutionContext* new_pec(const char* tag, const char* sub = "", int omp_num_tasks = 0)
  {
    return peca().create( tag, sub, & ParallelExecutionQueue::default_queue(), nullptr, omp_num_tasks );
  }

  // y[i] += a*x[i] : BASELINE config C1 (no such functor exists in onika; the *dispatch* is the reference's)
  struct AxpyFunctor
  {
    double * __restrict__ y; const double * __restrict__ x; double a;
    inline void operator () ( size_t i ) const { y[i] += a * x[i]; }
  };

  // per-element value add used for the 3D coverage pattern
  // (mirrors Grid3DBenchmarkFunctor, plugins/onikaParallelTutorial/parallel_for_3d_benchmark.cu:50-65)
  struct Grid3DFunctor
  {
    double * __restrict__ m; long n;
    inline void operator () ( onikaInt3_t c ) const
    {
      const unsigned long N = n, i = c.x, j = c.y, k = c.z;
      ONIKA_CU_ATOMIC_ADD( m[(k*N+j)*N+i] , 1.0 );
    }
  };

  using CellArrays = soatl::FieldArrays<64,16,__f_rx,__f_ry,__f_rz,__f_e>;

  // per-cell sum of field e + global sum: BASELINE config C4, written the way an exaNBody operator would
  // (block per cell, threads stride the cell's particles, atomics into return data)
  struct CellSumFunctor
  {
    const CellArrays * cells; double * cell_sum; double * total;
    inline void operator () ( size_t c ) const
    {
      const double * __restrict__ e = cells[c][f_e];
      const size_t n = cells[c].size();
      double s = 0.0;
      for(size_t p=ONIKA_CU_THREAD_IDX; p<n; p+=ONIKA_CU_BLOCK_SIZE) s += e[p];
      cell_sum[c] = s;
      ONIKA_CU_ATOMIC_ADD( *total , s );
    }
  };
}

// ---- templates used by the extern "C" entry points below ----
template<class FA, class... Fids>
static int layout_seq(FA fa, const uint64_t* sizes, int nsteps, uint64_t* cap, uint64_t* bytes, uint64_t* offs, Fids... fids)
{
  constexpr int nf = sizeof...(Fids);
  for(int s=0;s<nsteps;s++)
  {
    fa.resize( sizes[s] );
    cap[s] = fa.capacity();
    bytes[s] = fa.storage_size();
    const uint8_t* base = (const uint8_t*) fa.storage_ptr();
    const uint8_t* p[nf] = { (const uint8_t*) fa[fids] ... };
    for(int k=0;k<nf;k++) offs[s*8+k] = ( base==nullptr ) ? 0 : uint64_t( p[k] - base );
  }
  fa.resize(0);
  return nf;
}


template<size_t A>
static double soatl_rmw8(uint64_t n, double* const* cols, double s, const double* c)
{
  using namespace onika::soatl;
  auto arrays = make_hybrid_field_arrays( cst::align<A>(), cst::chunk<16>(), f_0,f_1,f_2,f_3,f_4,f_5,f_6,f_7 );
  arrays.resize(n);
  double* p[8] = { arrays[f_0],arrays[f_1],arrays[f_2],arrays[f_3],arrays[f_4],arrays[f_5],arrays[f_6],arrays[f_7] };
  for(int k=0;k<8;k++) { for(size_t i=n;i<arrays.capacity();i++) p[k][i]=0.0; std::memcpy( p[k], cols[k], n*sizeof(double) ); }
  const double c0=c[0],c1=c[1],c2=c[2],c3=c[3],c4=c[4],c5=c[5],c6=c[6],c7=c[7];
  auto t0 = clk::now();
  parallel_apply_simd( [=](double& a0,double& a1,double& a2,double& a3,double& a4,double& a5,double& a6,double& a7)
    {
      a0 = a0*s + c0; a1 = a1*s + c1; a2 = a2*s + c2; a3 = a3*s + c3;
      a4 = a4*s + c4; a5 = a5*s + c5; a6 = a6*s + c6; a7 = a7*s + c7;
    }
    , arrays, f_0,f_1,f_2,f_3,f_4,f_5,f_6,f_7 );
  const double t = secs(t0, clk::now());
  for(int k=0;k<8;k++) std::memcpy( cols[k], p[k], n*sizeof(double) );
  arrays.resize(0);
  return t;
}

extern "C"
{

int ref_num_threads() { return omp_get_max_threads(); }

// ---------------- C1: parallel_for axpy (include/onika/parallel/parallel_for.h:91-112) ----------------
double ref_parallel_for_axpy(double* y, const double* x, double a, uint64_t n)
{
  AxpyFunctor f{ y, x, a };
  auto t0 = clk::now();
  parallel_for( n, f, new_pec("axpy") );
  return secs(t0, clk::now());
}

// ---------------- parallel_memset (include/onika/parallel/memset.h:30-51) ----------------
double ref_parallel_memset_f64(double* d, uint64_t n, double v)
{
  auto t0 = clk::now();
  parallel_memset( d, n, v, new_pec("memset") );
  return secs(t0, clk::now());
}

// -------- block_parallel_for + BlockParallelValueAddFunctor (extras/block_parallel_value_add_functor.h:15-58) --------
// ndims = 1 : block_parallel_for(rows, functor)  ;  ndims = 2 : ParallelExecutionSpace<2>{ {0,0},{rows,cols} }
double ref_block_value_add(double* arr, uint64_t rows, uint64_t cols, double v, int ndims)
{
  extras::Array2D a; a.resize(rows, cols);
  std::memcpy( a.data(), arr, rows*cols*sizeof(double) );
  extras::BlockParallelValueAddFunctor<false> f = { a, v };
  auto t0 = clk::now();
  if( ndims == 2 )
  {
    const ParallelExecutionSpace<2> r = { {0,0} , { ssize_t(rows), ssize_t(cols) } };
    block_parallel_for( r, f, new_pec("vadd","2d") );
  }
  else
  {
    block_parallel_for( rows, f, new_pec("vadd","1d") );
  }
  const double t = secs(t0, clk::now());
  std::memcpy( arr, a.data(), rows*cols*sizeof(double) );
  return t;
}

// ---------------- C2: reduce-min (tests/gpu_reduce_min_fp64.cu:46-55), N = 1024*1024 per call ----------------
// nchunks consecutive 1 Mi-element chunks, each reduced by the reference's own cpu_compute; min of mins is exact.
double ref_reduce_min_chunks(const double* d, uint64_t nchunks, double* out)
{
  double m = std::numeric_limits<double>::max();
  auto t0 = clk::now();
  for(uint64_t c=0;c<nchunks;c++)
  {
    const double r = cpu_compute( const_cast<double*>( d + c*(1024ull*1024ull) ) );
    if( r < m ) m = r;
  }
  const double t = secs(t0, clk::now());
  *out = m;
  return t;
}
void ref_reduce_min_init_1m(double* d) { cpu_init(d); }

// ---------------- 3D parallel_for coverage (parallel_for.h:114-155; tutorial range parallel_for_3d_benchmark.cu:118-125) -----
double ref_parallel_for_3d(double* m, long n, const long* start, const long* end, int use_block_parallel_for)
{
  Grid3DFunctor f{ m, n };
  ParallelExecutionSpace<3> r = { { start[0],start[1],start[2] } , { end[0],end[1],end[2] } };
  auto t0 = clk::now();
  if( use_block_parallel_for ) block_parallel_for( r, f, new_pec("grid3d","block") );
  else parallel_for( r, f, new_pec("grid3d","pfor") );
  return secs(t0, clk::now());
}

// ---------------- C5: lane sequence (tests/parallel_queue_lane.cu:35-166, value-add kernels) ----------------
// mode 0 "hybrid-lane", 1 "hybrid-delayed-lane", 2 "hybrid-sequence"; omp_num_tasks=0 (pfor) unless use_tasks.
double ref_lane_sequence(double* a1_io, double* a2_io, uint64_t rows, uint64_t cols,
                         double v11, double v12, double v21, double v22, int mode, int use_tasks)
{
  extras::Array2D array1, array2;
  array1.resize(rows, cols); array2.resize(rows, cols);
  std::memcpy( array1.data(), a1_io, rows*cols*sizeof(double) );
  std::memcpy( array2.data(), a2_io, rows*cols*sizeof(double) );
  double t = 0.0;
  auto body = [&](int ntasks)
  {
    auto & pq = ParallelExecutionQueue::default_queue();
    extras::BlockParallelValueAddFunctor<false> a1k1 = { array1, v11 };
    extras::BlockParallelValueAddFunctor<false> a1k2 = { array1, v12 };
    extras::BlockParallelValueAddFunctor<false> a2k1 = { array2, v21 };
    extras::BlockParallelValueAddFunctor<false> a2k2 = { array2, v22 };
    auto t0 = clk::now();
    auto op11 = block_parallel_for( array1.rows(), a1k1, new_pec("Array1","Kernel1",ntasks) );
    auto op12 = block_parallel_for( array1.rows(), a1k2, new_pec("Array1","Kernel2",ntasks) );
    auto op21 = block_parallel_for( array2.rows(), a2k1, new_pec("Array2","Kernel1",ntasks) );
    auto op22 = block_parallel_for( array2.rows(), a2k2, new_pec("Array2","Kernel2",ntasks) );
    if( mode == 1 )
    {
      ParallelExecutionQueueBase qa, qb;
      qa << set_lane(0) << std::move(op11) << set_lane(1) << std::move(op21);
      qb << set_lane(0) << std::move(op12) << set_lane(1) << std::move(op22);
      pq << std::move(qa) << std::move(qb);
    }
    else if( mode == 0 )
    {
      pq << set_lane(0) << std::move(op11) << set_lane(1) << std::move(op21)
         << set_lane(0) << std::move(op12) << set_lane(1) << std::move(op22);
    }
    else
    {
      pq << std::move(op11) << std::move(op12) << std::move(op21) << std::move(op22);
    }
    pq << flush;
    pq << synchronize;
    t = secs(t0, clk::now());
  };
  if( use_tasks )
  {
#   pragma omp parallel
    {
#     pragma omp master
      {
#       pragma omp taskgroup
        { body( omp_get_max_threads() ); }
      }
    }
  }
  else body(0);
  std::memcpy( a1_io, array1.data(), rows*cols*sizeof(double) );
  std::memcpy( a2_io, array2.data(), rows*cols*sizeof(double) );
  return t;
}

// ---------------- SOATL layout (field_arrays.h:200-211,477-501; packed_field_arrays.h:36-178; alloc_strategy.h:44-67) -------
// cfg: 0 <64,16, rx,ry,rz,e>            (tests/benchmark.cpp)
//      1 <64,8,  e,atype,rx,mid,ry,rz>  (tests/soatlserializetest.cpp in/out)
//      2 <1,1,   e,atype,rx,mid,ry,rz>  (tests/soatlserializetest.cpp packed)
//      3 <64,16, rx,ry,rz,atype,e,mid>  (SURVEY 8a row a23 probe)
//      4 <64,16, f0..f7>                (config C3)
//      5 <256,16,f0..f7>                (config C3, CUDA alignment)
//      6 <32,4,  dist,tmp1,tmp2,atype,rx> (mixed small types)
// Applies the resize sequence sizes[0..nsteps) to one container and reports, after each step,
// capacity, storage bytes and the byte offset of every field (up to 8).  Returns number of fields.
int ref_soatl_layout(int cfg, const uint64_t* sizes, int nsteps, uint64_t* cap, uint64_t* bytes, uint64_t* offs)
{
  using namespace onika::soatl;
  switch(cfg)
  {
    case 0: return layout_seq( make_hybrid_field_arrays(cst::align<64>(),cst::chunk<16>(),f_rx,f_ry,f_rz,f_e), sizes,nsteps,cap,bytes,offs, f_rx,f_ry,f_rz,f_e );
    case 1: return layout_seq( make_hybrid_field_arrays(cst::align<64>(),cst::chunk<8>(),f_e,f_atype,f_rx,f_mid,f_ry,f_rz), sizes,nsteps,cap,bytes,offs, f_e,f_atype,f_rx,f_mid,f_ry,f_rz );
    case 2: return layout_seq( make_hybrid_field_arrays(cst::align<1>(),cst::chunk<1>(),f_e,f_atype,f_rx,f_mid,f_ry,f_rz), sizes,nsteps,cap,bytes,offs, f_e,f_atype,f_rx,f_mid,f_ry,f_rz );
    case 3: return layout_seq( make_hybrid_field_arrays(cst::align<64>(),cst::chunk<16>(),f_rx,f_ry,f_rz,f_atype,f_e,f_mid), sizes,nsteps,cap,bytes,offs, f_rx,f_ry,f_rz,f_atype,f_e,f_mid );
    case 4: return layout_seq( make_hybrid_field_arrays(cst::align<64>(),cst::chunk<16>(),f_0,f_1,f_2,f_3,f_4,f_5,f_6,f_7), sizes,nsteps,cap,bytes,offs, f_0,f_1,f_2,f_3,f_4,f_5,f_6,f_7 );
    case 5: return layout_seq( make_hybrid_field_arrays(cst::align<256>(),cst::chunk<16>(),f_0,f_1,f_2,f_3,f_4,f_5,f_6,f_7), sizes,nsteps,cap,bytes,offs, f_0,f_1,f_2,f_3,f_4,f_5,f_6,f_7 );
    case 6: return layout_seq( make_hybrid_field_arrays(cst::align<32>(),cst::chunk<4>(),f_dist,f_tmp1,f_tmp2,f_atype,f_rx), sizes,nsteps,cap,bytes,offs, f_dist,f_tmp1,f_tmp2,f_atype,f_rx );
    default: return -1;
  }
}

// ---------------- C3a: the stock 3R+1W SOATL benchmark kernel (tests/benchmark.cpp:88-110) ----------------
// arrays <64,16,rx,ry,rz,e>; (ax,ay,az) = element 0 as in benchmark.cpp:92-94; parallel_apply_simd as in :106-108.
double ref_soatl_dist(uint64_t n, const double* rx, const double* ry, const double* rz, double* e_out)
{
  using namespace onika::soatl;
  auto arrays = make_hybrid_field_arrays( cst::align<64>(), cst::chunk<16>(), f_rx, f_ry, f_rz, f_e );
  arrays.resize(n);
  // padding elements (touched by parallel_apply_simd up to ceil(n/16)*16, compute.h:221-231) get benign values
  for(size_t i=n;i<arrays.capacity();i++) { arrays[f_rx][i]=2.0; arrays[f_ry][i]=3.0; arrays[f_rz][i]=4.0; arrays[f_e][i]=0.0; }
  std::memcpy( arrays[f_rx], rx, n*sizeof(double) );
  std::memcpy( arrays[f_ry], ry, n*sizeof(double) );
  std::memcpy( arrays[f_rz], rz, n*sizeof(double) );
  std::memset( arrays[f_e], 0, n*sizeof(double) );
  const double ax = arrays[f_rx][0], ay = arrays[f_ry][0], az = arrays[f_rz][0];
  auto t0 = clk::now();
  parallel_apply_simd( [ax,ay,az](double& d, double x, double y, double z)
    {
      x = x - ax;
      y = y - ay;
      z = z - az;
      d = std::sqrt( x*x + y*y + z*z );
      x /= d;
      y /= d;
      z /= d;
      d += x/(y*z);
    }
    , arrays, f_e, f_rx, f_ry, f_rz );
  const double t = secs(t0, clk::now());
  std::memcpy( e_out, arrays[f_e], n*sizeof(double) );
  arrays.resize(0);
  return t;
}

// ---------------- C3b: 8-field read-modify-write  f_k[i] = f_k[i]*s + c_k   ----------------
// cols: 8 host pointers of n doubles each (in/out).  align 64 or 256.
double ref_soatl_rmw8(uint64_t n, double* const* cols, double s, const double* c, int align)
{
  return ( align == 256 ) ? soatl_rmw8<256>(n,cols,s,c) : soatl_rmw8<64>(n,cols,s,c);
}

// ---------------- SOATL serialisation (tests/soatlserializetest.cpp:61-107; soatl/copy.h:29-97) ----------------
// fills <64,8,e,atype,rx,mid,ry,rz>, soatl::copy into the <1,1,...> packed layout, returns packed bytes.
uint64_t ref_soatl_serialize(uint64_t n, const double* e, const uint8_t* atype, const double* rx, const int32_t* mid,
                             const double* ry, const double* rz, uint8_t* packed_out, uint64_t packed_capacity_bytes)
{
  using namespace onika::soatl;
  auto in_arrays  = make_hybrid_field_arrays( cst::align<64>(), cst::chunk<8>(), f_e,f_atype,f_rx,f_mid,f_ry,f_rz );
  auto ser_arrays = make_hybrid_field_arrays( cst::align<1>(),  cst::chunk<1>(), f_e,f_atype,f_rx,f_mid,f_ry,f_rz );
  in_arrays.resize(n);
  std::memcpy( in_arrays[f_e], e, n*sizeof(double) );
  std::memcpy( in_arrays[f_atype], atype, n );
  std::memcpy( in_arrays[f_rx], rx, n*sizeof(double) );
  std::memcpy( in_arrays[f_mid], mid, n*sizeof(int32_t) );
  std::memcpy( in_arrays[f_ry], ry, n*sizeof(double) );
  std::memcpy( in_arrays[f_rz], rz, n*sizeof(double) );
  ser_arrays.resize(n);
  copy( ser_arrays, in_arrays );
  const uint64_t sz = ser_arrays.storage_size();
  if( sz <= packed_capacity_bytes ) std::memcpy( packed_out, ser_arrays.storage_ptr(), sz );
  ser_arrays.resize(0);
  in_arrays.resize(0);
  return sz;
}

// ---------------- C4: cell grid, one FieldArrays<64,16,rx,ry,rz,e> per cell, per-cell sum of e + global sum -------------
// counts[c] particles in cell c; e_concat holds the e values cell after cell.  schedule: 0 dynamic 1 guided 2 static.
double ref_cell_sum(uint64_t ncells, const uint32_t* counts, const double* e_concat, double* cell_sum, double* total,
                    uint64_t* allocated_bytes, int schedule)
{
  std::vector<CellArrays> cells( ncells );
  uint64_t off = 0, bytes = 0;
  for(uint64_t c=0;c<ncells;c++)
  {
    cells[c].resize( counts[c] );
    if( counts[c] > 0 ) std::memcpy( cells[c][f_e], e_concat+off, counts[c]*sizeof(double) );
    off += counts[c];
    bytes += cells[c].storage_size();
  }
  if( allocated_bytes != nullptr ) *allocated_bytes = bytes;
  *total = 0.0;
  CellSumFunctor f{ cells.data(), cell_sum, total };
  BlockParallelForOptions opts; opts.omp_scheduling = OMPScheduling(schedule);
  auto t0 = clk::now();
  block_parallel_for( ncells, f, new_pec("cells"), opts );
  const double t = secs(t0, clk::now());
  for(auto & c : cells) c.resize(0);
  return t;
}

} // extern "C"
