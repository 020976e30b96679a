// One process per GPU, no MPI and no Python: the C++ multi-GPU helper (onika/parallel/multi_gpu.h).
//   multi_gpu_test <rank> <world> <rendezvous-dir> <n> <out-file> [transport: 0 device memory (default), 1 host shared memory]
// Each rank reduces its contiguous shard of a deterministic array with the fused reduce+exchange kernel (min, max, sum) and
// all-reduces three host scalars in the reference's all_reduce_multi call shape; results go to <out-file> as doubles.
#include <onika/cuda/cuda_context.h>
#include <onika/parallel/multi_gpu.h>
#include <onika/memory/allocator.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

static inline double lcg(uint64_t i, uint64_t seed, double lo, double hi)
{
  uint64_t h = ( i * 2654435761ull + seed * 40503ull + 12345ull ) & 0xFFFFFFFFull;
  h = ( ( h ^ ( h >> 15 ) ) * 2246822519ull ) & 0xFFFFFFFFull;
  h = ( h ^ ( h >> 13 ) ) & 0xFFFFFFFFull;
  return lo + ( hi - lo ) * ( double(h) / 4294967296.0 );
}

int main(int argc, char* argv[])
{
  if( argc < 6 ) { std::fprintf(stderr,"usage: %s rank world dir n out\n",argv[0]); return 2; }
  const int rank = std::atoi(argv[1]) , world = std::atoi(argv[2]);
  const std::string dir = argv[3];
  const uint64_t n = std::strtoull( argv[4] , nullptr , 10 );
  setenv( "ONIKA_B200_DEVICE" , argv[1] , 1 );
  if( onika::cuda::CudaContext::default_cuda_ctx() == nullptr ) return 9;
  using namespace onika::parallel;
  MultiGpuExchange x( rank , world , dir , "mgx" , argc > 6 ? std::atoi(argv[6]) : -1 );

  const uint64_t b = ( uint64_t(rank) * n ) / world , e = ( uint64_t(rank+1) * n ) / world;
  std::vector<double> h( e - b );
  for(uint64_t i=b;i<e;i++) h[i-b] = lcg( i , 7 , -1.0 , 1.0 );
  double *d = nullptr , *res = nullptr;
  ONIKA_B200_CHECK( onk_alloc( ( e - b + 1 ) * 8 , 256 , ONK_MEM_DEVICE , (void**) &d ) );
  ONIKA_B200_CHECK( onk_alloc( 3 * 8 , 64 , ONK_MEM_PINNED , (void**) &res ) );
  ONIKA_B200_CHECK( onk_memcpy( d , h.data() , ( e - b ) * 8 , 0 ) );
  std::vector<double> out;
  for(int rep=0;rep<2;rep++)
  {
    x.reduce_device( MultiGpuOp::MIN , d , e - b , res + 0 , 0 );
    x.reduce_device( MultiGpuOp::MAX , d , e - b , res + 1 , 0 );
    x.reduce_device( MultiGpuOp::SUM , d , e - b , res + 2 , 0 );
    ONIKA_B200_CHECK( onk_lane_sync(0) );
    out.insert( out.end() , res , res + 3 );
  }
  // the reference's call shape: all_reduce_multi( comm , op , T{} , values... )
  double a = 0.5 * ( rank + 1 ) , c = -3.0 + rank; float f = 1.25f * float( rank + 1 ); int cnt = 10 + rank;
  all_reduce_multi( x , MultiGpuOp::SUM , double{} , a , f , cnt );
  double lo = c , hi = c;
  all_reduce_multi( x , MultiGpuOp::MIN , double{} , lo );
  all_reduce_multi( x , MultiGpuOp::MAX , double{} , hi );
  out.push_back( a ); out.push_back( double(f) ); out.push_back( double(cnt) ); out.push_back( lo ); out.push_back( hi );
  out.push_back( x.healthy() ? 1.0 : 0.0 );
  x.barrier( 0 );
  ONIKA_B200_CHECK( onk_lane_sync(0) );
  FILE* fp = std::fopen( argv[5] , "wb" );
  std::fwrite( out.data() , sizeof(double) , out.size() , fp );
  std::fclose( fp );
  onk_free( d , 0 ); onk_free( res , 0 );
  return 0;
}
