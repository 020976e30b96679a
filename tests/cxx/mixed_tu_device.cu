// nvcc half of the mixed-compiler test: reads the default CUDA context that the g++-compiled translation unit created and
// runs a parallel operation whose grid is sized from it.  Exit code 0 = both compilers agree on every layout and value.
#include <onika/cuda/cuda_context.h>
#include <onika/parallel/block_parallel_for.h>
#include <onika/parallel/parallel_execution_context_allocator.h>
#include <onika/parallel/parallel_execution_queue.h>
#include <onika/memory/allocator.h>
#include <cstddef>
#include <cstdio>

extern "C" void mixed_host_tu_probe(long out[8]);

struct FillFunctor
{
  double* d; size_t cols;
  ONIKA_HOST_DEVICE_FUNC inline void operator () ( size_t i ) const { ONIKA_CU_BLOCK_SIMD_FOR(size_t, j, 0, cols) d[i*cols+j] = double(i) + 0.5; }
};
namespace onika { namespace parallel { template<> struct BlockParallelForFunctorTraits<FillFunctor> { static inline constexpr bool CudaCompatible = true; }; } }

int main()
{
  long h[8];
  mixed_host_tu_probe( h );
  auto* ctx = onika::cuda::CudaContext::default_cuda_ctx();
  const long mine[8] = { ctx ? long(ctx->m_devices[0].m_deviceProp.multiProcessorCount) : -1 , (long) sizeof(onika::cuda::CudaDevice) , (long) sizeof(onika::cuda::CudaContext)
                       , (long) offsetof(onikaDeviceProp_t, multiProcessorCount) , (long) sizeof(onikaDeviceProp_t) , (long) sizeof(onikaDim3_t)
                       , (long) sizeof(onika::parallel::ParallelExecutionContext) , (long) reinterpret_cast<std::uintptr_t>( ctx ) };
  for(int k=0;k<8;k++) if( h[k] != mine[k] ) { std::printf("mismatch #%d : host TU %ld , nvcc TU %ld\n", k, h[k], mine[k]); return 10 + k; }
  int sms = 0; ONIKA_B200_CHECK( onk_device_info( nullptr , &sms , nullptr , nullptr , 0 ) );
  if( mine[0] != sms || sms <= 0 ) { std::printf("SM count %ld vs %d\n", mine[0], sms); return 3; }
  // fixed-grid operation: the grid comes from multiProcessorCount of the shared context
  const size_t rows = 1000, cols = 257;
  onika::memory::CudaMMVector<double> v( rows*cols , 0.0 );
  onika::parallel::ParallelExecutionContextAllocator peca;
  auto & pq = onika::parallel::ParallelExecutionQueue::default_queue();
  onika::parallel::BlockParallelForOptions o; o.fixed_gpu_grid_size = true;
  auto* pec = peca.create( "mixed" , "fill" , &pq , ctx , 0 );
  const unsigned int grid_before = pec->m_grid_size.x;
  { auto op = onika::parallel::block_parallel_for( rows , FillFunctor{ v.data() , cols } , pec , o ); if( pec->m_grid_size.x < (unsigned) sms ) { std::printf("grid %u from %u\n", pec->m_grid_size.x, grid_before); return 4; } }
  pq << onika::parallel::flush << onika::parallel::synchronize;
  for(size_t i=0;i<rows;i++) for(size_t j=0;j<cols;j++) if( v[i*cols+j] != double(i) + 0.5 ) return 5;
  std::printf("mixed TU ok : %d SMs , CudaDevice %ld bytes in both compilers\n", sms, mine[1]);
  return 0;
}
