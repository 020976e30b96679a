// onika/app/api.h -- BOUNDARY SHIM of the application bootstrap, just enough for programs that drive the data-parallel hot
// path directly (the reference's tests/parallel_queue_lane.cu): a default-constructible ApplicationConfiguration holding
// the knobs that reach the hot path ("Config / flag system") and the two initialisation calls such programs
// make.  The real application layer (YAML, MPI start-up, plugin database, profiling summary ...) is out of scope.
#pragma once
#include <omp.h>
#include <memory>
#include <utility>
#include <vector>
#include <onika/cuda/cuda_context.h>
#include <onika/memory/allocator.h>
#include <onika/parallel/parallel_execution_context.h>
// what the reference's api.h brings along through the operator layer
#include <onika/parallel/parallel_execution_context_allocator.h>
#include <onika/parallel/block_parallel_for.h>
#include <onika/parallel/parallel_for.h>
#include <onika/scg/operator.h>

namespace onika { namespace app {

  struct ApplicationConfiguration
  {
    // configuration.onika.* (include/onika/app/default_app_config.h:97-106); negative = keep the built-in default
    int  parallel_task_core_mult = -1;
    int  parallel_task_core_add  = -1;
    int  gpu_sm_mult = -1;
    int  gpu_sm_add  = -1;
    int  gpu_block_size = -1;
    std::vector<long> gpu_block_dims = {};
    // top level flags
    bool nogpu = false;
    int  omp_num_threads = -1;
    int  omp_max_nesting = -1;
  };

  struct ApplicationContext
  {
    int m_cpucount = 0;
    int m_ngpus = 0;
  };

  // OpenMP side of the bootstrap: thread count as configured (host-only functors still run on host threads)
  inline int intialize_openmp( ApplicationConfiguration & configuration , int /*mpi_rank*/ = 0 , bool allow_openmp_conf = true )
  {
    if( allow_openmp_conf )
    {
      if( configuration.omp_num_threads > 0 ) omp_set_num_threads( configuration.omp_num_threads );
      if( configuration.omp_max_nesting > 0 ) omp_set_max_active_levels( configuration.omp_max_nesting );
    }
    return omp_get_max_threads();
  }

  // GPU side: tuning knobs into ParallelExecutionContext, device context created, host allocations GPU-addressable
  inline std::shared_ptr<ApplicationContext> initialize_gpu( const ApplicationConfiguration & configuration )
  {
    using PEC = onika::parallel::ParallelExecutionContext;
    if( configuration.parallel_task_core_mult >= 0 ) PEC::s_parallel_task_core_mult = configuration.parallel_task_core_mult;
    if( configuration.parallel_task_core_add  >= 0 ) PEC::s_parallel_task_core_add  = configuration.parallel_task_core_add;
    if( configuration.gpu_sm_mult >= 0 ) PEC::s_gpu_sm_mult = configuration.gpu_sm_mult;
    if( configuration.gpu_sm_add  >= 0 ) PEC::s_gpu_sm_add  = configuration.gpu_sm_add;
    if( configuration.gpu_block_size > 0 ) PEC::s_gpu_block_size = configuration.gpu_block_size;
    if( configuration.gpu_block_dims.size() == 3 )
      PEC::s_gpu_block_dims = onikaDim3_t{ (unsigned int) configuration.gpu_block_dims[0] , (unsigned int) configuration.gpu_block_dims[1] , (unsigned int) configuration.gpu_block_dims[2] };
    auto ctx = std::make_shared<ApplicationContext>();
    ctx->m_cpucount = omp_get_max_threads();
    onika::cuda::CudaContext::set_global_gpu_enable( ! configuration.nogpu );
    auto * cu = onika::cuda::CudaContext::default_cuda_ctx();
    ctx->m_ngpus = ( cu != nullptr ) ? int( cu->device_count() ) : 0;
    onika::memory::GenericHostAllocator::set_cuda_enabled( ctx->m_ngpus > 0 );
    return ctx;
  }

} }
