// onika/parallel/block_parallel_for_adapter.h -- device kernels and the type-erased adapter behind block_parallel_for.
//
// The reference launches one CUDA block per item (grid = N) through 10-line trampolines.  B200 design:
//   * bpf_items_kernel      persistent item loop: grid = min(N, occupancy x 148 SMs), CTA b runs items b, b+grid, ...
//                           (no 2M-block launches for cell-sized items; one barrier between items keeps functors that
//                           use block-shared scratch correct)
//   * bpf_sub_items_kernel  "sub-block items": for functors whose traits declare SubBlockItems and whose block is <= 32 threads
//                           (opts.max_block_size), a CTA of 256 threads holds 256/bs independent sub-blocks, each running the
//                           item loop on its own -- a cell of ~16 particles occupies 16 lanes instead of a 128-thread block,
//                           and an SM keeps 2048/bs items in flight instead of 16-32.  The ONIKA_CU_* block vocabulary means
//                           the sub-block inside such a CTA (include/onika/cuda/cuda.h), so functor sources do not change;
//                           opt-in because function-scope ONIKA_CU_BLOCK_SHARED variables would be shared by the sub-blocks.
//                           bpf_sub_worksteal_kernel is the same with a fixed grid: every sub-block draws its own tickets.
//   * bpf_box_kernel        same for 2D/3D boxes: items are box coordinates, first coordinate fastest; bpf_sub_box_kernel packs
//                           small items of such boxes into CTAs of sub-blocks (SubBlockItems + opts.max_block_size <= 32)
//   * bpf_worksteal_kernel  fixed grid + atomic ticket (opts.fixed_gpu_grid_size), 1D only, as in the reference
// Kernels live in the plugin's translation unit (device code for user functors never enters the runtime library);
// they are launched through the C ABI: cudaGetFuncBySymbol once per instantiation, then onk_launch(cufunc, ..., lane).
#pragma once
#include <chrono>
#include <onika/parallel/block_parallel_for_functor.h>
#include <onika/parallel/parallel_execution_stream.h>
#include <onika/log.h>

namespace onika { namespace parallel {

  // ========================== typed fast paths ==========================
  // Customisation point: a functor type whose whole operation is one of the typed sm_100a kernels of libonika_b200
  // (include/onika_b200.h) specialises this; the adapter then issues that kernel on the operation's lane instead of the
  // generic functor kernel.  launch() returns false when this particular call cannot use it (e.g. a sub-box that is not
  // one contiguous stream), in which case the generic path runs.  Results are identical by construction (same
  // arithmetic, parity-tested against the same oracle).
  template<class FuncT, unsigned int ND, unsigned int ElemND>
  struct B200TypedFastPath
  {
    static inline constexpr bool available = false;
    template<class SpaceT> static inline bool launch( const FuncT& , const SpaceT& , int /*lane*/ ) { return false; }
  };

  // ========================== device kernels ==========================
#if defined(__CUDACC__)
  namespace b200
  {
    template<bool Indexed, class ElementListT, class FuncT>
    __global__ void __launch_bounds__(ONIKA_CU_MAX_THREADS_PER_BLOCK)
    bpf_items_kernel( const ElementListT idx , __grid_constant__ const FuncT func , const ssize_t start , const unsigned long long n )
    {
      for(unsigned long long i = blockIdx.x; i < n; i += gridDim.x)
      {
        if constexpr ( ! Indexed ) func( start + ssize_t(i) );
        else func( idx[ start + ssize_t(i) ] );
        if( i + gridDim.x < n ) __syncthreads();
      }
    }

    // CTA = blockDim.y sub-blocks of blockDim.x <= 32 threads; sub-block w of W runs items w, w+W, ... (neighbouring lanes of
    // a warp work on neighbouring items: their index / offset loads share cache lines)
    template<bool Indexed, class ElementListT, class FuncT>
    __global__ void __launch_bounds__(256)
    bpf_sub_items_kernel( const ElementListT idx , __grid_constant__ const FuncT func , const ssize_t start , const unsigned long long n )
    {
      const unsigned long long W = (unsigned long long) gridDim.x * blockDim.y;
      for(unsigned long long i = (unsigned long long) blockIdx.x * blockDim.y + threadIdx.y; i < n; i += W)
      {
        if constexpr ( ! Indexed ) func( start + ssize_t(i) );
        else func( idx[ start + ssize_t(i) ] );
        ONIKA_CU_BLOCK_SYNC();
      }
    }

    // the same CTAs of sub-blocks with a fixed grid and work stealing (opts.fixed_gpu_grid_size): every sub-block draws its own
    // tickets -- its first lane one item ahead, the value reaches the others through a shuffle over the sub-block's lanes
    template<bool Indexed, class ElementListT, class FuncT>
    __global__ void __launch_bounds__(256)
    bpf_sub_worksteal_kernel( const unsigned long long n , GPUKernelExecutionScratch* scratch , const ElementListT idx , __grid_constant__ const FuncT func , const ssize_t start )
    {
#     if defined(__CUDA_ARCH__)
      const unsigned int lanes = ::onika::cuda::lane_group_mask( blockDim.x );
      const int leader = __ffs( int(lanes) ) - 1;
      unsigned long long * counter = & scratch->counters[GPUKernelExecutionScratch::WORKSTEALING_COUNTER];
      unsigned long long next = 0;
      if( threadIdx.x == 0 ) next = atomicAdd( counter , 1ull );
      next = __shfl_sync( lanes , next , leader );
      for(;;)
      {
        const unsigned long long i = next;
        if( i >= n ) break;
        if( threadIdx.x == 0 ) next = atomicAdd( counter , 1ull );     // the ticket of the following item travels while this one runs
        if constexpr ( ! Indexed ) func( start + ssize_t(i) );
        else func( idx[ start + ssize_t(i) ] );
        next = __shfl_sync( lanes , next , leader );
      }
#     endif
    }

    // 2D/3D blocks follow ParallelExecutionContext::gpu_block_dims() (8x8x4 = 256 threads by default): no 128-thread bound here
    template<class FuncT>
    __global__ void
    bpf_box_kernel( __grid_constant__ const FuncT func , const onikaInt3_t start , const unsigned int ex , const unsigned int ey , const unsigned long long n )
    {
      for(unsigned long long i = blockIdx.x; i < n; i += gridDim.x)
      {
        const unsigned long long jk = i / ex;
        const onikaInt3_t c = { start.x + ssize_t( i - jk * ex ) , start.y + ssize_t( jk % ey ) , start.z + ssize_t( jk / ey ) };
        func( c );
        if( i + gridDim.x < n ) __syncthreads();
      }
    }

    // 2D/3D boxes of small items: CTAs of sub-blocks as in bpf_sub_items_kernel, every sub-block a 1-D group of blockDim.x threads
    template<class FuncT>
    __global__ void __launch_bounds__(256)
    bpf_sub_box_kernel( __grid_constant__ const FuncT func , const onikaInt3_t start , const unsigned int ex , const unsigned int ey , const unsigned long long n )
    {
      const unsigned long long W = (unsigned long long) gridDim.x * blockDim.y;
      for(unsigned long long i = (unsigned long long) blockIdx.x * blockDim.y + threadIdx.y; i < n; i += W)
      {
        const unsigned long long jk = i / ex;
        const onikaInt3_t c = { start.x + ssize_t( i - jk * ex ) , start.y + ssize_t( jk % ey ) , start.z + ssize_t( jk / ey ) };
        func( c );
        ONIKA_CU_BLOCK_SYNC();
      }
    }

    template<bool Indexed, class ElementListT, class FuncT>
    __global__ void __launch_bounds__(ONIKA_CU_MAX_THREADS_PER_BLOCK)
    bpf_worksteal_kernel( const unsigned long long n , GPUKernelExecutionScratch* scratch , const ElementListT idx , __grid_constant__ const FuncT func , const ssize_t start )
    {
      // double-buffered ticket: thread 0 draws the NEXT item while the block works on the current one, so the atomic's
      // round trip overlaps the functor and one barrier per item suffices (tickets drawn past n are harmless)
      __shared__ unsigned long long item[2];
      if( ONIKA_CU_THREAD_IDX == 0 ) item[0] = atomicAdd( & scratch->counters[GPUKernelExecutionScratch::WORKSTEALING_COUNTER] , 1ull );
      __syncthreads();
      for(unsigned int k=0;;k^=1u)
      {
        const unsigned long long i = item[k];
        if( i >= n ) break;
        if( ONIKA_CU_THREAD_IDX == 0 ) item[k^1u] = atomicAdd( & scratch->counters[GPUKernelExecutionScratch::WORKSTEALING_COUNTER] , 1ull );
        if constexpr ( ! Indexed ) func( start + ssize_t(i) );
        else func( idx[ start + ssize_t(i) ] );
        __syncthreads();
      }
    }

    // grid for a persistent kernel: as many CTAs as can be resident, capped by the number of items
    template<class KernelT>
    inline unsigned int persistent_grid( KernelT kernel , unsigned int block_threads , unsigned long long nitems , int sm_count )
    {
      // the occupancy query costs microseconds: remembered per (kernel, block size) and thread
      static thread_local KernelT s_kernel = nullptr;
      static thread_local unsigned int s_threads = 0;
      static thread_local int s_per_sm = 0;
      if( s_kernel != kernel || s_threads != block_threads )
      {
        ONIKA_CU_CHECK_ERRORS( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &s_per_sm , kernel , int(block_threads) , 0 ) );
        if( s_per_sm < 1 ) s_per_sm = 1;
        s_kernel = kernel; s_threads = block_threads;
      }
      const int per_sm = s_per_sm;
      const unsigned long long cap = (unsigned long long) per_sm * (unsigned long long) sm_count;
      return (unsigned int)( nitems < cap ? nitems : cap );
    }

    template<class KernelT>
    inline void launch( KernelT kernel , unsigned int grid , onikaDim3_t block , int lane , void** args , int flags = 0 )
    {
      static thread_local KernelT s_kernel = nullptr;
      static thread_local cudaFunction_t s_func = nullptr;
      if( s_kernel != kernel ) { ONIKA_CU_CHECK_ERRORS( cudaGetFuncBySymbol( & s_func , (const void*) kernel ) ); s_kernel = kernel; }
      const unsigned g[3] = { grid , 1 , 1 } , b[3] = { block.x , block.y , block.z };
      ONIKA_B200_CHECK( onk_launch_ex( (void*) s_func , g , b , 0 , lane , args , flags ) );
    }
  }
#endif

  // =============== adapter: one instance is placement-constructed inside the PEC ===============
  template<class FuncT, bool GPUSupport , unsigned int ND=1, unsigned int ElemND=0, class ElementListT = std::span< const element_coord_t<ElemND> > >
  class BlockParallelForHostAdapter : public BlockParallelForHostFunctor
  {
    static inline constexpr unsigned int FuncParamDim = ( ElemND==0 ) ? ND : ElemND ;
    static_assert( FuncParamDim>=1 && FuncParamDim<=3 );
    using FuncParamType = std::conditional_t< FuncParamDim==1 , ssize_t , onikaInt3_t >;
    using ParExecSpaceT = ParallelExecutionSpace<ND,ElemND,ElementListT>;

    template<class... A> static inline constexpr bool has = lambda_is_compatible_with_v<FuncT,void,A...>;
    static inline constexpr bool functor_has_prolog            = has<block_parallel_for_prolog_t>;
    static inline constexpr bool functor_has_prolog_stream     = has<block_parallel_for_prolog_t,ParallelExecutionStream*>;
    static inline constexpr bool functor_has_cpu_prolog        = has<block_parallel_for_cpu_prolog_t>;
    static inline constexpr bool functor_has_gpu_prolog        = has<block_parallel_for_gpu_prolog_t>;
    static inline constexpr bool functor_has_gpu_prolog_stream = has<block_parallel_for_gpu_prolog_t,ParallelExecutionStream*>;
    static inline constexpr bool functor_has_epilog            = has<block_parallel_for_epilog_t>;
    static inline constexpr bool functor_has_epilog_stream     = has<block_parallel_for_epilog_t,ParallelExecutionStream*>;
    static inline constexpr bool functor_has_cpu_epilog        = has<block_parallel_for_cpu_epilog_t>;
    static inline constexpr bool functor_has_gpu_epilog        = has<block_parallel_for_gpu_epilog_t>;
    static inline constexpr bool functor_has_gpu_epilog_stream = has<block_parallel_for_gpu_epilog_t,ParallelExecutionStream*>;
    static inline constexpr bool functor_is_single_task        = has<block_parallel_for_single_task_t>;

    static_assert( has<FuncParamType> || functor_is_single_task , "User defined functor is not compatible with execution space");
    static_assert( !GPUSupport || !functor_is_single_task );

    const FuncT m_func;
    const ParExecSpaceT m_parallel_space;

  public:
    inline BlockParallelForHostAdapter( const FuncT& f , const ParExecSpaceT& ps ) : m_func(f) , m_parallel_space(ps) {}

    inline unsigned int execution_space_ndims() const override final { return m_parallel_space.space_nd(); }
    inline size_t execution_space_nitems() const override final { return m_parallel_space.number_of_items(); }
    inline bool is_single_task() const override final { return functor_is_single_task; }
    inline size_t execution_space_range(ssize_t * buf, size_t bufsz) const override final
    {
      return coord_range_to_array( buf , bufsz , m_parallel_space.m_start , m_parallel_space.m_end );
    }

    // ---------------- device target ----------------
    inline void stream_gpu_initialize(ParallelExecutionContext* , ParallelExecutionStream* pes) const override final
    {
      if constexpr ( GPUSupport )
      {
        if constexpr ( functor_has_gpu_prolog ) m_func( block_parallel_for_gpu_prolog_t{} );
        else if constexpr ( functor_has_gpu_prolog_stream ) m_func( block_parallel_for_gpu_prolog_t{} , pes );
        else if constexpr ( functor_has_prolog ) m_func( block_parallel_for_prolog_t{} );
        else if constexpr ( functor_has_prolog_stream ) m_func( block_parallel_for_prolog_t{} , pes );
      }
      else fatal_error() << "called stream_gpu_initialize with no GPU support" << std::endl;
    }

    inline void stream_gpu_kernel(ParallelExecutionContext* pec, ParallelExecutionStream*) const override final
    {
#     if defined(__CUDACC__)
      if constexpr ( GPUSupport )
      {
        const int sms = pec->m_cuda_ctx->m_devices[0].m_deviceProp.multiProcessorCount;
        const unsigned long long nitems = m_parallel_space.number_of_items();
        if( nitems == 0 ) return;
        const unsigned int threads = pec->m_block_size.x * pec->m_block_size.y * pec->m_block_size.z;
#       ifndef ONIKA_B200_DISABLE_TYPED_FAST_PATHS
        if constexpr ( B200TypedFastPath<FuncT,ND,ElemND>::available )
#       else
        if constexpr ( false )
#       endif
        {
          if( pec->m_grid_size.x == 0 && B200TypedFastPath<FuncT,ND,ElemND>::launch( m_func , m_parallel_space , pec->m_lane ) ) return;
        }
        if constexpr ( ND == 1 )
        {
          ssize_t start = m_parallel_space.m_start[0];
          auto elements = m_parallel_space.m_elements;
          unsigned long long n = nitems;
          if( pec->m_grid_size.x > 0 )
          {
            GPUKernelExecutionScratch* scratch = pec->m_cuda_scratch.get();
            void* args[] = { &n , &scratch , &elements , (void*) &m_func , &start };
            bool packed = false;
            if constexpr ( functor_sub_block_items_v<FuncT> )
            {
              if( threads <= 32 && ( threads & ( threads - 1 ) ) == 0 )
              {
                auto kernel = b200::bpf_sub_worksteal_kernel<(ElemND>=1),ElementListT,FuncT>;
                b200::launch( kernel , pec->m_grid_size.x , onikaDim3_t{ threads , 256u / threads , 1 } , pec->m_lane , args , ONK_LAUNCH_SUB_BLOCKS );
                packed = true;
              }
            }
            if( ! packed )
            {
              auto kernel = b200::bpf_worksteal_kernel<(ElemND>=1),ElementListT,FuncT>;
              b200::launch( kernel , pec->m_grid_size.x , pec->m_block_size , pec->m_lane , args );
            }
          }
          else
          {
            bool packed = false;
            if constexpr ( functor_sub_block_items_v<FuncT> )
            {
              if( threads <= 32 && ( threads & ( threads - 1 ) ) == 0 )
              {
                auto kernel = b200::bpf_sub_items_kernel<(ElemND>=1),ElementListT,FuncT>;
                const onikaDim3_t cta = { threads , 256u / threads , 1 };
                const unsigned long long nctas = ( n + cta.y - 1 ) / cta.y;
                void* args[] = { &elements , (void*) &m_func , &start , &n };
                b200::launch( kernel , b200::persistent_grid( kernel , 256 , nctas , sms ) , cta , pec->m_lane , args , ONK_LAUNCH_SUB_BLOCKS );
                packed = true;
              }
            }
            if( ! packed )
            {
              auto kernel = b200::bpf_items_kernel<(ElemND>=1),ElementListT,FuncT>;
              void* args[] = { &elements , (void*) &m_func , &start , &n };
              b200::launch( kernel , b200::persistent_grid( kernel , threads , n , sms ) , pec->m_block_size , pec->m_lane , args );
            }
          }
        }
        else
        {
          static_assert( ND == 1 || ElemND == 0 , "element indices must be 1D" );
          if( pec->m_grid_size.x > 0 ) fatal_error() << "Work stealing GPU execution only supported for 1D kernels" << std::endl;
          onikaInt3_t start = { m_parallel_space.m_start[0] , 0 , 0 };
          unsigned int ex = (unsigned int)( m_parallel_space.m_end[0] - m_parallel_space.m_start[0] ) , ey = 1;
          if constexpr ( ND >= 2 ) { start.y = m_parallel_space.m_start[1]; ey = (unsigned int)( m_parallel_space.m_end[1] - m_parallel_space.m_start[1] ); }
          if constexpr ( ND >= 3 ) { start.z = m_parallel_space.m_start[2]; }
          unsigned long long n = nitems;
          void* args[] = { (void*) &m_func , &start , &ex , &ey , &n };
          bool packed = false;
          if constexpr ( functor_sub_block_items_v<FuncT> )
          {
            // a caller that asks for blocks of <= 32 threads (opts.max_block_size) gets 1-D sub-blocks of that size instead of the
            // configured 2D/3D tile; pec->m_block_threads is only that small when the caller said so
            const unsigned int bs = pec->m_block_threads;
            if( bs >= 1 && bs <= 32 && ( bs & ( bs - 1 ) ) == 0 )
            {
              auto kernel = b200::bpf_sub_box_kernel<FuncT>;
              const onikaDim3_t cta = { bs , 256u / bs , 1 };
              const unsigned long long nctas = ( n + cta.y - 1 ) / cta.y;
              b200::launch( kernel , b200::persistent_grid( kernel , 256 , nctas , sms ) , cta , pec->m_lane , args , ONK_LAUNCH_SUB_BLOCKS );
              packed = true;
            }
          }
          if( ! packed )
          {
            auto kernel = b200::bpf_box_kernel<FuncT>;
            b200::launch( kernel , b200::persistent_grid( kernel , threads , n , sms ) , pec->m_block_size , pec->m_lane , args );
          }
        }
      }
      else
#     endif
      { (void)pec; fatal_error() << "called stream_gpu_kernel with no GPU support" << std::endl; }
    }

    inline void stream_gpu_finalize(ParallelExecutionContext* , ParallelExecutionStream* pes) const override final
    {
      if constexpr ( GPUSupport )
      {
        if constexpr ( functor_has_gpu_epilog_stream ) m_func( block_parallel_for_gpu_epilog_t{} , pes );
        else if constexpr ( functor_has_epilog_stream ) m_func( block_parallel_for_epilog_t{} , pes );
      }
      else fatal_error() << "called stream_gpu_finalize with no GPU support" << std::endl;
    }

    inline bool needs_gpu_epilog(const ParallelExecutionContext* pec) const override final
    {
      return functor_has_gpu_epilog || functor_has_epilog || ! pec->m_execution_end_callback.is_null();
    }

    // runs on a runtime thread once the lane reached this point: must not call the CUDA API
    inline void execute_gpu_epilog(const ParallelExecutionContext* pec) const override final
    {
      if constexpr ( functor_has_gpu_epilog ) m_func( block_parallel_for_gpu_epilog_t{} );
      else if constexpr ( functor_has_epilog ) m_func( block_parallel_for_epilog_t{} );
      pec->m_execution_end_callback.execute();
    }

    // ---------------- host target (user-declared host-only functors; runs as a host node of the lane) ----------------
    inline void execute_prolog( ParallelExecutionContext* , ParallelExecutionStream* pes ) const override final
    {
      if constexpr (functor_has_cpu_prolog) m_func( block_parallel_for_cpu_prolog_t{} );
      else if constexpr (functor_has_prolog) m_func( block_parallel_for_prolog_t{} );
      else if constexpr (functor_has_prolog_stream) m_func( block_parallel_for_prolog_t{} , pes );
    }

    inline void execute_epilog( ParallelExecutionContext* pec, ParallelExecutionStream* pes ) const override final
    {
      if constexpr (functor_has_cpu_epilog) m_func( block_parallel_for_cpu_epilog_t{} );
      else if constexpr (functor_has_epilog) m_func( block_parallel_for_epilog_t{} );
      else if constexpr (functor_has_epilog_stream) m_func( block_parallel_for_epilog_t{} , pes );
      pec->m_execution_end_callback.execute();
    }

    inline void execute_omp_parallel_region( ParallelExecutionContext* pec, ParallelExecutionStream* pes ) const override final
    {
      const auto T0 = std::chrono::high_resolution_clock::now();
      execute_prolog( pec , pes );
      const auto & ps = m_parallel_space;
      // the caller's OpenMP schedule (opts.omp_scheduling, default guided) through the run-time schedule of this thread: one loop
      // per dimensionality instead of one per (dimensionality, schedule) as in the reference (block_parallel_for_adapter.h:272-360)
#     if defined(_OPENMP)
      omp_set_schedule( pec->m_omp_sched == OMP_SCHED_STATIC ? omp_sched_static : ( pec->m_omp_sched == OMP_SCHED_DYNAMIC ? omp_sched_dynamic : omp_sched_guided ) , 0 );
#     endif
      if constexpr ( functor_is_single_task ) { m_func( block_parallel_for_single_task_t{} ); }
      else if constexpr ( ND == 1 )
      {
        const auto * idx = ps.m_elements.data();
        (void) idx;
#       pragma omp parallel for schedule(runtime)
        for(ssize_t i=ps.m_start[0]; i<ps.m_end[0]; i++)
        {
          if constexpr ( ElemND==0 ) m_func( i );
          else m_func( idx[i] );
        }
      }
      else if constexpr ( ND == 2 )
      {
#       pragma omp parallel for collapse(2) schedule(runtime)
        for(ssize_t j=ps.m_start[1]; j<ps.m_end[1]; j++)
        for(ssize_t i=ps.m_start[0]; i<ps.m_end[0]; i++) m_func( onikaInt3_t{i,j,0} );
      }
      else
      {
#       pragma omp parallel for collapse(3) schedule(runtime)
        for(ssize_t k=ps.m_start[2]; k<ps.m_end[2]; k++)
        for(ssize_t j=ps.m_start[1]; j<ps.m_end[1]; j++)
        for(ssize_t i=ps.m_start[0]; i<ps.m_end[0]; i++) m_func( onikaInt3_t{i,j,k} );
      }
      execute_epilog( pec , pes );
      pec->m_total_cpu_execution_time = std::chrono::duration<double,std::milli>( std::chrono::high_resolution_clock::now() - T0 ).count();
      if( pes != nullptr ) pes->m_omp_execution_count.fetch_sub(1);
    }

    inline void execute_omp_tasks( ParallelExecutionContext* pec, ParallelExecutionStream* pes, unsigned int ) const override final
    {
      execute_omp_parallel_region( pec , pes );
    }

    inline ~BlockParallelForHostAdapter() override {}
  };

} }
