// onika/parallel/block_parallel_for_functor.h -- the functor protocol of block_parallel_for:
// traits (GPU capability), optional prolog/epilog/single-task overloads, and the type-erased interface the queue calls.
#pragma once
#include <concepts>
#include <type_traits>
#include <onika/cuda/cuda.h>
#include <onika/parallel/parallel_execution_context.h>
#include <onika/lambda_tools.h>

namespace onika { namespace parallel {

  // specialise with CudaCompatible = true for functors whose call operator may run on the device
  template<class FuncT> struct BlockParallelForFunctorTraits { static inline constexpr bool CudaCompatible = false; };

  // optional member of a traits specialisation: `static inline constexpr bool SubBlockItems = true;` lets the runtime pack
  // several small items into one CTA (block_parallel_for_adapter.h, bpf_sub_items_kernel) when the operation's block size is
  // <= 32 threads.  Declare it for functors that keep no function-scope ONIKA_CU_BLOCK_SHARED state and synchronise only
  // through the ONIKA_CU_* vocabulary.
  template<class FuncT> static inline constexpr bool functor_sub_block_items_v = []{
      if constexpr ( requires { BlockParallelForFunctorTraits<FuncT>::SubBlockItems; } ) return bool( BlockParallelForFunctorTraits<FuncT>::SubBlockItems );
      else return false; }();

  template<class FuncT> struct FunctorNeedsGPUSupport : public
# if defined(__CUDACC__)
    std::integral_constant<bool,BlockParallelForFunctorTraits<FuncT>::CudaCompatible>
# else
    std::integral_constant<bool,false>
# endif
  {};
  template<class FuncT> static inline constexpr bool functor_gpu_support_v = FunctorNeedsGPUSupport<FuncT>::value;

  // tag types selecting optional overloads of the functor's call operator:
  //   f(block_parallel_for_prolog_t [,stream*])      before the operation, any target
  //   f(block_parallel_for_gpu_prolog_t [,stream*])  before, device target only     f(block_parallel_for_cpu_prolog_t) host target only
  //   f(block_parallel_for_epilog_t [,stream*]) / _gpu_ / _cpu_ variants             after completion
  struct block_parallel_for_prolog_t {};
  struct block_parallel_for_gpu_prolog_t : public block_parallel_for_prolog_t {};
  struct block_parallel_for_cpu_prolog_t : public block_parallel_for_prolog_t {};
  struct block_parallel_for_epilog_t {};
  struct block_parallel_for_gpu_epilog_t : public block_parallel_for_epilog_t {};
  struct block_parallel_for_cpu_epilog_t : public block_parallel_for_epilog_t {};

  // f(block_parallel_for_single_task_t): one host call for the whole execution space
  struct block_parallel_for_single_task_t {};
  template<class T> concept block_parallel_for_single_task = std::is_same_v<T,block_parallel_for_single_task_t>;

  template<long long FunctorSize, long long MaxSize, class FuncT>
  struct AssertFunctorSizeFitIn { std::enable_if_t< (FunctorSize <= MaxSize) , int > x = 0; };

  // what ParallelExecutionQueue::schedule() / sync_and_remove() see of a functor
  class BlockParallelForHostFunctor
  {
  public:
    // host execution (user-declared host-only functors)
    virtual inline void execute_prolog( ParallelExecutionContext*, ParallelExecutionStream* ) const {}
    virtual inline void execute_omp_parallel_region( ParallelExecutionContext*, ParallelExecutionStream* ) const {}
    virtual inline void execute_omp_tasks( ParallelExecutionContext*, ParallelExecutionStream*, unsigned int ) const {}
    virtual inline void execute_epilog( ParallelExecutionContext*, ParallelExecutionStream* ) const {}
    // device execution
    virtual inline void stream_gpu_initialize(ParallelExecutionContext*,ParallelExecutionStream*) const {}
    virtual inline void stream_gpu_kernel(ParallelExecutionContext*,ParallelExecutionStream*) const {}
    virtual inline void stream_gpu_finalize(ParallelExecutionContext*,ParallelExecutionStream*) const {}
    virtual inline void execute_gpu_epilog(const ParallelExecutionContext*) const {}
    virtual inline bool needs_gpu_epilog(const ParallelExecutionContext*) const { return false; }
    // execution space
    virtual inline unsigned int execution_space_ndims() const { return 0; }
    virtual inline size_t execution_space_nitems() const { return 0; }
    virtual inline bool is_single_task() const { return false; }
    virtual inline size_t execution_space_range(ssize_t *, size_t) const { return 0; }
    virtual inline ~BlockParallelForHostFunctor() {}
  };

  template<class FuncT> requires std::invocable<FuncT&>
  struct BlockParallelSingleTaskFunctor
  {
    FuncT m_func;
    inline void operator () ( parallel::block_parallel_for_single_task_t ) const { m_func(); }
  };
  template<class FuncT> requires std::invocable<FuncT&>
  inline auto make_single_task_block_parallel_functor( const FuncT& f ) { return BlockParallelSingleTaskFunctor<FuncT>{ f }; }

  inline const BlockParallelForHostFunctor & get_block_parallel_functor( const ParallelExecutionContext* pec )
  {
    return * reinterpret_cast<const BlockParallelForHostFunctor*>( pec->m_host_scratch.functor_data );
  }

} }
