// onika/soatl/compute.h -- soatl::apply / parallel_apply / apply_simd / parallel_apply_simd : f( col0[i], col1[i], ... ).
//
// Reference: host-only OpenMP loops.  B200 front end: when the operator can run on the device (an extended
// __host__ __device__ lambda or a functor with __device__ call operator, in a .cu translation unit) and the columns
// are device addressable, the parallel variants launch a grid-stride kernel over the columns on the default lane and
// synchronise; otherwise (plain host lambdas, host-only TUs) they are sequential host loops, which is the caller's own
// host code.  The chunked parallel_apply_simd sweeps ceil(N/C)*C elements like the reference does (padding included).
#pragma once
#include <onika/soatl/field_id.h>
#include <onika/soatl/constants.h>
#include <onika/memory/simd.h>
#include <onika/cuda/cuda.h>
#include <onika/cuda/cuda_error.h>
#include <onika_b200.h>

namespace onika { namespace soatl {

  // A functor CLASS whose call operator is ONIKA_HOST_DEVICE_FUNC can declare itself runnable on the device (extended lambdas
  // are recognised by the compiler; named types are not), in the style of parallel::BlockParallelForFunctorTraits:
  //   namespace onika { namespace soatl { template<> struct ApplyFunctorTraits<MyOp> { static constexpr bool CudaCompatible = true; }; } }
  template<class OperatorT> struct ApplyFunctorTraits { static inline constexpr bool CudaCompatible = false; };

  namespace compute_details
  {
#   if defined(__CUDACC__)
    template<class OperatorT, class... T>
    __global__ void __launch_bounds__(256) apply_kernel( __grid_constant__ const OperatorT f , const size_t n , T* __restrict__ ... cols )
    {
      for(size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) f( cols[i] ... );
    }
    template<class OperatorT> static inline constexpr bool device_callable_v =
#     if defined(__CUDACC_EXTENDED_LAMBDA__)
      __nv_is_extended_host_device_lambda_closure_type(OperatorT) || __nv_is_extended_device_lambda_closure_type(OperatorT) ||
#     endif
      ApplyFunctorTraits<OperatorT>::CudaCompatible;
    template<class OperatorT, class... T>
    static inline bool try_device_apply( const OperatorT& f , size_t n , T* ... cols )
    {
      if constexpr ( device_callable_v<OperatorT> )
      {
        int ok = 1;
        ( ... , ( [&]{ int yes = 0; ONIKA_B200_CHECK( onk_is_gpu_addressable( cols , &yes ) ); ok = ok && yes; }() ) );
        if( ! ok ) return false;
        if( n == 0 ) return true;
        int sms = 148;
        ONIKA_B200_CHECK( onk_device_info( nullptr , &sms , nullptr , nullptr , 0 ) );
        const size_t need = ( n + 255 ) / 256;
        const unsigned grid = (unsigned)( need < size_t(sms) * 8 ? need : size_t(sms) * 8 );
        void* s = nullptr;
        ONIKA_B200_CHECK( onk_lane_stream( ONK_LANE_DEFAULT , &s ) );
        apply_kernel<<< grid , 256 , 0 , (cudaStream_t) s >>>( f , n , cols ... );
        ONIKA_CU_CHECK_ERRORS( cudaGetLastError() );
        ONIKA_B200_CHECK( onk_lane_sync( ONK_LANE_DEFAULT ) );
        return true;
      }
      else return false;
    }
#   else
    template<class OperatorT, class... T> static inline bool try_device_apply( const OperatorT& , size_t , T* ... ) { return false; }
#   endif
    template<class OperatorT, class... T>
    static inline void host_apply( OperatorT f , size_t n , T* __restrict__ ... cols ) { for(size_t i=0;i<n;i++) f( cols[i] ... ); }
  }

  // ---------------- sequential ----------------
  template<typename OperatorT, typename... T>
  static inline void apply( OperatorT f, size_t N, T* __restrict__ ... arraypack ) { compute_details::host_apply( f , N , arraypack ... ); }
  template<typename OperatorT, typename... T>
  static inline void apply( OperatorT f, size_t first, size_t N, T* __restrict__ ... arraypack ) { apply( f , N , arraypack+first ... ); }
  template<typename OperatorT, typename FieldArraysT, typename... ids>
  static inline void apply( OperatorT f, size_t first, size_t N, FieldArraysT& arrays, const FieldId<ids> & ... fids ) { apply( f , N , arrays[fids]+first ... ); }
  template<typename OperatorT, typename FieldArraysT, typename... ids>
  static inline void apply( OperatorT f, size_t N, FieldArraysT& arrays, const FieldId<ids> & ... fids ) { apply( f , N , arrays[fids] ... ); }
  template<typename OperatorT, typename FieldArraysT, typename... ids>
  static inline void apply( OperatorT f, FieldArraysT& arrays, const FieldId<ids> & ... fids ) { apply( f , arrays.size() , arrays[fids] ... ); }

  template<typename OperatorT, typename... T>
  static inline void apply_simd( OperatorT f, size_t N, T* __restrict__ ... arraypack ) { compute_details::host_apply( f , N , arraypack ... ); }
  template<typename OperatorT, size_t V, typename... T>
  static inline void apply_simd( OperatorT f, size_t N, cst::chunk<V>, T* __restrict__ ... arraypack ) { compute_details::host_apply( f , ( ( N + V - 1 ) / V ) * V , arraypack ... ); }
  template<typename OperatorT, typename... T>
  static inline void apply_simd( OperatorT f, size_t first, size_t N, T* __restrict__ ... arraypack ) { apply_simd( f , N , arraypack+first ... ); }
  template<typename OperatorT, typename FieldArraysT, typename... ids>
  static inline void apply_simd( OperatorT f, size_t first, size_t N, FieldArraysT& arrays, const FieldId<ids>& ... fids ) { apply_simd( f , N , arrays[fids]+first ... ); }
  template<typename OperatorT, typename FieldArraysT, typename... ids>
  static inline void apply_simd( OperatorT f, size_t N, FieldArraysT& arrays, const FieldId<ids>& ... fids ) { apply_simd( f , N , cst::chunk<FieldArraysT::ChunkSize>() , arrays[fids] ... ); }
  template<typename OperatorT, typename FieldArraysT, typename... ids>
  static inline void apply_simd( OperatorT f, FieldArraysT& arrays, const FieldId<ids>& ... fids ) { apply_simd( f , arrays.size() , cst::chunk<FieldArraysT::ChunkSize>() , arrays[fids] ... ); }

  // ---------------- parallel ----------------
  template<typename OperatorT, typename... T>
  static inline void parallel_apply( OperatorT f, size_t N, T* __restrict__ ... arraypack )
  {
    if( ! compute_details::try_device_apply( f , N , arraypack ... ) ) compute_details::host_apply( f , N , arraypack ... );
  }
  template<typename OperatorT, typename... T>
  static inline void parallel_apply( OperatorT f, size_t first, size_t N, T* __restrict__ ... arraypack ) { parallel_apply( f , N , arraypack+first ... ); }
  template<typename OperatorT, typename FieldArraysT, typename... ids>
  static inline void parallel_apply( OperatorT f, size_t first, size_t N, FieldArraysT& arrays, const FieldId<ids> & ... fids ) { parallel_apply( f , N , arrays[fids]+first ... ); }
  template<typename OperatorT, typename FieldArraysT, typename... ids>
  static inline void parallel_apply( OperatorT f, size_t N, FieldArraysT& arrays, const FieldId<ids> & ... fids ) { parallel_apply( f , N , arrays[fids] ... ); }
  template<typename OperatorT, typename FieldArraysT, typename... ids>
  static inline void parallel_apply( OperatorT f, FieldArraysT& arrays, const FieldId<ids> & ... fids ) { parallel_apply( f , arrays.size() , arrays[fids] ... ); }

  template<typename OperatorT, typename... T>
  static inline void parallel_apply_simd( OperatorT f, size_t N, T* __restrict__ ... arraypack ) { parallel_apply( f , N , arraypack ... ); }
  template<typename OperatorT, size_t V, typename... T>
  static inline void parallel_apply_simd( OperatorT f, size_t N, cst::chunk<V>, T* __restrict__ ... arraypack ) { parallel_apply( f , ( ( N + V - 1 ) / V ) * V , arraypack ... ); }
  template<typename OperatorT, typename... T>
  static inline void parallel_apply_simd( OperatorT f, size_t first, size_t N, T* __restrict__ ... arraypack ) { parallel_apply_simd( f , N , arraypack+first ... ); }
  template<typename OperatorT, typename FieldArraysT, typename... ids>
  static inline void parallel_apply_simd( OperatorT f, size_t first, size_t N, FieldArraysT& arrays, const FieldId<ids>& ... fids ) { parallel_apply_simd( f , N , arrays[fids]+first ... ); }
  template<typename OperatorT, typename FieldArraysT, typename... ids>
  static inline void parallel_apply_simd( OperatorT f, size_t N, FieldArraysT& arrays, const FieldId<ids>& ... fids ) { parallel_apply_simd( f , N , cst::chunk<FieldArraysT::ChunkSize>() , arrays[fids] ... ); }
  template<typename OperatorT, typename FieldArraysT, typename... ids>
  static inline void parallel_apply_simd( OperatorT f, FieldArraysT& arrays, const FieldId<ids>& ... fids ) { parallel_apply_simd( f , arrays.size() , cst::chunk<FieldArraysT::ChunkSize>() , arrays[fids] ... ); }

} }
