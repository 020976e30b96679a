// onika/soatl/compute.h -- soatl::apply / parallel_apply / apply_simd / parallel_apply_simd : f( col0[i], col1[i], ... ).
//
// Reference: host-only OpenMP loops.  B200 front end: when the operator can run on the device (an extended
// __host__ __device__ lambda or a functor with __device__ call operator, in a .cu translation unit) and the columns
// are device addressable, the parallel variants launch a vectorised kernel over the columns (256-bit accesses, stores only
// for the columns the operator takes by non-const reference) on the default lane and synchronise, or on the caller's lane
// without synchronising inside a DeviceApplyScope; otherwise (plain host lambdas, host-only TUs) they are sequential host loops, which is the caller's own
// host code.  The chunked parallel_apply_simd sweeps ceil(N/C)*C elements like the reference does (padding included).
#pragma once
#include <onika/soatl/field_id.h>
#include <onika/soatl/constants.h>
#include <onika/memory/simd.h>
#include <onika/cuda/cuda.h>
#include <onika/cuda/cuda_error.h>
#include <onika_b200.h>
#include <algorithm>
#include <cstring>
#include <tuple>
#include <type_traits>
#include <utility>

namespace onika { namespace soatl {

  // A functor CLASS whose call operator is ONIKA_HOST_DEVICE_FUNC can declare itself runnable on the device (extended lambdas
  // are recognised by the compiler; named types are not), in the style of parallel::BlockParallelForFunctorTraits:
  //   namespace onika { namespace soatl { template<> struct ApplyFunctorTraits<MyOp> { static constexpr bool CudaCompatible = true; }; } }
  template<class OperatorT> struct ApplyFunctorTraits { static inline constexpr bool CudaCompatible = false; };

  // Where the device path of parallel_apply* runs: execution lane and whether the call returns only when the sweep is done
  // (the reference's host loops are synchronous, hence the default).  A DeviceApplyScope in the caller's scope puts the
  // sweeps of this thread on the caller's lane, optionally without the host synchronisation, so that they order with the
  // parallel operations queued on that lane:   { soatl::DeviceApplyScope s( lane , false ); soatl::parallel_apply_simd(...); }
  namespace compute_details { struct ApplyTarget { int lane = ONK_LANE_DEFAULT; bool synchronize = true; };
                              inline ApplyTarget& apply_target() { static thread_local ApplyTarget t; return t; } }
  struct DeviceApplyScope
  {
    compute_details::ApplyTarget m_saved;
    inline DeviceApplyScope( int lane , bool synchronize = false ) : m_saved( compute_details::apply_target() ) { compute_details::apply_target() = { lane , synchronize }; }
    inline ~DeviceApplyScope() { compute_details::apply_target() = m_saved; }
    DeviceApplyScope( const DeviceApplyScope& ) = delete;
    DeviceApplyScope& operator = ( const DeviceApplyScope& ) = delete;
  };

  namespace compute_details
  {
#   if defined(__CUDACC__)
    // ---- which columns does the operator write ? ----
    // An operator with ONE non-template call operator (lambdas, plain functors) tells through its parameter types: a column
    // received by value or by const reference is only read, so the kernel never stores it; anything else (generic lambdas,
    // overload sets) is treated as read-write for every column that is not const-qualified.
    template<class Sig> struct call_signature;
    template<class C, class R, class... A> struct call_signature< R (C::*)(A...) const > { template<size_t K> using arg = std::tuple_element_t< K , std::tuple<A...> >; };
    template<class C, class R, class... A> struct call_signature< R (C::*)(A...) >       { template<size_t K> using arg = std::tuple_element_t< K , std::tuple<A...> >; };
    template<class C, class R, class... A> struct call_signature< R (C::*)(A...) const noexcept > { template<size_t K> using arg = std::tuple_element_t< K , std::tuple<A...> >; };
    template<class C, class R, class... A> struct call_signature< R (C::*)(A...) noexcept >       { template<size_t K> using arg = std::tuple_element_t< K , std::tuple<A...> >; };
    template<class F, size_t K, class = void> struct writes_argument : std::true_type {};
    template<class F, size_t K> struct writes_argument< F , K , std::void_t< decltype( & F::operator() ) > >
    {
      using A = typename call_signature< decltype( & F::operator() ) >::template arg<K>;
      static inline constexpr bool value = std::is_lvalue_reference_v<A> && ! std::is_const_v< std::remove_reference_t<A> >;
    };

    // ---- one vector of E elements of T: 32 / 16 / 8 / 4 / 2 / 1 bytes moved by ONE global access ----
    template<class T, unsigned E> struct alignas( E * sizeof(T) ) ApplyVec { T v[E]; };

    // loads are plain (non-volatile) asm statements: when the operator overwrites a column without reading it, nothing uses
    // the loaded registers and the compiler drops the load -- a write-only column costs one store, as in a hand-written kernel
    template<bool ReadOnly, class V> __device__ __forceinline__ V load_vec( const V* p )
    {
      V r;
      if constexpr ( sizeof(V) == 32 )
      {
        unsigned long long w[4];
        if constexpr ( ReadOnly ) asm("ld.global.nc.L1::no_allocate.L2::evict_first.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(w[0]), "=l"(w[1]), "=l"(w[2]), "=l"(w[3]) : "l"(p));
        else                      asm("ld.global.L1::no_allocate.L2::evict_first.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(w[0]), "=l"(w[1]), "=l"(w[2]), "=l"(w[3]) : "l"(p));
        memcpy( &r , w , 32 );
      }
      else if constexpr ( sizeof(V) == 16 )
      {
        unsigned long long w[2];
        if constexpr ( ReadOnly ) asm("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(w[0]), "=l"(w[1]) : "l"(p));
        else                      asm("ld.global.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(w[0]), "=l"(w[1]) : "l"(p));
        memcpy( &r , w , 16 );
      }
      else if constexpr ( sizeof(V) == 8 ) { unsigned long long w; if constexpr ( ReadOnly ) asm("ld.global.nc.u64 %0, [%1];" : "=l"(w) : "l"(p)); else asm("ld.global.u64 %0, [%1];" : "=l"(w) : "l"(p)); memcpy( &r , &w , 8 ); }
      else if constexpr ( sizeof(V) == 4 ) { unsigned int w;       if constexpr ( ReadOnly ) asm("ld.global.nc.u32 %0, [%1];" : "=r"(w) : "l"(p)); else asm("ld.global.u32 %0, [%1];" : "=r"(w) : "l"(p)); memcpy( &r , &w , 4 ); }
      else if constexpr ( sizeof(V) == 2 ) { unsigned short w;     if constexpr ( ReadOnly ) asm("ld.global.nc.u16 %0, [%1];" : "=h"(w) : "l"(p)); else asm("ld.global.u16 %0, [%1];" : "=h"(w) : "l"(p)); memcpy( &r , &w , 2 ); }
      else { r = *p; }
      return r;
    }
    template<class V> __device__ __forceinline__ void store_vec( V* p , const V& r )
    {
      if constexpr ( sizeof(V) == 32 )
      {
        unsigned long long w[4]; memcpy( w , &r , 32 );
        asm volatile("st.global.L1::no_allocate.v4.u64 [%0], {%1,%2,%3,%4};" :: "l"(p), "l"(w[0]), "l"(w[1]), "l"(w[2]), "l"(w[3]) : "memory");
      }
      else if constexpr ( sizeof(V) == 16 )
      {
        unsigned long long w[2]; memcpy( w , &r , 16 );
        asm volatile("st.global.L1::no_allocate.v2.u64 [%0], {%1,%2};" :: "l"(p), "l"(w[0]), "l"(w[1]) : "memory");
      }
      else { *p = r; }
    }

    template<class... T> static inline constexpr size_t max_sizeof_v = std::max( { sizeof(T) ... } );
    template<class T> static inline constexpr bool vectorisable_v = std::is_trivially_copyable_v<T> && ( sizeof(T) & ( sizeof(T) - 1 ) ) == 0 && sizeof(T) <= 16;

    // Vector body: thread t of a 256-thread CTA takes elements [ (tile*256+t)*E , +E ) of EVERY column with one access per
    // column (256-bit for the widest type), all loads issued before the first call, then E calls of the operator on the
    // register copies, then one store per written column.  Grid = resident CTAs x SMs walking tiles grid-stride.  The scalar
    // tail (< 256*E elements) and unaligned column sets use one element per thread.
    template<unsigned E, class OperatorT, size_t... K, class... T>
    __device__ __forceinline__ void apply_tile( const OperatorT& f , const size_t i , std::index_sequence<K...> , T* ... cols )
    {
      std::tuple< ApplyVec< std::remove_const_t<T> , E > ... > v = { load_vec< std::is_const_v<T> || ! writes_argument<OperatorT,K>::value >( reinterpret_cast< const ApplyVec< std::remove_const_t<T> , E > * >( cols + i ) ) ... };
#     pragma unroll
      for(unsigned j=0;j<E;j++) f( std::get<K>(v).v[j] ... );
      ( ... , ( [&]{
          if constexpr ( ! std::is_const_v<T> && writes_argument<OperatorT,K>::value )
            store_vec( reinterpret_cast< ApplyVec< std::remove_const_t<T> , E > * >( cols + i ) , std::get<K>(v) );
        }() ) );
    }

    template<unsigned E, class OperatorT, class... T>
    __global__ void __launch_bounds__(256) apply_vec_kernel( __grid_constant__ const OperatorT f , const size_t n , const size_t ntiles , T* ... cols )
    {
      for(size_t t = blockIdx.x; t < ntiles; t += gridDim.x)
        apply_tile<E>( f , ( t * 256 + threadIdx.x ) * E , std::index_sequence_for<T...>{} , cols ... );
      for(size_t i = ntiles * 256 * E + size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) f( cols[i] ... );
    }

    template<class OperatorT, class... T>
    __global__ void __launch_bounds__(256) apply_kernel( __grid_constant__ const OperatorT f , const size_t n , T* ... cols )
    {
      for(size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) f( cols[i] ... );
    }
    template<class OperatorT> static inline constexpr bool device_callable_v =
#     if defined(__CUDACC_EXTENDED_LAMBDA__)
      __nv_is_extended_host_device_lambda_closure_type(OperatorT) || __nv_is_extended_device_lambda_closure_type(OperatorT) ||
#     endif
      ApplyFunctorTraits<OperatorT>::CudaCompatible;

    template<class KernelT> static inline unsigned int resident_grid( KernelT kernel , size_t nctas )
    {
      static thread_local KernelT s_kernel = nullptr;
      static thread_local size_t s_cap = 0;
      if( s_kernel != kernel )
      {
        int per_sm = 1 , sms = 148;
        ONIKA_CU_CHECK_ERRORS( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &per_sm , kernel , 256 , 0 ) );
        ONIKA_B200_CHECK( onk_device_info( nullptr , &sms , nullptr , nullptr , 0 ) );
        s_cap = size_t( per_sm < 1 ? 1 : per_sm ) * size_t(sms); s_kernel = kernel;
      }
      return (unsigned int)( nctas < s_cap ? ( nctas ? nctas : 1 ) : s_cap );
    }

    template<class OperatorT, class... T>
    static inline bool try_device_apply( const OperatorT& f , size_t n , T* ... cols )
    {
      if constexpr ( device_callable_v<OperatorT> )
      {
        int ok = 1;
        ( ... , ( [&]{ int yes = 0; ONIKA_B200_CHECK( onk_is_gpu_addressable( cols , &yes ) ); ok = ok && yes; }() ) );
        if( ! ok ) return false;
        if( n == 0 ) return true;
        const ApplyTarget target = apply_target();
        void* s = nullptr;
        ONIKA_B200_CHECK( onk_lane_stream( target.lane , &s ) );
        bool launched = false;
        if constexpr ( ( ... && vectorisable_v<T> ) )
        {
          constexpr unsigned E = unsigned( 32 / max_sizeof_v<T...> );
          const bool aligned = ( ... && ( ( reinterpret_cast<uintptr_t>( cols ) % ( E * sizeof(T) ) ) == 0 ) );
          const size_t ntiles = n / ( 256 * size_t(E) );
          if( aligned && ntiles > 0 )
          {
            auto kernel = apply_vec_kernel< E , OperatorT , T... >;
            kernel<<< resident_grid( kernel , ntiles ) , 256 , 0 , (cudaStream_t) s >>>( f , n , ntiles , cols ... );
            launched = true;
          }
        }
        if( ! launched )
        {
          auto kernel = apply_kernel< OperatorT , T... >;
          kernel<<< resident_grid( kernel , ( n + 255 ) / 256 ) , 256 , 0 , (cudaStream_t) s >>>( f , n , cols ... );
        }
        ONIKA_CU_CHECK_ERRORS( cudaGetLastError() );
        if( target.synchronize ) ONIKA_B200_CHECK( onk_lane_sync( target.lane ) );
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
