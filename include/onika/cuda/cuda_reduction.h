// onika/cuda/cuda_reduction.h -- block_reduce_add / block_reduce_min / block_reduce_max for user functors.
//
// Same call shape as the reference helpers (value, IntConst<BlockSize>) -> reduced value, valid in every thread.
// B200 design: butterfly shuffles inside each warp (no shared memory, no barrier), ONE shared-memory exchange of the
// BlockSize/32 warp results, then a second butterfly -- 2 barriers instead of 2*log2(BlockSize) in the reference's
// shared-memory tree.  Comparison semantics follow the reference: `a > b ? b : a` for min, `a < b ? b : a` for max.
// In the host pass (a block is one thread) they are the identity, as in the reference.
#pragma once
#include <onika/cuda/cuda.h>
#include <onika/integral_constant.h>

namespace onika { namespace cuda {

  namespace reduction_details
  {
    struct OpAdd { template<class T> ONIKA_HOST_DEVICE_FUNC static inline T apply(const T& a, const T& b) { return a + b; } };
    struct OpMin { template<class T> ONIKA_HOST_DEVICE_FUNC static inline T apply(const T& a, const T& b) { return ( a > b ) ? b : a; } };
    struct OpMax { template<class T> ONIKA_HOST_DEVICE_FUNC static inline T apply(const T& a, const T& b) { return ( a < b ) ? b : a; } };

    template<class Op, class T, int BlockSize>
    ONIKA_HOST_DEVICE_FUNC static inline T block_reduce(const T& x)
    {
#     if defined(__CUDA_ARCH__)
      static_assert( sizeof(T) <= 8 && std::is_trivially_copyable_v<T> , "shuffle based reduction needs a <= 8 byte trivially copyable type" );
      if constexpr ( BlockSize <= 32 )
      {
        // a block (or a sub-block of a CTA that packs several small items) inside one warp: butterfly over its own lanes
        // only -- no shared memory, no barrier, and the other lanes of the warp may be anywhere else in the code
        static_assert( BlockSize >= 1 && ( BlockSize & ( BlockSize - 1 ) ) == 0 , "a block smaller than a warp must be a power of two" );
        const unsigned int lanes = lane_group_mask( BlockSize );
        T v = x;
#       pragma unroll
        for(int o=BlockSize/2;o>0;o>>=1) v = Op::apply( v , __shfl_xor_sync( lanes , v , o ) );
        return v;
      }
      else
      {
      static_assert( BlockSize % 32 == 0 && BlockSize <= 1024 , "block size must be a multiple of the warp size" );
      constexpr int NW = BlockSize / 32;
      __shared__ alignas(8) unsigned char s_raw[ ( NW + 1 ) * sizeof(T) ];
      T* s_part = reinterpret_cast<T*>( s_raw );
      T v = x;
#     pragma unroll
      for(int o=16;o>0;o>>=1) v = Op::apply( v , __shfl_xor_sync( 0xffffffffu , v , o ) );
      const int tid = ONIKA_CU_THREAD_IDX;
      if( ( tid & 31 ) == 0 ) s_part[ tid >> 5 ] = v;
      __syncthreads();
      if( tid < 32 )
      {
        T r = s_part[ ( tid < NW ) ? tid : 0 ];
        if( tid >= NW ) r = s_part[0];                      // idempotent filler for min/max; masked out for add below
        if constexpr ( std::is_same_v<Op,OpAdd> ) { if( tid >= NW ) r = T{}; }
#       pragma unroll
        for(int o=16;o>0;o>>=1) r = Op::apply( r , __shfl_xor_sync( 0xffffffffu , r , o ) );
        if( tid == 0 ) s_part[NW] = r;
      }
      __syncthreads();
      const T result = s_part[NW];
      __syncthreads();                                       // s_raw may be reused by the next call
      return result;
      }
#     else
      return x;
#     endif
    }
  }

  template<class T, int BlockSize>
  ONIKA_HOST_DEVICE_FUNC static inline T block_reduce_add(const T& x , onika::IntConst<BlockSize> ) { return reduction_details::block_reduce<reduction_details::OpAdd,T,BlockSize>(x); }
  template<class T, int BlockSize>
  ONIKA_HOST_DEVICE_FUNC static inline T block_reduce_min(const T& x , onika::IntConst<BlockSize> ) { return reduction_details::block_reduce<reduction_details::OpMin,T,BlockSize>(x); }
  template<class T, int BlockSize>
  ONIKA_HOST_DEVICE_FUNC static inline T block_reduce_max(const T& x , onika::IntConst<BlockSize> ) { return reduction_details::block_reduce<reduction_details::OpMax,T,BlockSize>(x); }

} }
