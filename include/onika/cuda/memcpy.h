// onika/cuda/memcpy.h -- cu_block_memcpy / cu_block_memmove: all threads of a block copy one byte range together.
// B200 version: 16-byte words when source and destination are 16-byte congruent, 8-byte words when 8-byte congruent,
// bytes otherwise; coalesced (consecutive threads take consecutive words).
#pragma once
#include <cstdint>
#include <cstring>
#include <onika/cuda/cuda.h>

namespace onika { namespace cuda {

  ONIKA_HOST_DEVICE_FUNC static inline void cu_block_memcpy(void* dst, const void* src, size_t n)
  {
#   if defined(__CUDA_ARCH__)
    uint8_t* d = static_cast<uint8_t*>(dst);
    const uint8_t* s = static_cast<const uint8_t*>(src);
    const unsigned tid = ONIKA_CU_THREAD_IDX, nt = ONIKA_CU_BLOCK_SIZE;
    const uintptr_t da = reinterpret_cast<uintptr_t>(d), sa = reinterpret_cast<uintptr_t>(s);
    size_t head = 0;
    unsigned w = 1;
    if( ( ( da ^ sa ) & 15 ) == 0 ) w = 16; else if( ( ( da ^ sa ) & 7 ) == 0 ) w = 8;
    if( w > 1 ) { head = ( w - ( da & (w-1) ) ) & (w-1); if( head > n ) head = n; }
    for(size_t i=tid;i<head;i+=nt) d[i] = s[i];
    const size_t nw = ( w > 1 ) ? ( n - head ) / w : 0;
    if( w == 16 ) { const uint4* sv = reinterpret_cast<const uint4*>(s+head); uint4* dv = reinterpret_cast<uint4*>(d+head); for(size_t i=tid;i<nw;i+=nt) dv[i] = sv[i]; }
    else if( w == 8 ) { const uint2* sv = reinterpret_cast<const uint2*>(s+head); uint2* dv = reinterpret_cast<uint2*>(d+head); for(size_t i=tid;i<nw;i+=nt) dv[i] = sv[i]; }
    for(size_t i=head+nw*w+tid;i<n;i+=nt) d[i] = s[i];
#   else
    std::memcpy(dst, src, n);
#   endif
  }

  // overlapping ranges: the block copies in chunks of one word per thread, separated by barriers, walking in the safe direction
  ONIKA_HOST_DEVICE_FUNC static inline void cu_block_memmove(void* dst, const void* src, size_t n)
  {
#   if defined(__CUDA_ARCH__)
    uint8_t* d = static_cast<uint8_t*>(dst);
    const uint8_t* s = static_cast<const uint8_t*>(src);
    const unsigned tid = ONIKA_CU_THREAD_IDX, nt = ONIKA_CU_BLOCK_SIZE;
    if( d == s || n == 0 ) return;
    const bool fwd = ( d < s ) || ( d >= s + n );
    const size_t rounds = ( n + nt - 1 ) / nt;
    for(size_t r=0;r<rounds;r++)
    {
      const size_t chunk = fwd ? r : ( rounds - 1 - r );
      const size_t i = chunk * nt + tid;
      uint8_t b = 0;
      if( i < n ) b = s[i];
      ONIKA_CU_BLOCK_SYNC();
      if( i < n ) d[i] = b;
      ONIKA_CU_BLOCK_SYNC();
    }
#   else
    std::memmove(dst, src, n);
#   endif
  }

} }
