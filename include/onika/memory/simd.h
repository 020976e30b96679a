// onika/memory/simd.h -- alignment / chunk-size requirements used as SOATL defaults.
// Host SIMD on the reference's target is AVX-512 (64-byte vectors); the same constants are kept so layouts are
// byte-compatible with the reference's defaults (DEFAULT_ALIGNMENT 64, DEFAULT_CHUNK_SIZE 16).
#pragma once
#include <cstddef>
#include <cstdint>

namespace onika { namespace memory {

  static inline constexpr size_t SIMD_REGISTER_BYTES = 64;
  inline const char* simd_arch() { return "sm_100a (host layout: 64-byte vectors)"; }

  template<class T> struct SimdRequirements
  {
    static constexpr size_t alignment = SIMD_REGISTER_BYTES;
    static constexpr size_t chunksize = ( SIMD_REGISTER_BYTES / sizeof(T) ) > 0 ? ( SIMD_REGISTER_BYTES / sizeof(T) ) : 1;
  };

  // debug helpers: true when [p, p+n) ranges are pairwise disjoint / aligned
  template<class... T> inline bool check_pointers_aliasing(size_t n, T* ... p)
  {
    const uintptr_t b[] = { reinterpret_cast<uintptr_t>(p) ... };
    const size_t sz[] = { n * sizeof(T) ... };
    constexpr size_t N = sizeof...(T);
    for(size_t i=0;i<N;i++) for(size_t j=i+1;j<N;j++) if( b[i] < b[j] + sz[j] && b[j] < b[i] + sz[i] ) return false;
    return true;
  }
  template<class... T> inline bool check_simd_pointers(size_t n, T* ... p)
  {
    return check_pointers_aliasing(n, p...);
  }

} }
