// onika/memory/allocator.h -- host-visible allocators of the B200 front end.
//
// GenericHostAllocator keeps the reference's interface (allocate(bytes,align) / deallocate(ptr,bytes) /
// is_gpu_addressable / allocates_gpu_addressable / set_gpu_addressable_allocation, policy MALLOC | CUDA_HOST).
// CUDA_HOST memory is CUDA managed memory from the runtime library (onk_alloc(ONK_MEM_MANAGED)), so pointers stay
// dereferenceable on the host as plugins expect (e.g. tutorials read results on the host after synchronize).
// Instead of the reference's 12-byte {size,flags} trailer after every payload, the library keeps a side table keyed
// by pointer; deallocate(ptr,bytes) still checks the byte count (memory_info() reads that table).
// Deliberate difference: a default-constructed GenericHostAllocator allocates GPU-ADDRESSABLE (managed) memory -- the
// reference defaults to MALLOC and applications switch containers over with set_gpu_addressable_allocation(true) once a GPU
// is enabled; this library has no run without a GPU, so containers are device-ready from the start.  The policy switch itself
// (set_gpu_addressable_allocation(false) -> posix_memalign) works as in the reference.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <onika/memory/simd.h>
#include <onika/cuda/cuda.h>
#include <onika/cuda/cuda_error.h>
#include <onika_b200.h>

namespace onika { namespace memory {

  enum class HostAllocationPolicy { MALLOC = 0 , CUDA_HOST = 1 };

  static inline constexpr HostAllocationPolicy CUDA_FALLBACK_ALLOC_POLICY = HostAllocationPolicy::CUDA_HOST;

# ifndef ONIKA_DEFAULT_ALIGNMENT
  static inline constexpr size_t DEFAULT_ALIGNMENT = SimdRequirements<double>::alignment;
# else
  static inline constexpr size_t DEFAULT_ALIGNMENT = ONIKA_DEFAULT_ALIGNMENT;
# endif
# ifndef ONIKA_DEFAULT_CHUNK_SIZE
  static inline constexpr size_t DEFAULT_CHUNK_SIZE = SimdRequirements<float>::chunksize;
# else
  static inline constexpr size_t DEFAULT_CHUNK_SIZE = ONIKA_DEFAULT_CHUNK_SIZE;
# endif
# ifndef ONIKA_MINIMUM_CUDA_ALIGNMENT
  static inline constexpr size_t MINIMUM_CUDA_ALIGNMENT = 256;
# else
  static inline constexpr size_t MINIMUM_CUDA_ALIGNMENT = ONIKA_MINIMUM_CUDA_ALIGNMENT;
# endif

  // what the allocator knows about a block it handed out (the reference decodes this from the trailer it writes after the payload)
  struct MemoryChunkInfo
  {
    void * alloc_base = nullptr;
    size_t alloc_size = 0;
    uint32_t alloc_flags = 0;          // policy in the low byte, alignment above
    inline HostAllocationPolicy mem_type() const { return static_cast<HostAllocationPolicy>( alloc_flags & 0xFF ); }
    inline unsigned int alignment() const { return alloc_flags >> 8; }
    inline void * base_ptr() const { return alloc_base; }
    inline size_t size() const { return alloc_size; }
  };

  struct GenericHostAllocator
  {
    static inline constexpr size_t DefaultAlignBytes = std::max( MINIMUM_CUDA_ALIGNMENT , DEFAULT_ALIGNMENT );
    static inline bool s_enable_debug_log = false;
    static inline void set_debug_log(bool b) { s_enable_debug_log = b; }
    static inline bool s_enable_cuda = true;
    static inline bool cuda_enabled() { return s_enable_cuda; }
    static inline void set_cuda_enabled(bool yn) { s_enable_cuda = yn; }

    HostAllocationPolicy m_alloc_policy = HostAllocationPolicy::CUDA_HOST;

    inline HostAllocationPolicy get_policy() const { return cuda_enabled() ? m_alloc_policy : HostAllocationPolicy::MALLOC; }
    inline bool allocates_gpu_addressable() const { return get_policy() == HostAllocationPolicy::CUDA_HOST; }
    inline void set_gpu_addressable_allocation(bool yn) { m_alloc_policy = yn ? HostAllocationPolicy::CUDA_HOST : HostAllocationPolicy::MALLOC; }
    inline bool operator == (const GenericHostAllocator& o) const { return m_alloc_policy == o.m_alloc_policy; }

    // base / size / policy / alignment of a block of `bytes` bytes handed out by allocate(); the alignment reported is the
    // largest power of two (up to 4096) the address satisfies
    inline MemoryChunkInfo memory_info(void* ptr, size_t bytes) const
    {
      MemoryChunkInfo info;
      if( ptr == nullptr ) return info;
      int gpu = 0;
      ONIKA_B200_CHECK( onk_is_gpu_addressable( ptr , &gpu ) );
      unsigned int a = 1; while( a < 4096u && ( reinterpret_cast<uintptr_t>(ptr) % ( 2u * a ) ) == 0 ) a *= 2u;
      info.alloc_base = ptr; info.alloc_size = bytes;
      info.alloc_flags = ( a << 8 ) | static_cast<uint32_t>( gpu ? HostAllocationPolicy::CUDA_HOST : HostAllocationPolicy::MALLOC );
      return info;
    }

    inline void* allocate(size_t bytes, size_t align) const
    {
      if( bytes == 0 ) return nullptr;
      void* p = nullptr;
      if( get_policy() == HostAllocationPolicy::CUDA_HOST )
      {
        ONIKA_B200_CHECK( onk_alloc( bytes , align , ONK_MEM_MANAGED , &p ) );
      }
      else
      {
        align = std::max( align , sizeof(void*) );
        if( posix_memalign( &p , align , bytes ) != 0 ) { std::fprintf(stderr,"Allocation failed. aborting.\n"); std::abort(); }
      }
      if( s_enable_debug_log ) std::fprintf( stderr , "GenericHostAllocator: allocate %zu bytes, alignment %zu, %s -> %p\n" , bytes , align , get_policy()==HostAllocationPolicy::CUDA_HOST ? "managed" : "malloc" , p );
      return p;
    }

    inline void deallocate(void* ptr, size_t bytes) const
    {
      if( ptr == nullptr ) return;
      int gpu = 0;
      ONIKA_B200_CHECK( onk_is_gpu_addressable( ptr , &gpu ) );
      if( gpu ) ONIKA_B200_CHECK( onk_free( ptr , bytes ) );
      else std::free( ptr );
    }

    inline bool is_gpu_addressable(void* ptr, size_t) const
    {
      int gpu = 0;
      ONIKA_B200_CHECK( onk_is_gpu_addressable( ptr , &gpu ) );
      return gpu != 0;
    }
  };

  // STL-compatible allocator of host-visible, device-addressable memory
  template<class T>
  struct CudaManagedAllocator
  {
    using value_type = T;
    CudaManagedAllocator() = default;
    template<class U> CudaManagedAllocator(const CudaManagedAllocator<U>&) {}
    static inline T* allocate(std::size_t n)
    {
      constexpr size_t al = std::max( alignof(T) , MINIMUM_CUDA_ALIGNMENT );
      return static_cast<T*>( GenericHostAllocator{ CUDA_FALLBACK_ALLOC_POLICY }.allocate( sizeof(T) * n , al ) );
    }
    static inline void deallocate(T* p, std::size_t n) { GenericHostAllocator{ CUDA_FALLBACK_ALLOC_POLICY }.deallocate( p , sizeof(T) * n ); }
    template<class U> inline bool operator == (const CudaManagedAllocator<U>&) const { return true; }
    template<class U> inline bool operator != (const CudaManagedAllocator<U>&) const { return false; }
  };

  using DefaultAllocator = GenericHostAllocator;

  template<class T>
  struct NullAllocator
  {
    using value_type = T;
    ONIKA_HOST_DEVICE_FUNC inline T* allocate(std::size_t) { return nullptr; }
    ONIKA_HOST_DEVICE_FUNC inline void deallocate(T*, std::size_t) {}
    template<class U> ONIKA_HOST_DEVICE_FUNC inline bool operator == (const U&) const { return false; }
    ONIKA_HOST_DEVICE_FUNC inline bool operator == (const NullAllocator<T>&) const { return true; }
    template<class U> ONIKA_HOST_DEVICE_FUNC inline bool operator != (const U& o) const { return !( *this == o ); }
  };

  // host/device visible std::vector
  template<class T> using CudaMMVector = std::vector< T , CudaManagedAllocator<T> >;

  // fixed array in managed memory; resize() reallocates and does NOT keep the elements (as in the reference)
  template<class T>
  struct CudaMMArray
  {
    T* m_data_pointer = nullptr;
    size_t m_size = 0;
    CudaMMArray() = default;
    CudaMMArray(const CudaMMArray&) = delete;
    CudaMMArray& operator = (const CudaMMArray&) = delete;
    inline CudaMMArray(CudaMMArray&& o) : m_data_pointer(o.m_data_pointer), m_size(o.m_size) { o.m_data_pointer = nullptr; o.m_size = 0; }
    ONIKA_HOST_DEVICE_FUNC inline const T& operator [] (size_t i) const { return m_data_pointer[i]; }
    ONIKA_HOST_DEVICE_FUNC inline T& operator [] (size_t i) { return m_data_pointer[i]; }
    ONIKA_HOST_DEVICE_FUNC inline size_t size() const { return m_size; }
    inline void resize(size_t sz)
    {
      if( m_data_pointer != nullptr ) CudaManagedAllocator<T>::deallocate( m_data_pointer , m_size );
      m_size = sz;
      m_data_pointer = ( sz > 0 ) ? CudaManagedAllocator<T>::allocate( sz ) : nullptr;
    }
    inline void clear() { resize(0); }
    inline T* begin() const { return m_data_pointer; }
    inline T* end() const { return m_data_pointer + m_size; }
    inline ~CudaMMArray() { clear(); }
  };

# ifndef __CUDACC__
# define ONIKA_ASSUME_ALIGNED(x) x = ( decltype(x) __restrict__ ) __builtin_assume_aligned( x , ::onika::memory::DEFAULT_ALIGNMENT )
# define ONIKA__bultin_assume_aligned(x,a) __builtin_assume_aligned( x , a )
# else
# define ONIKA_ASSUME_ALIGNED(x) while(false)
# define ONIKA__bultin_assume_aligned(x,a) (x)
# endif

} }
