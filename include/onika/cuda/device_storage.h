// onika/cuda/device_storage.h -- CudaDeviceStorage<T>: a shared handle on an array of T in device memory.
// Same interface as the reference handle (New / get / -> / * / reset / comparisons); the memory comes from the runtime
// library's device allocator (onk_alloc, include/onika_b200.h) and the reference count is atomic.
#pragma once
#include <atomic>
#include <cassert>
#include <cstddef>
#include <onika/cuda/cuda_context.h>

namespace onika { namespace cuda {

  struct CudaDeviceStorageShared
  {
    size_t m_array_size = 0;
    std::atomic<size_t> m_refcount { 0 };
  };

  template<class T>
  struct CudaDeviceStorage
  {
    T* m_ptr = nullptr;
    CudaDeviceStorageShared* m_shared = nullptr;

    CudaDeviceStorage() = default;
    inline CudaDeviceStorage(const CudaDeviceStorage& o) : m_ptr(o.m_ptr), m_shared(o.m_shared) { retain(); }
    inline CudaDeviceStorage(CudaDeviceStorage&& o) : m_ptr(o.m_ptr), m_shared(o.m_shared) { o.m_ptr = nullptr; o.m_shared = nullptr; }
    inline ~CudaDeviceStorage() { reset(); }

    static inline CudaDeviceStorage<T> New(CudaDevice&, size_t n = 1)
    {
      void* p = nullptr;
      ONIKA_B200_CHECK( onk_alloc( sizeof(T) * n , 256 , ONK_MEM_DEVICE , &p ) );
      CudaDeviceStorage<T> s;
      s.m_ptr = static_cast<T*>(p);
      s.m_shared = new CudaDeviceStorageShared{};
      s.m_shared->m_array_size = n;
      s.m_shared->m_refcount = 1;
      return s;
    }

    inline void reset()
    {
      if( m_shared != nullptr && m_shared->m_refcount.fetch_sub(1) == 1 )
      {
        if( m_ptr != nullptr ) ONIKA_B200_CHECK( onk_free( m_ptr , sizeof(T) * m_shared->m_array_size ) );
        delete m_shared;
      }
      m_ptr = nullptr;
      m_shared = nullptr;
    }

    inline CudaDeviceStorage& operator = (const CudaDeviceStorage& o)
    {
      if( m_shared != o.m_shared ) { reset(); m_ptr = o.m_ptr; m_shared = o.m_shared; retain(); }
      return *this;
    }
    inline CudaDeviceStorage& operator = (CudaDeviceStorage&& o)
    {
      if( this != &o ) { reset(); m_ptr = o.m_ptr; m_shared = o.m_shared; o.m_ptr = nullptr; o.m_shared = nullptr; }
      return *this;
    }

    inline bool operator == (const CudaDeviceStorage& o) const { return m_ptr == o.m_ptr; }
    inline bool operator == (const T* p) const { return m_ptr == p; }
    inline bool operator == (std::nullptr_t) const { return m_ptr == nullptr; }
    inline bool operator != (const CudaDeviceStorage& o) const { return m_ptr != o.m_ptr; }
    inline bool operator != (const T* p) const { return m_ptr != p; }
    inline bool operator != (std::nullptr_t) const { return m_ptr != nullptr; }

    inline const T& operator *  () const { assert(m_ptr != nullptr); return *m_ptr; }
    inline       T& operator *  ()       { assert(m_ptr != nullptr); return *m_ptr; }
    inline const T* operator -> () const { assert(m_ptr != nullptr); return  m_ptr; }
    inline       T* operator -> ()       { assert(m_ptr != nullptr); return  m_ptr; }
    inline const T* get() const { return m_ptr; }
    inline       T* get()       { return m_ptr; }

  private:
    inline void retain() { if( m_shared != nullptr ) m_shared->m_refcount.fetch_add(1); }
  };

} }
