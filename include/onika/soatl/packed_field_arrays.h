// onika/soatl/packed_field_arrays.h -- byte layout of packed field arrays and their allocators.
// Layout arithmetic is the runtime library's (onk_soatl_layout): column k starts where column k-1 ends, rounded up to
// the alignment, for a capacity that is a whole number of chunks -- byte-identical to the reference's calculator.
#pragma once
#include <cassert>
#include <cstdint>
#include <cstdlib>
#include <onika/cuda/cuda.h>
#include <onika/soatl/constants.h>
#include <onika/soatl/stdtypes.h>
#include <onika/soatl/traits.h>
#include <onika/soatl/field_id.h>
#include <onika/memory/allocator.h>
#include <onika_b200.h>

namespace onika { namespace soatl {

  namespace pfa_details
  {
    template<size_t A, class... ids>
    ONIKA_HOST_DEVICE_FUNC static inline size_t storage_size( FieldIds<ids...> , size_t capacity )
    {
      size_t cursor = 0;
      ( ... , ( cursor += ( capacity * sizeof( typename FieldId<ids>::value_type ) + ( A - 1 ) ) & ~( A - 1 ) ) );
      return cursor;
    }
  }

  template<size_t _A, size_t _C, typename Fids>
  ONIKA_HOST_DEVICE_FUNC static inline size_t pfa_storage_size(size_t capacity)
  {
    assert( capacity % _C == 0 );
    return pfa_details::storage_size<_A>( Fids{} , capacity );
  }
  template<size_t _A, size_t _C, typename id, typename... ids>
  ONIKA_HOST_DEVICE_FUNC static inline size_t pfa_pointer_offset(size_t capacity)
  {
    return pfa_details::storage_size<_A>( preceding_field_ids_t< id , ids... >{} , capacity );
  }

  class PackedFieldArraysAllocator
  {
  public:
    virtual size_t allocation_bytes(size_t n_elements) const =0;
    virtual void* allocate(size_t n_elements) const =0;
    virtual void deallocate(void* ptr, size_t n_elements) const =0;
    virtual bool is_gpu_addressable(void* ptr, size_t n_elements) const =0;
    virtual bool allocates_gpu_addressable() const =0;
    virtual void set_gpu_addressable_allocation(bool) =0;
    virtual ~PackedFieldArraysAllocator() = default;
  };

  template<class BaseAllocT, size_t Alignment, size_t ChunkSize, class... ids>
  class PackedFieldArraysAllocatorImpl : public PackedFieldArraysAllocator
  {
  public:
    PackedFieldArraysAllocatorImpl() = default;
    inline PackedFieldArraysAllocatorImpl(const BaseAllocT& a) : m_alloc(a) {}
    inline size_t allocation_bytes(size_t n) const override final { return pfa_storage_size<Alignment,ChunkSize,FieldIds<ids...> >( n ); }
    inline void* allocate(size_t n) const override final { return m_alloc.allocate( allocation_bytes(n) , std::max( Alignment , size_t(8) ) ); }
    inline void deallocate(void* p, size_t n) const override final { m_alloc.deallocate( p , allocation_bytes(n) ); }
    inline bool is_gpu_addressable(void* p, size_t n) const override final { return m_alloc.is_gpu_addressable( p , allocation_bytes(n) ); }
    inline bool allocates_gpu_addressable() const override final { return m_alloc.allocates_gpu_addressable(); }
    inline void set_gpu_addressable_allocation(bool yn) override final { m_alloc.set_gpu_addressable_allocation( yn ); }
    inline BaseAllocT& base_allocator() { return m_alloc; }
  private:
    BaseAllocT m_alloc;
  };

  template<class BaseAllocT, size_t A, size_t C, class Fids> struct PackedFieldArraysAllocatorImplFromFieldIds;
  template<class BaseAllocT, size_t A, size_t C, class... ids>
  struct PackedFieldArraysAllocatorImplFromFieldIds<BaseAllocT,A,C, FieldIds<ids...> > { using type = PackedFieldArraysAllocatorImpl<BaseAllocT,A,C,ids...>; };
  template<class BaseAllocT, size_t A, size_t C, class Fids>
  using pfa_allocator_impl_from_field_ids_t = typename PackedFieldArraysAllocatorImplFromFieldIds<BaseAllocT,A,C,Fids>::type;

} }
