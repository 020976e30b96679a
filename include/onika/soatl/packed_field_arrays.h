// onika/soatl/packed_field_arrays.h -- byte layout of packed field arrays and their allocators.
// Layout arithmetic is the runtime library's (onk_soatl_layout): column k starts where column k-1 ends, rounded up to
// the alignment, for a capacity that is a whole number of chunks -- byte-identical to the reference's calculator.
#pragma once
#include <cassert>
#include <cstdint>
#include <cstdlib>
#include <type_traits>
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
    ONIKA_HOST_DEVICE_FUNC static inline constexpr size_t storage_size( FieldIds<ids...> , size_t capacity )
    {
      size_t cursor = 0;
      ( ... , ( cursor += ( capacity * sizeof( typename FieldId<ids>::value_type ) + ( A - 1 ) ) & ~( A - 1 ) ) );
      return cursor;
    }
  }

  namespace pfa_details
  {
    // the reference's cross-check of its size calculator (include/onika/soatl/packed_field_arrays.h:105-118 there): bytes of
    // every column plus the padding that brings the next column back onto the alignment, summed field by field
    template<size_t A, size_t C, class... ids>
    ONIKA_HOST_DEVICE_FUNC static inline constexpr size_t check_size_for_capacity( std::integral_constant<size_t,A> , std::integral_constant<size_t,C> , FieldIds<ids...> , size_t capacity )
    {
      static_assert( A >= 1 && C >= 1 );
      size_t total = 0;
      ( ... , ( total += capacity * sizeof( typename FieldId<ids>::value_type ) , total += ( A - total % A ) % A ) );
      return total;
    }

    // size calculator type: storage_size(capacity) for a whole number of chunks.  The reference builds a compile-time
    // histogram of field sizes so that its host loop multiplies instead of adding (packed_field_arrays.h:36-103,120-158);
    // here the fold over the field list IS the constant-folded form, and the run-time twin is onk_soatl_layout of the C ABI.
    template<size_t A, size_t C, class Fids> struct SizeCalculator
    {
      static_assert( A >= 1 && C >= 1 && ( A & ( A - 1 ) ) == 0 , "alignment must be a power of two" );
      static inline constexpr size_t alignment = A , chunk_size = C;
      ONIKA_HOST_DEVICE_FUNC static inline constexpr size_t storage_size( size_t capacity ) { return pfa_details::storage_size<A>( Fids{} , capacity ); }
      template<class StreamT> static inline StreamT& print( StreamT& out ) { out << "SizeCalculator<" << A << "," << C << ">\n"; return out; }
    };
  }

  // same spelling and argument order as the reference (alignment, field ids, chunk size)
  template<size_t _A, typename Fids, size_t _C> using pfa_size_calculator_t = pfa_details::SizeCalculator<_A,_C,Fids>;

  template<size_t _A, size_t _C, typename Fids>
  ONIKA_HOST_DEVICE_FUNC static inline size_t pfa_storage_size(size_t capacity)
  {
    assert( capacity % _C == 0 );
    assert( ( pfa_size_calculator_t<_A,Fids,_C>::storage_size(capacity) == pfa_details::check_size_for_capacity( std::integral_constant<size_t,_A>{} , std::integral_constant<size_t,_C>{} , Fids{} , capacity ) ) );
    return pfa_size_calculator_t<_A,Fids,_C>::storage_size( capacity );
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
