// onika/soatl/field_arrays.h -- FieldArrays<A,C,ids...>: struct-of-arrays container in ONE packed allocation.
//
// Public interface and byte layout follow the reference container (column k at sum_{j<k} roundup(capacity*sizeof(T_j),A),
// capacity a whole number of chunks of C elements growing geometrically, 32-bit size/capacity by default,
// operator[](FieldId) -> aligned restrict pointer, resize/assign/push_back/set_tuple/read_tuple/write_tuple/
// operator[](i)/swap/shrink_to_fit/clear/copy_from/field_pointers/capture_pointers/apply_arrays/storage_ptr/storage_size/
// unsafe_make_view), so a raw dump of storage_ptr()/storage_size() is interchangeable with the reference's.
// Implementation: only the base pointer is stored; column addresses are a compile-time fold over the preceding field
// sizes (a handful of integer ops, usable in kernels).  Storage comes from the allocator passed in (default: managed
// memory from libonika_b200, device addressable and host dereferenceable).
#pragma once
#include <cassert>
#include <cstring>
#include <tuple>
#include <type_traits>
#include <utility>
#include <onika/cuda/cuda.h>
#include <onika/flat_tuple.h>
#include <onika/memory/alloc_strategy.h>
#include <onika/soatl/packed_field_arrays.h>
#include <onika/soatl/field_tuple.h>
#include <onika/soatl/field_pointer_tuple.h>
#include <onika/soatl/field_combiner.h>
#include <onika/soatl/traits.h>
#include <onika/soatl/copy.h>

namespace onika { namespace soatl {

  template<class FuncT , class PointerTuple , class Seq> struct FieldArrayCombinerAccessor;
  template<class FuncT , class PointerTuple , std::size_t ... I>
  struct FieldArrayCombinerAccessor< FuncT , PointerTuple , std::index_sequence<I...> >
  {
    const FuncT m_func;
    const PointerTuple m_pointers;
    ONIKA_HOST_DEVICE_FUNC inline auto operator [] (size_t i) const { return m_func( m_pointers.template get_nth<I>()[i] ... ); }
  };

  template<size_t _Alignment, size_t _ChunkSize, typename _DefaultAllocator, size_t _NbStoredPointers, typename... ids >
  struct FieldArraysWithAllocator
  {
    static_assert( _NbStoredPointers <= sizeof...(ids) , "Cannot store more pointers than fields" );
    static_assert( _Alignment > 0 && IsPowerOf2<_Alignment>::value , "alignment must be a power of 2" );
    static_assert( _ChunkSize > 0 , "Chunk size must be strictly positive" );

    static constexpr size_t NbStoredPointers = _NbStoredPointers;
    static constexpr size_t Alignment = _Alignment;
    static constexpr size_t AlignmentLog2 = Log2<_Alignment>::value;
    static constexpr size_t AlignmentLowMask = Alignment - 1;
    static constexpr size_t AlignmentHighMask = ~AlignmentLowMask;
    static constexpr size_t ChunkSize = _ChunkSize;
    static constexpr size_t TupleSize = sizeof...(ids);

    using TupleValueType = FieldTuple<ids...>;
    using FieldIdsTuple = std::tuple< FieldId<ids> ... >;
    using AllocStrategy = memory::DefaultAllocationStrategy;
    using DefaultAllocator = _DefaultAllocator;
    template<typename _id> using field_value_t = typename FieldId<_id>::value_type;
    template<typename _id> using field_pointer_t = field_value_t<_id> * __restrict__ ;

    template<class fid> struct HasFieldHelper { static constexpr bool value = find_index_of_id_v<fid,ids...> != bad_field_index; };
    template<class F , class... fids> struct HasFieldHelper< FieldCombiner<F,fids...> > { static constexpr bool value = ( ... && ( HasFieldHelper<fids>::value ) ); };
    template<class f> using HasField = std::integral_constant< bool , HasFieldHelper<f>::value >;
    template<class _id> static inline constexpr bool has_field( FieldId<_id> ) { return HasField<_id>::value; }
    template<class F , class... fids> static inline constexpr bool has_field( FieldCombiner<F,fids...> ) { return HasField< FieldCombiner<F,fids...> >::value; }

    static constexpr size_t alignment() { return Alignment; }
    static constexpr size_t chunksize() { return ChunkSize; }

    // ---------------- size / capacity ----------------
    ONIKA_HOST_DEVICE_FUNC inline size_t size() const { return m_size; }
    ONIKA_HOST_DEVICE_FUNC inline bool empty() const { return m_size == 0; }
    ONIKA_HOST_DEVICE_FUNC inline size_t capacity() const { return m_capacity; }
    ONIKA_HOST_DEVICE_FUNC inline size_t chunk_ceil() const { return ( ( size() + ChunkSize - 1 ) / ChunkSize ) * ChunkSize; }
    ONIKA_HOST_DEVICE_FUNC inline size_t minimal_capacity() const { return chunk_ceil(); }
    ONIKA_HOST_DEVICE_FUNC inline size_t payload_bytes() const { return ( 0 + ... + sizeof(field_value_t<ids>) ) * size(); }
    ONIKA_HOST_DEVICE_FUNC inline void* storage_ptr() const { return m_storage; }

    template<class Allocator = DefaultAllocator> inline size_t storage_size(const Allocator& alloc = Allocator{}) const { return alloc.allocation_bytes( capacity() ); }
    template<class Allocator = DefaultAllocator> inline size_t memory_bytes(const Allocator& alloc = Allocator{}) const { return storage_size(alloc) + sizeof(*this); }
    template<class Allocator = DefaultAllocator> inline bool is_gpu_addressable(const Allocator& alloc = Allocator{}) const { return alloc.is_gpu_addressable( storage_ptr() , capacity() ); }

    // ---------------- construction / destruction ----------------
    inline FieldArraysWithAllocator() = default;
    template<class Allocator = DefaultAllocator>
    inline FieldArraysWithAllocator( size_t n , const Allocator& alloc = Allocator{} ) { resize( n , alloc ); }
    template<class Allocator = DefaultAllocator>
    inline FieldArraysWithAllocator( const FieldArraysWithAllocator & other , const Allocator& alloc = Allocator{} ) { copy_from( other , alloc ); }
    inline FieldArraysWithAllocator( FieldArraysWithAllocator && o ) : m_size(o.m_size) , m_capacity(o.m_capacity) , m_storage(o.m_storage) { o.reset(); }
    inline FieldArraysWithAllocator& operator = ( FieldArraysWithAllocator && o )
    {
      clear();
      m_size = o.m_size; m_capacity = o.m_capacity; m_storage = o.m_storage;
      o.reset();
      return *this;
    }
    // containers must be emptied (clear()/resize(0)) before they die: the allocator instance is not stored
    inline ~FieldArraysWithAllocator() { assert( empty() && capacity() == 0 && storage_ptr() == nullptr ); }

    // ---------------- resize family ----------------
    template<class Allocator = DefaultAllocator>
    inline void resize(size_t s , const Allocator& alloc = Allocator{} )
    {
      if( s == m_size ) return;
      const size_t cap = AllocStrategy::update_capacity( s , capacity() , ChunkSize );
      if( cap != m_capacity ) reallocate( cap , alloc );
      m_size = s;
    }
    template<class Allocator = DefaultAllocator> inline void clear( const Allocator& alloc = Allocator{} ) { resize( 0 , alloc ); }
    template<class Allocator = DefaultAllocator>
    inline void shrink_to_fit( const Allocator& alloc = Allocator{} , bool force_reallocate=false ) { reallocate( minimal_capacity() , alloc , force_reallocate ); }
    template<class Allocator = DefaultAllocator>
    inline void assign(size_t s , const TupleValueType& value , const Allocator& alloc = Allocator{} )
    {
      resize( s , alloc );
      for(size_t i=0;i<s;i++) set_tuple( i , value );
    }
    template<class Allocator = DefaultAllocator>
    inline void push_back( const TupleValueType& value , const Allocator& alloc = Allocator{} )
    {
      const size_t s = size();
      resize( s+1 , alloc );
      set_tuple( s , value );
    }
    template<class Allocator = DefaultAllocator>
    inline FieldArraysWithAllocator& copy_from ( const FieldArraysWithAllocator & other , const Allocator& alloc = Allocator{} )
    {
      const size_t n = other.size();
      m_size = 0;                       // nothing to preserve across the next resize
      resize( n , alloc );
      ( ... , std::memcpy( (*this)[FieldId<ids>{}] , other[FieldId<ids>{}] , n * sizeof(field_value_t<ids>) ) );
      return *this;
    }

    // ---------------- element access ----------------
    ONIKA_HOST_DEVICE_FUNC inline void set_tuple ( size_t i, const TupleValueType& v ) { assert( i < capacity() ); ( ... , ( (*this)[FieldId<ids>{}][i] = v[FieldId<ids>{}] ) ); }
    ONIKA_HOST_DEVICE_FUNC inline void set_tuple ( size_t i, const field_value_t<ids>& ... args ) { set_tuple( i , TupleValueType(args...) ); }
    inline void set_tuple ( size_t i, const std::tuple< field_value_t<ids> ... > & v ) { set_tuple( i , TupleValueType(v) ); }
    template<typename... other> ONIKA_HOST_DEVICE_FUNC inline void read_tuple( size_t i, FieldTuple<other...>& tp ) const
    {
      assert( i < capacity() );
      ( ... , read_one<other>( i , tp ) );
    }
    template<typename... other> ONIKA_HOST_DEVICE_FUNC inline void write_tuple( size_t i, const FieldTuple<other...>& tp ) const
    {
      assert( i < capacity() );
      ( ... , write_one<other>( i , tp ) );
    }
    ONIKA_HOST_DEVICE_FUNC inline TupleValueType operator [] ( size_t i ) const { assert( i < capacity() ); return TupleValueType( (*this)[FieldId<ids>{}][i] ... ); }
    inline void swap( size_t a, size_t b ) { auto t = (*this)[a]; set_tuple( a , (*this)[b] ); set_tuple( b , t ); }
    static inline TupleValueType make_tuple(const field_value_t<ids> & ... args) { return TupleValueType( args ... ); }

    // ---------------- column access ----------------
    template<typename _id>
    ONIKA_HOST_DEVICE_FUNC ONIKA_ALWAYS_INLINE field_pointer_t<_id> operator [] ( FieldId<_id> ) const
    {
      static_assert( find_index_of_id_v<_id,ids...> != bad_field_index , "no such field in this container" );
      return column_pointer<_id>();
    }
    template<typename _id>
    ONIKA_HOST_DEVICE_FUNC ONIKA_ALWAYS_INLINE field_pointer_t<_id> field_pointer_or_null ( FieldId<_id> ) const
    {
      if constexpr ( find_index_of_id_v<_id,ids...> != bad_field_index ) return column_pointer<_id>();
      else return nullptr;
    }
    template<class FuncT , class... fids>
    ONIKA_HOST_DEVICE_FUNC ONIKA_ALWAYS_INLINE auto operator [] ( const FieldCombiner<FuncT,fids...>& c ) const
    {
      using PointerTuple = FlatTuple< const field_value_t<fids> * ... >;
      return FieldArrayCombinerAccessor<FuncT, PointerTuple , std::make_index_sequence<sizeof...(fids)> > { c.m_func , PointerTuple( (*this)[FieldId<fids>{}] ... ) };
    }
    inline FieldPointerTuple<Alignment,ChunkSize,ids...> field_pointers() const { return FieldPointerTuple<Alignment,ChunkSize,ids...>( (*this)[FieldId<ids>{}] ... ); }
    template<typename... other>
    ONIKA_HOST_DEVICE_FUNC inline void capture_pointers( FieldPointerTuple<Alignment,ChunkSize,other...>& pt ) const
    {
      ( ... , ( pt.set_pointer( FieldId<other>{} , field_pointer_or_null( FieldId<other>{} ) ) ) );
    }
    template<class FuncT> inline void apply_arrays(FuncT f) { ( ... , ( f( (*this)[FieldId<ids>{}] , size() ) ) ); }

    // ---------------- raw storage manipulation (serialisation, views) ----------------
    template<class Allocator = DefaultAllocator>
    inline void unsafe_storage_free( const Allocator& alloc = Allocator{} )
    {
      if( m_storage != nullptr ) { alloc.deallocate( m_storage , capacity() ); m_storage = nullptr; }
    }
    inline void unsafe_make_view(void * ptr , size_t sz) { assert( size()==0 && capacity()==0 && m_storage==nullptr ); m_size = m_capacity = sz; m_storage = ptr; }
    inline void unsafe_reset() { reset(); }

  private:
    inline void reset() { m_size = 0; m_capacity = 0; m_storage = nullptr; }

    template<typename _id>
    ONIKA_HOST_DEVICE_FUNC ONIKA_ALWAYS_INLINE field_pointer_t<_id> column_pointer() const
    {
      if( m_storage == nullptr ) return nullptr;
      const size_t off = pfa_details::storage_size<Alignment>( preceding_field_ids_t< _id , ids... >{} , size_t(m_capacity) );
      return reinterpret_cast< field_value_t<_id>* >( static_cast<uint8_t*>( m_storage ) + off );
    }
    template<class oid, typename... other> ONIKA_HOST_DEVICE_FUNC inline void read_one( size_t i, FieldTuple<other...>& tp ) const
    {
      if constexpr ( find_index_of_id_v<oid,ids...> != bad_field_index ) tp[ FieldId<oid>{} ] = (*this)[FieldId<oid>{}][i];
    }
    template<class oid, typename... other> ONIKA_HOST_DEVICE_FUNC inline void write_one( size_t i, const FieldTuple<other...>& tp ) const
    {
      if constexpr ( find_index_of_id_v<oid,ids...> != bad_field_index ) (*this)[FieldId<oid>{}][i] = tp[ FieldId<oid>{} ];
    }

    template<class Allocator = DefaultAllocator>
    inline void reallocate(size_t new_capacity , const Allocator& alloc = Allocator{} , bool force = false)
    {
      assert( new_capacity % ChunkSize == 0 );
      if( new_capacity == capacity() && ! force ) return;
      if( new_capacity == 0 )
      {
        if( m_storage != nullptr ) alloc.deallocate( m_storage , capacity() );
        reset();
        return;
      }
      void* fresh = alloc.allocate( new_capacity );
      const size_t keep = ( m_storage != nullptr ) ? std::min( new_capacity , size_t(m_size) ) : 0;   // storage freed beforehand => nothing to carry over
      if( keep > 0 )
      {
        FieldArraysWithAllocator view;
        view.unsafe_make_view( fresh , new_capacity );
        ( ... , std::memcpy( view[FieldId<ids>{}] , (*this)[FieldId<ids>{}] , keep * sizeof(field_value_t<ids>) ) );
        view.unsafe_reset();
      }
      if( m_storage != nullptr ) alloc.deallocate( m_storage , capacity() );
      m_storage = fresh;
      m_capacity = new_capacity;
    }

    array_size_t m_size = 0;
    array_size_t m_capacity = 0;
    void* m_storage = nullptr;
  };

  template<size_t A, size_t C, typename... ids >
  using FieldArrays = FieldArraysWithAllocator< A, C, PackedFieldArraysAllocatorImpl<memory::DefaultAllocator,A,C,ids...> , 1 , ids... >;

  template<typename... ids>
  inline FieldArrays<memory::DEFAULT_ALIGNMENT,memory::DEFAULT_CHUNK_SIZE,ids...> make_hybrid_field_arrays(const FieldId<ids>& ...) { return {}; }
  template<size_t A, size_t C,typename... ids>
  inline FieldArrays<A,C,ids...> make_hybrid_field_arrays(cst::align<A>, cst::chunk<C>, const FieldId<ids>& ...) { return {}; }

  template<size_t A, size_t C, typename Al, size_t N, typename... Ids >
  struct IsFieldArrays< FieldArraysWithAllocator<A,C,Al,N,Ids...> >  : public std::true_type {};

} }
