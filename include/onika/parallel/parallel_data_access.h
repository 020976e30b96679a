// onika/parallel/parallel_data_access.h -- declarative data-access stencils attached to parallel operations
// (queue << access << op).  64-byte POD descriptors, at most 2 per operation and 28 stencil elements each, each element
// packing an access mode and a relative (i,j,k) offset in {-1,0,1}^3 into one byte.
#pragma once
#include <cassert>
#include <cmath>
#include <concepts>
#include <cstdint>
#include <cstring>
#include <string_view>
#include <onika/oarray.h>

namespace onika { namespace parallel {

  struct AccessStencilElement
  {
    static inline constexpr uint8_t NA = 0x00;
    static inline constexpr uint8_t RO = 0x01;
    static inline constexpr uint8_t WO = 0x02;
    static inline constexpr uint8_t RW = 0x03;
    static inline constexpr const char * ACC_STR[4] = { "NA","RO","WO","RW" };
    uint8_t m_relative_access = 0;   // bits 0-1 mode, 2-3 i+1, 4-5 j+1, 6-7 k+1
    inline constexpr AccessStencilElement( unsigned int mode = RO , int ri=0 , int rj=0 , int rk=0 )
      : m_relative_access( uint8_t( ( mode & 3u ) | ( unsigned(ri+1) << 2 ) | ( unsigned(rj+1) << 4 ) | ( unsigned(rk+1) << 6 ) ) ) {}
    inline constexpr unsigned int mode() const { return m_relative_access & 3u; }
    inline constexpr const char* mode_str() const { return ACC_STR[mode()]; }
    inline constexpr int ri() const { return int( ( m_relative_access >> 2 ) & 3u ) - 1; }
    inline constexpr int rj() const { return int( ( m_relative_access >> 4 ) & 3u ) - 1; }
    inline constexpr int rk() const { return int( ( m_relative_access >> 6 ) & 3u ) - 1; }
  };

  struct ParallelDataAccess
  {
    static inline constexpr size_t MAX_NAME_LENGTH = 16;
    static inline constexpr size_t MAX_STENCIL_SIZE = 28;
    const void * m_data_ptr = nullptr;   // identity of the accessed data (array start, object address, ...)
    uint32_t m_ndim = 0;                 // 0: not correlated with the execution space
    onika::inplace_vector<AccessStencilElement,MAX_STENCIL_SIZE> m_stencil;
    char m_name[MAX_NAME_LENGTH] = {'\0',};
    inline void set_ndim(unsigned int nd) { assert( nd <= 3 ); m_ndim = nd; }
    inline unsigned int ndim() const { return m_ndim; }
    inline void reset() { m_data_ptr = nullptr; m_ndim = 0; m_stencil.clear(); m_name[0] = '\0'; }
    inline void set_name(std::string_view sv)
    {
      const size_t n = ( sv.size() < MAX_NAME_LENGTH-1 ) ? sv.size() : MAX_NAME_LENGTH-1;
      if( n > 0 ) std::memcpy( m_name , sv.data() , n );
      m_name[n] = '\0';
    }
    inline void set_address(const void * p) { assert( p != nullptr ); m_data_ptr = p; }
    inline const void* address() const { return m_data_ptr; }
    inline const char* name() const { return m_name; }
  };
  static_assert( sizeof(ParallelDataAccess) == 64 );

  static inline constexpr size_t MAX_PARALLEL_DATA_ACCESSES = 2;
  using ParallelDataAccessVector = onika::inplace_vector<ParallelDataAccess,MAX_PARALLEL_DATA_ACCESSES>;
  static inline constexpr size_t MAX_CONFLICT_MAP_SIZE = 128;
  using AccessConflictMap = onika::inplace_vector< oarray_t<int,3> , MAX_CONFLICT_MAP_SIZE >;

  // two access sets may conflict when they name the same data and their combined access modes intersect: the reference's
  // test (src/parallel/parallel_data_access.cpp:28-45), kept bit for bit because the reference PRINTS its outcome.  Note what
  // it answers: with RO = 1 and WO = 2 a pure writer and a pure reader of the same data do not "conflict" while two readers
  // do -- so it must not be used to ORDER operations; data_access_dependency() below is the ordering rule.
  inline bool concurrent_data_access(const ParallelDataAccessVector& a , const ParallelDataAccessVector& b )
  {
    auto modes = [](const ParallelDataAccess& d) { unsigned m = 0; for(auto e : d.m_stencil) m |= e.mode(); return m; };
    for(const auto & da : a) for(const auto & db : b)
      if( da.m_data_ptr == db.m_data_ptr && ( modes(da) & modes(db) ) != 0 ) return true;
    return false;
  }

  // must operation b (queued after a) wait for a ?  Same data, both touch it, at least one writes: read-after-write,
  // write-after-read and write-after-write all order; two readers do not.  This is the element rule of
  // concurrent_data_access_conflict_map applied to whole access sets, and what the lane placement uses.
  inline bool data_access_dependency(const ParallelDataAccessVector& a , const ParallelDataAccessVector& b )
  {
    auto modes = [](const ParallelDataAccess& d) { unsigned m = 0; for(auto e : d.m_stencil) m |= e.mode(); return m; };
    for(const auto & da : a) for(const auto & db : b)
    {
      const unsigned ma = modes(da) , mb = modes(db);
      if( da.m_data_ptr == db.m_data_ptr && ma != 0 && mb != 0 && ( ( ma | mb ) & AccessStencilElement::WO ) != 0 ) return true;
    }
    return false;
  }

  // relative offsets at which an element of one operation touches what an element of the other writes
  inline void concurrent_data_access_conflict_map(const ParallelDataAccessVector& a , const ParallelDataAccessVector& b , AccessConflictMap & out )
  {
    out.clear();
    for(const auto & da : a) for(const auto & db : b)
    {
      if( da.m_data_ptr != db.m_data_ptr ) continue;
      for(auto ea : da.m_stencil) for(auto eb : db.m_stencil)
      {
        const bool both_touch = ea.mode() != 0 && eb.mode() != 0;
        const bool one_writes = ( ( ea.mode() | eb.mode() ) & AccessStencilElement::WO ) != 0;
        if( both_touch && one_writes && out.size() < MAX_CONFLICT_MAP_SIZE )
          out.push_back( oarray_t<int,3>{ { ea.ri()-eb.ri() , ea.rj()-eb.rj() , ea.rk()-eb.rk() } } );
      }
    }
  }

  inline ParallelDataAccess local_access(const void * p , unsigned int nd=3, uint8_t mode = AccessStencilElement::RW , std::string_view name = {} )
  {
    ParallelDataAccess d; d.m_data_ptr = p; d.m_ndim = nd; d.m_stencil.push_back( AccessStencilElement{mode} ); d.set_name(name); return d;
  }
  inline ParallelDataAccess access_cross_2d(const void * p , uint8_t center = AccessStencilElement::RW, uint8_t side = AccessStencilElement::RO , std::string_view name = {} )
  {
    ParallelDataAccess d; d.m_data_ptr = p; d.m_ndim = 2;
    d.m_stencil.push_back( {center,0,0} );
    d.m_stencil.push_back( {side,-1,0} ); d.m_stencil.push_back( {side,1,0} );
    d.m_stencil.push_back( {side,0,-1} ); d.m_stencil.push_back( {side,0,1} );
    d.set_name(name); return d;
  }
  inline ParallelDataAccess access_cross_3d(const void * p , uint8_t center = AccessStencilElement::RW, uint8_t side = AccessStencilElement::RO , std::string_view name = {} )
  {
    ParallelDataAccess d; d.m_data_ptr = p; d.m_ndim = 3;
    d.m_stencil.push_back( {center,0,0,0} );
    d.m_stencil.push_back( {side,-1,0,0} ); d.m_stencil.push_back( {side,1,0,0} );
    d.m_stencil.push_back( {side,0,-1,0} ); d.m_stencil.push_back( {side,0,1,0} );
    d.m_stencil.push_back( {side,0,0,-1} ); d.m_stencil.push_back( {side,0,0,1} );
    d.set_name(name); return d;
  }
  inline ParallelDataAccess access_nbh_3d(const void * p , uint8_t center = AccessStencilElement::RW, uint8_t side = AccessStencilElement::RO , std::string_view name = {} )
  {
    ParallelDataAccess d; d.m_data_ptr = p; d.m_ndim = 3;
    for(int k=-1;k<=1;k++) for(int j=-1;j<=1;j++) for(int i=-1;i<=1;i++)
      d.m_stencil.push_back( { ( i==0 && j==0 && k==0 ) ? center : side , i , j , k } );
    d.set_name(name); return d;
  }

} }
