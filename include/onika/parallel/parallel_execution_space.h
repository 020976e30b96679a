// onika/parallel/parallel_execution_space.h -- the index space a parallel operation runs over:
// a 1D/2D/3D box [m_start,m_end), optionally (1D only) an explicit list of element coordinates.
#pragma once
#include <vector>
#include <cassert>
#include <cstddef>
#include <span>
#include <type_traits>
#include <algorithm>
#include <onika/oarray.h>
#include <onika/memory/allocator.h>

namespace onika { namespace parallel {

  template<unsigned int ND=1> struct ElementCoordT { using type = onika::oarray_t<ssize_t,ND>; };
  template<> struct ElementCoordT<1> { using type = ssize_t; };
  template<unsigned int ND> using element_coord_t = typename ElementCoordT<ND>::type;

  // dimensionality of a coordinate type: integral -> 1, oarray_t<.,N> -> N, anything else -> 0
  template<class T> static inline constexpr unsigned int element_coord_nd_v =
      std::is_integral_v<T> ? 1u : ( is_oarray_v<T> ? []{ if constexpr (is_oarray_v<T>) return (unsigned int)T::array_size; else return 0u; }() : 0u );

  inline constexpr size_t coord_range_size(ssize_t a, ssize_t b) { return ( b > a ) ? size_t( b - a ) : 0; }
  template<some_integral_oarray_t A>
  inline constexpr size_t coord_range_size(const A& a, const A& b)
  {
    if( A::array_size == 0 ) return 0;
    size_t n = 1;
    for(size_t d=0; d<A::array_size; d++) n *= ( b[d] > a[d] ) ? size_t( b[d] - a[d] ) : 0;
    return n;
  }

  // advance `it` inside [a,b), first coordinate fastest; returns false (and parks `it` at b) when the range is exhausted
  template<some_oarray_t A>
  inline constexpr bool coord_range_iterator_increment(A& it, const A& a, const A& b)
  {
    for(size_t d=0; d<A::array_size; d++)
    {
      if( ++it[d] < b[d] ) return true;
      it[d] = a[d];
    }
    it = b;
    return false;
  }
  inline bool coord_range_iterator_increment(ssize_t& it, ssize_t, ssize_t b)
  {
    if( ++it > b ) { it = b; return false; }
    return true;
  }

  template<class T> concept some_space_coord = some_integral_oarray_t<T> || std::is_integral_v<T>;

  // explicit (managed-memory) list of every coordinate of a box, in iteration order
  template<some_space_coord C>
  inline auto coord_range_to_list(const C& a, const C& b)
  {
    using Coord = element_coord_t< element_coord_nd_v<C> >;
    const size_t n = coord_range_size(a, b);
    memory::CudaMMArray<Coord> elements;
    if( n > 0 )
    {
      elements.resize(n);
      auto it = a; size_t k = 0;
      do { elements[k++] = it; } while( k < n && coord_range_iterator_increment(it, a, b) );
    }
    return elements;
  }

  template<some_space_coord C>
  inline size_t coord_range_to_array(ssize_t* out, size_t outsz, const C& a, const C& b)
  {
    constexpr size_t ND = element_coord_nd_v<C>;
    if( 2 * ND > outsz ) return 0;
    if constexpr ( std::is_integral_v<C> ) { out[0] = a; out[1] = b; }
    else { for(size_t d=0;d<ND;d++) { out[d] = a[d]; out[ND+d] = b[d]; } }
    return 2 * ND;
  }

  template<unsigned int _NDim=1, unsigned int _ElementListNDim=0, class _ElementListT = std::span< const element_coord_t<_ElementListNDim> > >
  struct ParallelExecutionSpace
  {
    static_assert( _NDim>=1 && _NDim<=3 && _ElementListNDim<=3 );
    static_assert( _ElementListNDim==0 || _NDim==1 , "Element lists are only supported for 1D parallel execution spaces" );
    static inline constexpr unsigned int NDim = _NDim;
    static inline constexpr unsigned int ElementListNDim = _ElementListNDim;
    static inline constexpr unsigned int SpaceNDim = ( ElementListNDim==0 ) ? NDim : ElementListNDim ;
    using space_coord_t = onika::oarray_t<ssize_t,SpaceNDim>;
    using coord_t = onika::oarray_t<ssize_t,NDim>;
    using element_list_t = _ElementListT;
    using element_t = std::remove_cv_t< std::remove_reference_t< decltype( _ElementListT{}[0] ) > >;
    coord_t m_start;
    coord_t m_end;
    element_list_t m_elements = {};
    static inline constexpr unsigned int space_nd() { return SpaceNDim; }
    inline constexpr size_t number_of_items() const { return coord_range_size( m_start , m_end ); }
  };

  template<class T> struct is_parallel_execution_space_t : std::false_type {};
  template<unsigned int N,unsigned int E,class L> struct is_parallel_execution_space_t< ParallelExecutionSpace<N,E,L> > : std::true_type {};
  template<class T> static inline constexpr bool is_parallel_execution_space_v = is_parallel_execution_space_t<T>::value;
  template<class T> concept SortOfParallelExecutionSpace = is_parallel_execution_space_v<T>;
  template<class T, class P> concept CompatibleParallelExecutionSpace = is_parallel_execution_space_v<T> && is_parallel_execution_space_v<P> && T::SpaceNDim == P::SpaceNDim;

  // Which elements of a second execution space touch, through `conflict_stencil` (relative offsets, e.g. the conflict map
  // of two operations' data accesses), an element of a first one?  Splits pes2 into the elements that must wait for pes1
  // (`pes2_dependent_elements`) and those that may run concurrently with it (`pes2_remaining_elements`), both in
  // iteration order.  (Declared but left empty in the reference, include/onika/parallel/parallel_execution_space.h:152-159.)
  // Spaces given by element lists are handled conservatively: every element of pes2 is reported dependent.
  template<SortOfParallelExecutionSpace PES1, CompatibleParallelExecutionSpace<PES1> PES2 >
  inline void concurrent_execution_space_elements( const PES1& pes1, const PES2& pes2
                                                 , std::span< element_coord_t<PES1::SpaceNDim> > conflict_stencil
                                                 , std::vector< element_coord_t<PES1::SpaceNDim> > & pes2_dependent_elements
                                                 , std::vector< element_coord_t<PES1::SpaceNDim> > & pes2_remaining_elements )
  {
    constexpr unsigned int ND = PES1::SpaceNDim;
    using Coord = element_coord_t<ND>;
    pes2_dependent_elements.clear();
    pes2_remaining_elements.clear();
    auto comp = [](const Coord& c, unsigned int d) -> ssize_t { if constexpr ( ND == 1 ) { (void)d; return c; } else return c[d]; };
    auto shifted_inside_pes1 = [&](const Coord& e, const Coord& s) -> bool
    {
      for(unsigned int d=0;d<ND;d++)
      {
        const ssize_t x = comp(e,d) + comp(s,d);
        if( x < pes1.m_start[d] || x >= pes1.m_end[d] ) return false;
      }
      return true;
    };
    const bool listed = ( PES1::ElementListNDim != 0 ) || ( PES2::ElementListNDim != 0 );
    if( pes2.number_of_items() == 0 ) return;
    auto it = pes2.m_start;
    do
    {
      Coord e;
      if constexpr ( ND == 1 ) e = it[0]; else e = it;
      bool dep = listed;
      for(size_t k=0; !dep && k<conflict_stencil.size(); k++) dep = shifted_inside_pes1( e , conflict_stencil[k] );
      ( dep ? pes2_dependent_elements : pes2_remaining_elements ).push_back( e );
    } while( coord_range_iterator_increment( it , pes2.m_start , pes2.m_end ) );
  }

} }
