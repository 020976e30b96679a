// onika/soatl/constants.h -- compile-time alignment / chunk tags.
#pragma once
#include <cstdlib>
namespace onika { namespace soatl {
  template<size_t N> struct Log2 { static constexpr size_t value = Log2<N/2>::value + 1; };
  template<> struct Log2<1> { static constexpr size_t value = 0; };
  template<> struct Log2<0> { static constexpr size_t value = 0; };
  template<size_t n> struct IsPowerOf2 { static constexpr bool value = ( n != 0 ) && ( ( n & ( n - 1 ) ) == 0 ); };
  namespace cst
  {
    template<size_t A> struct align { static_assert(A>0 && IsPowerOf2<A>::value,"alignment must be a power of 2"); static constexpr size_t value = A; };
    template<size_t C> struct chunk { static_assert(C>0,"chunk size must be > 0"); static constexpr size_t value = C; };
    template<size_t> struct at {};
    template<size_t> struct count {};
    static constexpr align<1> unaligned{}; static constexpr align<1> aligned_1{}; static constexpr align<2> aligned_2{};
    static constexpr align<4> aligned_4{}; static constexpr align<8> aligned_8{}; static constexpr align<16> aligned_16{};
    static constexpr align<32> aligned_32{}; static constexpr align<64> aligned_64{}; static constexpr align<256> aligned_256{};
    static constexpr chunk<1> no_chunk{}; static constexpr chunk<1> chunk_1{}; static constexpr chunk<2> chunk_2{};
    static constexpr chunk<4> chunk_4{}; static constexpr chunk<8> chunk_8{}; static constexpr chunk<16> chunk_16{};
  }
} }
