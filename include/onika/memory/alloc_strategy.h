// onika/memory/alloc_strategy.h -- capacity growth policies of SOATL containers.  The arithmetic lives in the runtime
// library (onk_soatl_update_capacity) so that every front end (C++, Python) computes identical capacities.
#pragma once
#include <cstddef>
#include <onika_b200.h>

namespace onika { namespace memory {

  // exact fit, whole chunks; shrinks only when two chunks or more are free
  struct ChunkIncrementalAllocationStrategy
  {
    static inline size_t update_capacity(size_t s, size_t capacity, size_t chunksize)
    {
      const bool resize = ( s > capacity ) || ( s + 2 * chunksize <= capacity ) || ( s == 0 );
      return resize ? ( ( s + chunksize - 1 ) / chunksize ) * chunksize : capacity;
    }
  };

  // geometric growth / shrink-at-half, whole chunks (the default)
  struct ChunkLogAllocationStrategy
  {
    static inline size_t update_capacity(size_t s, size_t capacity, size_t chunksize)
    {
      return onk_soatl_update_capacity( s , capacity , chunksize );
    }
  };

  using DefaultAllocationStrategy = ChunkLogAllocationStrategy;

} }
