// onika/parallel/multi_gpu.h -- the cross-GPU step of the hot path for plugin code: one process per GPU of a node, the only
// exchange is a handful of fp64 partials per reduction.
//
// Reference counterpart: mpi::all_reduce_multi( comm , op , T{} , a , b , c ... ) (include/onika/mpi/all_reduce_multi.h:30-37:
// k scalars packed into one MPI_Allreduce with one operator).  Here the "communicator" is a MultiGpuExchange: exchange buffers
// in device memory, mapped into every peer process through CUDA IPC, and the all-reduce is one 128-thread kernel per rank over
// NVLink peer memory (onk_xchg_allreduce_f64) that folds the ranks' values in RANK ORDER: every rank gets the same bits and
// sums are reproducible.  No MPI, no collective library: the 64-byte IPC handles travel once, at construction, through any
// all-gather the application already has (an MPI_Allgather, a socket, ...) or through the file rendezvous below.
#pragma once
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>
#include <onika/cuda/cuda_error.h>
#include <onika/log.h>
#include <onika_b200.h>

namespace onika { namespace parallel {

  enum class MultiGpuOp : int { MIN = ONK_OP_MIN , MAX = ONK_OP_MAX , SUM = ONK_OP_SUM };

  // all-gather of `bytes` bytes per rank through files in a directory every process of the node can see (/dev/shm, /tmp):
  // rank r writes <dir>/<tag>.<r> atomically (write + rename) and reads the others' as they appear.  A transport for
  // programs that have none; anything with all-gather semantics can be passed to MultiGpuExchange instead.  The files stay
  // behind (no rank knows when the last reader is done): use a directory or tag that is new for every run (job id, start
  // time agreed through the launcher), otherwise the handles of an earlier run would be read.
  inline bool file_allgather( const std::string& dir , const std::string& tag , int rank , int world , const void* mine , void* all , size_t bytes , double timeout_s = 120.0 )
  {
    const std::string base = dir + "/" + tag + ".";
    {
      const std::string tmp = base + std::to_string(rank) + ".tmp";
      std::ofstream f( tmp , std::ios::binary ); f.write( (const char*) mine , std::streamsize(bytes) ); f.close();
      if( std::rename( tmp.c_str() , ( base + std::to_string(rank) ).c_str() ) != 0 ) return false;
    }
    const auto t0 = std::chrono::steady_clock::now();
    for(int r=0;r<world;r++)
    {
      for(;;)
      {
        std::ifstream f( base + std::to_string(r) , std::ios::binary );
        if( f.good() ) { f.read( (char*) all + size_t(r) * bytes , std::streamsize(bytes) ); if( size_t(f.gcount()) == bytes ) break; }
        if( std::chrono::duration<double>( std::chrono::steady_clock::now() - t0 ).count() > timeout_s ) return false;
        std::this_thread::sleep_for( std::chrono::milliseconds(2) );
      }
    }
    return true;
  }

  class MultiGpuExchange
  {
  public:
    // allgather( const void* mine , void* all , size_t bytes_per_rank ) -> bool : collective over the `world` processes
    // transport: ONK_XCHG_TRANSPORT_DEVICE (device memory over NVLink, needs peer access) or ONK_XCHG_TRANSPORT_HOST (pinned host
    // memory in POSIX shared memory, works everywhere); -1 = the library's default (ONK_XCHG_TRANSPORT in the environment)
    template<class AllGatherF>
    inline MultiGpuExchange( int rank , int world , AllGatherF && allgather , int transport = -1 ) : m_rank(rank) , m_world(world)
    {
      if( transport < 0 ) ONIKA_B200_CHECK( onk_xchg_create( rank , world , & m_x ) );
      else ONIKA_B200_CHECK( onk_xchg_create_ex( rank , world , transport , & m_x ) );
      ONIKA_B200_CHECK( onk_alloc( ONK_XCHG_MAX_VALUES * sizeof(double) , 64 , ONK_MEM_PINNED , (void**) & m_values ) );
      if( world > 1 )
      {
        unsigned char mine[ONK_IPC_HANDLE_BYTES];
        std::vector<unsigned char> all( size_t(world) * ONK_IPC_HANDLE_BYTES );
        ONIKA_B200_CHECK( onk_xchg_export( m_x , mine ) );
        if( ! allgather( (const void*) mine , (void*) all.data() , size_t(ONK_IPC_HANDLE_BYTES) ) ) fatal_error() << "MultiGpuExchange: handle all-gather failed" << std::endl;
        ONIKA_B200_CHECK( onk_xchg_connect( m_x , all.data() ) );
      }
    }
    // file rendezvous in `dir` (unique `tag` per exchange)
    inline MultiGpuExchange( int rank , int world , const std::string& dir , const std::string& tag , int transport = -1 )
      : MultiGpuExchange( rank , world , [&](const void* mine, void* all, size_t n) { return file_allgather( dir , tag , rank , world , mine , all , n ); } , transport ) {}
    inline ~MultiGpuExchange() { if( m_values ) onk_free( m_values , 0 ); if( m_x ) onk_xchg_destroy( m_x ); }
    MultiGpuExchange( const MultiGpuExchange& ) = delete;
    MultiGpuExchange& operator = ( const MultiGpuExchange& ) = delete;

    inline int rank() const { return m_rank; }
    inline int world() const { return m_world; }
    inline onk_xchg* handle() { return m_x; }

    // in-place all-reduce of `count` <= 8 doubles in device-accessible memory, asynchronous on `lane`
    inline void all_reduce_device( MultiGpuOp op , double* values_dev , int count , int lane = ONK_LANE_DEFAULT ) { ONIKA_B200_CHECK( onk_xchg_allreduce_f64( m_x , int(op) , values_dev , count , lane ) ); }
    // local reduction of in_dev[0:n] fused with the cross-GPU combine: ONE kernel per rank, result (same bits everywhere) in out_dev
    inline void reduce_device( MultiGpuOp op , const double* in_dev , uint64_t n , double* out_dev , int lane = ONK_LANE_DEFAULT ) { ONIKA_B200_CHECK( onk_reduce_xchg_f64( int(op) , in_dev , n , out_dev , m_x , lane ) ); }
    inline void barrier( int lane = ONK_LANE_DEFAULT ) { ONIKA_B200_CHECK( onk_xchg_barrier( m_x , lane ) ); }
    // healthy() is false once a peer failed to show up in a call (the calls after that are refused by the library)
    inline bool healthy() const { uint64_t f = 0; ONIKA_B200_CHECK( onk_xchg_status( m_x , &f ) ); return f == 0; }

    // host scalars, synchronous -- the reference's call shape
    template<class T, class... U>
    inline void all_reduce_multi( MultiGpuOp op , T , U& ... u )
    {
      static_assert( sizeof...(U) >= 1 && sizeof...(U) <= ONK_XCHG_MAX_VALUES , "1..8 values per call" );
      static_assert( std::is_arithmetic_v<T> && ( ... && std::is_arithmetic_v<U> ) , "arithmetic scalars only" );
      static_assert( std::is_floating_point_v<T> || sizeof(T) <= 4 , "values travel as fp64: integers wider than 32 bits would lose bits" );
      int k = 0;
      ( ... , ( m_values[k++] = double( T(u) ) ) );
      all_reduce_device( op , m_values , int(sizeof...(U)) , ONK_LANE_DEFAULT );
      ONIKA_B200_CHECK( onk_lane_sync( ONK_LANE_DEFAULT ) );
      k = 0;
      ( ... , ( u = U( T( m_values[k++] ) ) ) );
    }

  private:
    onk_xchg* m_x = nullptr;
    double* m_values = nullptr;     // pinned + mapped: written by the host, read and written in place by the kernel
    int m_rank = 0 , m_world = 1;
  };

  // free-function spelling of the reference's all_reduce_multi( comm , op , T{} , values... )
  template<class T, class... U>
  inline void all_reduce_multi( MultiGpuExchange& x , MultiGpuOp op , T t , U& ... u ) { x.all_reduce_multi( op , t , u ... ); }

} }
