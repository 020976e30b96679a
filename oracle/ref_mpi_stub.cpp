// oracle/ref_mpi_stub.cpp -- TEST INFRASTRUCTURE: thread-per-rank implementation of oracle/stub/mpi.h.
#include <mpi.h>
#include <barrier>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

namespace
{
  struct Posted { const void* sendbuf; const int* counts; const int* displs; int count1; int elem; };
  struct World
  {
    int size = 1;
    std::unique_ptr<std::barrier<>> bar;
    std::vector<Posted> posted;
  };
  World* g_world = nullptr;
  thread_local int t_rank = 0;
}

extern "C" {

int MPI_Comm_rank(MPI_Comm, int* rank) { *rank = t_rank; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm, int* size) { *size = g_world ? g_world->size : 1; return MPI_SUCCESS; }

int MPI_Alltoall(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int, MPI_Datatype, MPI_Comm)
{
  World& w = *g_world;
  w.posted[t_rank] = Posted{ sendbuf, nullptr, nullptr, sendcount, sendtype };
  w.bar->arrive_and_wait();
  const size_t chunk = size_t(sendcount) * size_t(sendtype);
  for(int p=0;p<w.size;p++)
    std::memcpy( (char*)recvbuf + size_t(p)*chunk, (const char*)w.posted[p].sendbuf + size_t(t_rank)*chunk, chunk );
  w.bar->arrive_and_wait();
  return MPI_SUCCESS;
}

int MPI_Alltoallv(const void* sendbuf, const int* sendcounts, const int* sdispls, MPI_Datatype sendtype,
                  void* recvbuf, const int* recvcounts, const int* rdispls, MPI_Datatype, MPI_Comm)
{
  World& w = *g_world;
  w.posted[t_rank] = Posted{ sendbuf, sendcounts, sdispls, 0, sendtype };
  w.bar->arrive_and_wait();
  for(int p=0;p<w.size;p++)
  {
    const Posted& s = w.posted[p];
    if( s.counts[t_rank] != recvcounts[p] ) { std::fprintf(stderr,"mpi stub: alltoallv count mismatch %d -> %d\n",p,t_rank); std::abort(); }
    std::memcpy( (char*)recvbuf + size_t(rdispls[p])*size_t(sendtype),
                 (const char*)s.sendbuf + size_t(s.displs[t_rank])*size_t(s.elem), size_t(recvcounts[p])*size_t(sendtype) );
  }
  w.bar->arrive_and_wait();
  return MPI_SUCCESS;
}

int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype* newtype) { *newtype = count * oldtype; return MPI_SUCCESS; }
int MPI_Type_commit(MPI_Datatype*) { return MPI_SUCCESS; }
int MPI_Type_free(MPI_Datatype*) { return MPI_SUCCESS; }
int MPI_Abort(MPI_Comm, int code) { std::fprintf(stderr,"mpi stub: MPI_Abort(%d)\n",code); std::abort(); }

void ref_mpi_run(int world, void (*fn)(int, void*), void* user)
{
  World w; w.size = world; w.bar = std::make_unique<std::barrier<>>( world ); w.posted.resize( world );
  g_world = &w;
  std::vector<std::thread> th;
  for(int r=0;r<world;r++) th.emplace_back( [=]{ t_rank = r; fn(r,user); } );
  for(auto& t : th) t.join();
  g_world = nullptr;
}

}
