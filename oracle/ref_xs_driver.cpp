// oracle/ref_xs_driver.cpp -- TEST INFRASTRUCTURE: runs the UNMODIFIED reference id-based redistribution
// (src/mpi/xs_data_move.cpp:communication_scheme_from_ids + include/onika/mpi/xs_data_move_impl.h:data_move) for `world`
// in-process ranks (oracle/stub/mpi.h) and returns every rank's index tables and moved payloads.
// Inputs/outputs are rank-concatenated arrays; before_off/after_off have world+1 entries (prefix sums of the counts).
#include <onika/mpi/xs_data_move.h>
#include <onika/test/unit_test.h>
#include <cstdint>
#include <cstring>
#include <vector>

namespace
{
  struct Elem32 { unsigned char b[32]; };
  struct Job
  {
    int world; uint64_t id_min, id_max;
    const uint64_t *before_off, *ids_before, *after_off, *ids_after;
    int32_t *send_indices, *recv_indices, *send_count, *send_displ, *recv_count, *recv_displ;   // counts/displs: world x world
    const double* in8; double* out8; const unsigned char* in32; unsigned char* out32;
  };

  void rank_main(int rank, void* user)
  {
    using namespace onika::mpi;
    Job& j = *static_cast<Job*>(user);
    const uint64_t b0 = j.before_off[rank], nb = j.before_off[rank+1] - b0;
    const uint64_t a0 = j.after_off[rank],  na = j.after_off[rank+1] - a0;
    std::vector<int> si, sc, sd, ri, rc, rd;
    communication_scheme_from_ids( MPI_COMM_WORLD, j.id_min, j.id_max, nb, j.ids_before + b0, na, j.ids_after + a0, si, sc, sd, ri, rc, rd );
    for(uint64_t i=0;i<nb;i++) j.send_indices[b0+i] = si[i];
    for(uint64_t i=0;i<na;i++) j.recv_indices[a0+i] = ri[i];
    for(int p=0;p<j.world;p++)
    {
      j.send_count[rank*j.world+p] = sc[p]; j.send_displ[rank*j.world+p] = sd[p];
      j.recv_count[rank*j.world+p] = rc[p]; j.recv_displ[rank*j.world+p] = rd[p];
    }
    if( j.in8 != nullptr ) data_move( MPI_COMM_WORLD, si, sc, sd, ri, rc, rd, j.in8 + b0, j.out8 + a0 );
    if( j.in32 != nullptr )
      data_move( MPI_COMM_WORLD, si, sc, sd, ri, rc, rd, reinterpret_cast<const Elem32*>(j.in32) + b0, reinterpret_cast<Elem32*>(j.out32) + a0 );
  }
}

extern "C" void ref_xs_data_move(int world, uint64_t id_min, uint64_t id_max,
                                 const uint64_t* before_off, const uint64_t* ids_before, const uint64_t* after_off, const uint64_t* ids_after,
                                 int32_t* send_indices, int32_t* recv_indices, int32_t* send_count, int32_t* send_displ, int32_t* recv_count, int32_t* recv_displ,
                                 const double* in8, double* out8, const unsigned char* in32, unsigned char* out32)
{
  Job j{ world, id_min, id_max, before_off, ids_before, after_off, ids_after, send_indices, recv_indices, send_count, send_displ, recv_count, recv_displ, in8, out8, in32, out32 };
  ref_mpi_run( world, rank_main, &j );
}

// The range utilities register the reference's own unit tests (xs_data_move_range_utils.h:96-129) at load time through
// onika::UnitTest / plugin_db_register, whose definitions live in src/core/unit_test.cpp + src/core/plugin.cpp (dlopen,
// plugin database: not needed here).  These minimal definitions keep the registered tests so they can be run below.
namespace onika
{
  UnitTest* UnitTest::s_unit_tests_list = nullptr;
  void UnitTest::register_unit_test(const std::string& name , std::function<void()> F )
  {
    auto* t = new UnitTest{ name , F , s_unit_tests_list };
    s_unit_tests_list = t;
  }
  void plugin_db_register( const std::string& , const std::string& ) {}
}

// runs every registered reference unit test; returns the number of failures and writes the number run
extern "C" int ref_run_unit_tests(int* n_run)
{
  int failed = 0, run = 0;
  for(auto* t = onika::UnitTest::s_unit_tests_list; t != nullptr; t = t->m_next)
  {
    ++run;
    try { t->m_test(); } catch( ... ) { ++failed; }
  }
  if( n_run != nullptr ) *n_run = run;
  return failed;
}
