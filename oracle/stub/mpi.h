/* oracle/stub/mpi.h -- TEST INFRASTRUCTURE.  In-process stand-in for the handful of MPI calls the reference's
 * src/mpi/xs_data_move.cpp and include/onika/mpi/xs_data_move_impl.h make, so that the UNMODIFIED reference sources can
 * be compiled and run here (no MPI in this image): every "rank" is a thread of one process, collectives are memory copies
 * between barriers (oracle/ref_mpi_stub.cpp).  Only what those two files use is declared. */
#pragma once
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;            /* = element size in bytes */
#define MPI_COMM_WORLD 0
#define MPI_CHAR 1
#define MPI_BYTE 1
#define MPI_INT  4
#define MPI_SUCCESS 0

int MPI_Comm_rank(MPI_Comm comm, int* rank);
int MPI_Comm_size(MPI_Comm comm, int* size);
int MPI_Alltoall(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount, MPI_Datatype recvtype, MPI_Comm comm);
int MPI_Alltoallv(const void* sendbuf, const int* sendcounts, const int* sdispls, MPI_Datatype sendtype,
                  void* recvbuf, const int* recvcounts, const int* rdispls, MPI_Datatype recvtype, MPI_Comm comm);
int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype* newtype);
int MPI_Type_commit(MPI_Datatype* t);
int MPI_Type_free(MPI_Datatype* t);
int MPI_Abort(MPI_Comm comm, int code);

/* driver side: run fn(rank, user) on `world` threads that form MPI_COMM_WORLD */
void ref_mpi_run(int world, void (*fn)(int rank, void* user), void* user);

#ifdef __cplusplus
}
#endif
