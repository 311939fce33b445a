// TEST INFRASTRUCTURE — in-process stand-in for <mpi.h> (no MPI in this image).
//
// Ranks are THREADS of one process: oracle/ref_harness.cc starts one thread per
// subdomain and calls minifem_ref_mpi_bind() on it.  Only the calls the reference's
// bulk-synchronous XMPI path makes exist (halo.cc:58,91,96; FEM.cc:113;
// main.cc:101-103,381).  Sends are buffered, so the Irecv / Send / Waitall sequence
// of halo.cc:52-96 cannot deadlock.
#ifndef MINIFEM_ORACLE_MPI_SHIM_H
#define MINIFEM_ORACLE_MPI_SHIM_H

#include <cstdint>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef struct { void *buf; int count, source, tag, bytes; } MPI_Request;
typedef struct { int unused; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_DOUBLE     8
#define MPI_UINT64_T   9
#define MPI_MAX        1
#define MPI_SUCCESS    0
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)

int MPI_Init (int *argc, char ***argv);
int MPI_Finalize ();
int MPI_Comm_size (MPI_Comm comm, int *size);
int MPI_Comm_rank (MPI_Comm comm, int *rank);
int MPI_Irecv (void *buf, int count, MPI_Datatype type, int source, int tag,
               MPI_Comm comm, MPI_Request *req);
int MPI_Send (const void *buf, int count, MPI_Datatype type, int dest, int tag,
              MPI_Comm comm);
int MPI_Waitall (int count, MPI_Request *reqs, MPI_Status *statuses);
int MPI_Reduce (const void *sendbuf, void *recvbuf, int count, MPI_Datatype type,
                MPI_Op op, int root, MPI_Comm comm);

// Harness side: set the world size once, then bind each rank thread.
void minifem_ref_mpi_world (int nranks);
void minifem_ref_mpi_bind (int rank);

#endif
