/* minimpi — a tiny single-node stand-in for <mpi.h>.
 *
 * TEST/BASELINE INFRASTRUCTURE ONLY.  It exists so that the UNMODIFIED reference
 * sources under /root/reference (which include <mpi.h> and call 11 MPI functions,
 * see SURVEY.md §2.3) can be compiled in an image that ships no MPI at all.
 * Nothing in the product path (matrix_solver_b200/) includes this header.
 *
 * Ranks are processes created by fork() inside MPI_Init; the rank count comes from
 * the environment variable MINIMPI_NP (default 1).  Collectives go through one
 * anonymous shared mapping created before the fork.
 *
 * Only what the reference uses is provided:
 *   MPI_Init, MPI_Finalize, MPI_Comm_rank, MPI_Comm_size, MPI_Bcast, MPI_Send,
 *   MPI_Recv, MPI_Gatherv, MPI_Allreduce, MPI_Reduce  (+ MPI_Barrier for drivers)
 *   datatypes MPI_DOUBLE, MPI_INT, MPI_UNSIGNED_LONG; op MPI_SUM; MPI_COMM_WORLD.
 */
#ifndef MINIMPI_MPI_H
#define MINIMPI_MPI_H

/* The reference's init.cpp relies on the real mpi.h dragging these in. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0

#define MPI_DOUBLE 1
#define MPI_INT 2
#define MPI_UNSIGNED_LONG 3

#define MPI_SUM 1

#define MPI_STATUS_IGNORE ((MPI_Status *)0)

int MPI_Init (int *argc, char ***argv);
int MPI_Finalize (void);
int MPI_Comm_rank (MPI_Comm comm, int *rank);
int MPI_Comm_size (MPI_Comm comm, int *size);
int MPI_Barrier (MPI_Comm comm);
int MPI_Bcast (void *buf, int count, MPI_Datatype type, int root, MPI_Comm comm);
int MPI_Send (const void *buf, int count, MPI_Datatype type, int dest, int tag,
              MPI_Comm comm);
int MPI_Recv (void *buf, int count, MPI_Datatype type, int source, int tag,
              MPI_Comm comm, MPI_Status *status);
int MPI_Gatherv (const void *sendbuf, int sendcount, MPI_Datatype sendtype,
                 void *recvbuf, const int *recvcounts, const int *displs,
                 MPI_Datatype recvtype, int root, MPI_Comm comm);
int MPI_Allreduce (const void *sendbuf, void *recvbuf, int count,
                   MPI_Datatype type, MPI_Op op, MPI_Comm comm);
int MPI_Reduce (const void *sendbuf, void *recvbuf, int count, MPI_Datatype type,
                MPI_Op op, int root, MPI_Comm comm);

#ifdef __cplusplus
}
#endif
#endif
