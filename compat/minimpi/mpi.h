/*
 * minimpi -- a tiny single-node MPI stand-in (NOT an MPI implementation).
 *
 * Why it exists: neither the development container nor the B200 boxes have an
 * MPI installation, yet the drop-in boundary of this project
 * (include/contour3d.h -> sdecomp.h) includes <mpi.h>, and the reference
 * program (NaokiHori/Contour3D, src/main.c) is an MPI program.  This header +
 * minimpi.c provide exactly the subset of MPI those two users need, with
 * ranks living in processes of ONE node that share a memory arena:
 *
 *   - fork mode   : MINIMPI_NP=<n> ./prog      MPI_Init forks n-1 children
 *   - attach mode : RANK / WORLD_SIZE / MASTER_PORT (torchrun) or
 *                   MINIMPI_RANK / MINIMPI_SIZE / MINIMPI_SESSION
 *                   -> independently launched processes attach to a named
 *                   shared-memory segment
 *   - default     : one rank
 *
 * Every communication call is COLLECTIVE in this stand-in: all ranks must
 * issue the same sequence of Sendrecv / Reduce / Allreduce / Bcast / Barrier
 * calls (true for halo exchanges on Cartesian shifts and for the image
 * reduction, the only patterns used here).
 *
 * On a machine with a real MPI, do not put this directory on the include
 * path; the project then compiles against the real <mpi.h> unchanged.
 */
#if !defined(MINIMPI_MPI_H)
#define MINIMPI_MPI_H

#include <stddef.h>

#if defined(__cplusplus)
extern "C" {
#endif

#define MINIMPI 1

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef long MPI_Aint;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
typedef void MPI_User_function(void *invec, void *inoutvec, int *len, MPI_Datatype *datatype);

#define MPI_SUCCESS        0
#define MPI_ERR_OTHER      15

#define MPI_COMM_NULL      0
#define MPI_COMM_WORLD     1
#define MPI_COMM_SELF      2

#define MPI_DATATYPE_NULL  0
#define MPI_BYTE           1
#define MPI_CHAR           2
#define MPI_UNSIGNED_CHAR  3
#define MPI_INT            4
#define MPI_FLOAT          5
#define MPI_DOUBLE         6
#define MPI_UNSIGNED_LONG  7
#define MPI_UINT64_T       8

#define MPI_OP_NULL        0
#define MPI_MIN            1
#define MPI_MAX            2
#define MPI_SUM            3

#define MPI_PROC_NULL      (-1)
#define MPI_ANY_TAG        (-1)
#define MPI_STATUS_IGNORE  ((MPI_Status *)0)
#define MPI_IN_PLACE       ((void *)-1)

int MPI_Init(int *argc, char ***argv);
int MPI_Initialized(int *flag);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int errorcode);
double MPI_Wtime(void);

int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_dup(MPI_Comm comm, MPI_Comm *newcomm);
int MPI_Comm_free(MPI_Comm *comm);

int MPI_Dims_create(int nnodes, int ndims, int dims[]);
int MPI_Cart_create(MPI_Comm comm_old, int ndims, const int dims[], const int periods[], int reorder, MPI_Comm *comm_cart);
int MPI_Cart_get(MPI_Comm comm, int maxdims, int dims[], int periods[], int coords[]);
int MPI_Cart_shift(MPI_Comm comm, int direction, int disp, int *rank_source, int *rank_dest);
int MPI_Cart_sub(MPI_Comm comm, const int remain_dims[], MPI_Comm *newcomm);

int MPI_Type_create_hvector(int count, int blocklength, MPI_Aint stride, MPI_Datatype oldtype, MPI_Datatype *newtype);
int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype *newtype);
int MPI_Type_create_struct(int count, const int blocklengths[], const MPI_Aint displacements[], const MPI_Datatype types[], MPI_Datatype *newtype);
int MPI_Type_commit(MPI_Datatype *datatype);
int MPI_Type_free(MPI_Datatype *datatype);

int MPI_Op_create(MPI_User_function *fn, int commute, MPI_Op *op);
int MPI_Op_free(MPI_Op *op);

int MPI_Barrier(MPI_Comm comm);
int MPI_Bcast(void *buffer, int count, MPI_Datatype datatype, int root, MPI_Comm comm);
int MPI_Sendrecv(const void *sendbuf, int sendcount, MPI_Datatype sendtype, int dest, int sendtag,
                 void *recvbuf, int recvcount, MPI_Datatype recvtype, int source, int recvtag,
                 MPI_Comm comm, MPI_Status *status);
int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype datatype, MPI_Op op, int root, MPI_Comm comm);
int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype datatype, MPI_Op op, MPI_Comm comm);
int MPI_Alltoallw(const void *sendbuf, const int sendcounts[], const int sdispls[], const MPI_Datatype sendtypes[],
                  void *recvbuf, const int recvcounts[], const int rdispls[], const MPI_Datatype recvtypes[], MPI_Comm comm);

#if defined(__cplusplus)
}
#endif

#endif /* MINIMPI_MPI_H */
