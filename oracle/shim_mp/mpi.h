/* Multi-PROCESS MPI stand-in used ONLY to run the reference oracle on several ranks inside one container
 * (oracle/_ref/libdktref_mp_*.so).  TEST INFRASTRUCTURE - not part of the product path.
 *
 * The reference (paralab/Dendro-KT) is an MPI program; this image has no MPI.  oracle/shim/mpi.h serves the single-rank
 * oracle.  This header declares the subset of MPI the reference calls on its distributed path (inventory: DESIGN.md,
 * SURVEY.md 8f N3) and oracle/shim_mp/mpi_mp.cpp implements it for ranks that are fork()ed children of one parent,
 * talking through an anonymous shared-memory arena: eager buffered point-to-point messages with MPI's matching rules
 * (source, tag, communicator; non-overtaking), collectives built on them in rank order (deterministic reductions).
 * Datatypes carry their size and kind; user-defined contiguous types and user reduction functions are supported.
 */
#ifndef DKT_ORACLE_MPI_MP_H
#define DKT_ORACLE_MPI_MP_H

#include <stddef.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef int MPI_Group;
typedef long MPI_Aint;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; long dkt_bytes; } MPI_Status;
typedef void (MPI_User_function)(void *, void *, int *, MPI_Datatype *);

#define MPI_SUCCESS 0
#define MPI_COMM_NULL 0
#define MPI_COMM_WORLD 1
#define MPI_COMM_SELF 2
#define MPI_GROUP_NULL 0
#define MPI_UNDEFINED (-32766)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
#define MPI_REQUEST_NULL 0
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_IN_PLACE ((void *)1)

/* datatype = kind << 24 | bytes;  kind: 0 opaque bytes, 1 signed integer, 2 unsigned integer, 3 floating point */
#define DKT_MPI_T(kind, bytes) (((kind) << 24) | (bytes))
#define MPI_BYTE DKT_MPI_T(0, 1)
#define MPI_CHAR DKT_MPI_T(1, 1)
#define MPI_UNSIGNED_CHAR DKT_MPI_T(2, 1)
#define MPI_SHORT DKT_MPI_T(1, 2)
#define MPI_UNSIGNED_SHORT DKT_MPI_T(2, 2)
#define MPI_INT DKT_MPI_T(1, 4)
#define MPI_UNSIGNED DKT_MPI_T(2, 4)
#define MPI_FLOAT DKT_MPI_T(3, 4)
#define MPI_LONG DKT_MPI_T(1, 8)
#define MPI_UNSIGNED_LONG DKT_MPI_T(2, 8)
#define MPI_LONG_LONG_INT DKT_MPI_T(1, 8)
#define MPI_LONG_LONG DKT_MPI_T(1, 8)
#define MPI_UNSIGNED_LONG_LONG DKT_MPI_T(2, 8)
#define MPI_DOUBLE DKT_MPI_T(3, 8)
#define MPI_LONG_DOUBLE DKT_MPI_T(3, 16)

#define MPI_SUM 1
#define MPI_MIN 2
#define MPI_MAX 3
#define MPI_LAND 4
#define MPI_LOR 5
#define MPI_PROD 6

#ifdef __cplusplus
extern "C" {
#endif
int MPI_Init(int *, char ***);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm, int);
int MPI_Comm_rank(MPI_Comm, int *);
int MPI_Comm_size(MPI_Comm, int *);
int MPI_Barrier(MPI_Comm);
double MPI_Wtime(void);
int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm);
int MPI_Reduce(const void *, void *, int, MPI_Datatype, MPI_Op, int, MPI_Comm);
int MPI_Allreduce(const void *, void *, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Scan(const void *, void *, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Gather(const void *, int, MPI_Datatype, void *, int, MPI_Datatype, int, MPI_Comm);
int MPI_Allgather(const void *, int, MPI_Datatype, void *, int, MPI_Datatype, MPI_Comm);
int MPI_Allgatherv(const void *, int, MPI_Datatype, void *, const int *, const int *, MPI_Datatype, MPI_Comm);
int MPI_Alltoall(const void *, int, MPI_Datatype, void *, int, MPI_Datatype, MPI_Comm);
int MPI_Alltoallv(const void *, const int *, const int *, MPI_Datatype, void *, const int *, const int *, MPI_Datatype, MPI_Comm);
int MPI_Isend(const void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *);
int MPI_Issend(const void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *);
int MPI_Send(const void *, int, MPI_Datatype, int, int, MPI_Comm);
int MPI_Irecv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *);
int MPI_Recv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *);
int MPI_Sendrecv(const void *, int, MPI_Datatype, int, int, void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *);
int MPI_Wait(MPI_Request *, MPI_Status *);
int MPI_Waitall(int, MPI_Request *, MPI_Status *);
int MPI_Get_count(const MPI_Status *, MPI_Datatype, int *);
int MPI_Type_contiguous(int, MPI_Datatype, MPI_Datatype *);
int MPI_Type_commit(MPI_Datatype *);
int MPI_Type_free(MPI_Datatype *);
int MPI_Op_create(MPI_User_function *, int, MPI_Op *);
int MPI_Op_free(MPI_Op *);
int MPI_Comm_split(MPI_Comm, int, int, MPI_Comm *);
int MPI_Comm_dup(MPI_Comm, MPI_Comm *);
int MPI_Comm_free(MPI_Comm *);
int MPI_Comm_group(MPI_Comm, MPI_Group *);
int MPI_Group_incl(MPI_Group, int, const int *, MPI_Group *);
int MPI_Group_free(MPI_Group *);
int MPI_Comm_create(MPI_Comm, MPI_Group, MPI_Comm *);

/* control: the parent creates the world BEFORE fork(); every child then names its rank */
int dktmp_world_create(int nranks, size_t arena_bytes);
int dktmp_set_rank(int rank);
long dktmp_bytes_sent(void);
#ifdef __cplusplus
}
#endif
#endif
