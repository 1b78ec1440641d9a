/* Single-rank MPI stand-in used ONLY to build the reference oracle (oracle/_ref).
 * TEST INFRASTRUCTURE - not part of the product path.
 *
 * The reference (paralab/Dendro-KT) is an MPI program; this image has no MPI.
 * Every nProc==1 branch of the reference avoids point-to-point traffic
 * (include/oda.tcc:216,283,324,387; include/nsort.tcc:519; src/tsort.cpp:254), so a
 * communicator of size one with copy-collectives is sufficient.  Datatypes are encoded
 * as their size in bytes.  Point-to-point calls abort: reaching one means a multi-rank
 * branch was taken, which this shim cannot honour.
 */
#ifndef DKT_ORACLE_MPI_SHIM_H
#define DKT_ORACLE_MPI_SHIM_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef int MPI_Group;
typedef long MPI_Aint;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
typedef void (MPI_User_function)(void *, void *, int *, MPI_Datatype *);

#define MPI_SUCCESS 0
#define MPI_COMM_WORLD 1
#define MPI_COMM_NULL 0
#define MPI_COMM_SELF 2
#define MPI_UNDEFINED (-32766)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
#define MPI_REQUEST_NULL 0
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_IN_PLACE ((void *)1)

#define MPI_BYTE 1
#define MPI_CHAR 1
#define MPI_UNSIGNED_CHAR 1
#define MPI_SHORT 2
#define MPI_UNSIGNED_SHORT 2
#define MPI_INT 4
#define MPI_UNSIGNED 4
#define MPI_FLOAT 4
#define MPI_LONG 8
#define MPI_UNSIGNED_LONG 8
#define MPI_LONG_LONG_INT 8
#define MPI_LONG_LONG 8
#define MPI_UNSIGNED_LONG_LONG 8
#define MPI_DOUBLE 8
#define MPI_LONG_DOUBLE 16

#define MPI_SUM 1
#define MPI_MIN 2
#define MPI_MAX 3
#define MPI_LAND 4
#define MPI_LOR 5
#define MPI_PROD 6

static inline void dkt_shim_p2p_(const char *what)
{
  fprintf(stderr, "[oracle mpi shim] %s reached: a multi-rank branch ran on the single-rank shim\n", what);
  abort();
}
static inline void dkt_shim_copy_(const void *s, void *r, long bytes)
{
  if (s != MPI_IN_PLACE && s != r && bytes > 0) memmove(r, s, (size_t)bytes);
}

static inline int MPI_Init(int *, char ***) { return MPI_SUCCESS; }
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
static inline int MPI_Abort(MPI_Comm, int code) { exit(code); return 0; }
static inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm, int *s) { *s = 1; return MPI_SUCCESS; }
static inline int MPI_Barrier(MPI_Comm) { return MPI_SUCCESS; }
static inline double MPI_Wtime(void)
{
  struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return MPI_SUCCESS; }
static inline int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op, int, MPI_Comm)
{ dkt_shim_copy_(s, r, (long)n * t); return MPI_SUCCESS; }
static inline int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op, MPI_Comm)
{ dkt_shim_copy_(s, r, (long)n * t); return MPI_SUCCESS; }
static inline int MPI_Scan(const void *s, void *r, int n, MPI_Datatype t, MPI_Op, MPI_Comm)
{ dkt_shim_copy_(s, r, (long)n * t); return MPI_SUCCESS; }
static inline int MPI_Gather(const void *s, int n, MPI_Datatype t, void *r, int, MPI_Datatype, int, MPI_Comm)
{ dkt_shim_copy_(s, r, (long)n * t); return MPI_SUCCESS; }
static inline int MPI_Allgather(const void *s, int n, MPI_Datatype t, void *r, int, MPI_Datatype, MPI_Comm)
{ dkt_shim_copy_(s, r, (long)n * t); return MPI_SUCCESS; }
static inline int MPI_Alltoall(const void *s, int n, MPI_Datatype t, void *r, int, MPI_Datatype, MPI_Comm)
{ dkt_shim_copy_(s, r, (long)n * t); return MPI_SUCCESS; }
static inline int MPI_Allgatherv(const void *s, int n, MPI_Datatype t, void *r, const int *, const int *displs,
                                 MPI_Datatype rt, MPI_Comm)
{ dkt_shim_copy_(s, (char *)r + (long)displs[0] * rt, (long)n * t); return MPI_SUCCESS; }
static inline int MPI_Alltoallv(const void *s, const int *sc, const int *sd, MPI_Datatype t, void *r, const int *,
                                const int *rd, MPI_Datatype rt, MPI_Comm)
{ dkt_shim_copy_((const char *)s + (long)sd[0] * t, (char *)r + (long)rd[0] * rt, (long)sc[0] * t); return MPI_SUCCESS; }

static inline int MPI_Isend(const void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *) { dkt_shim_p2p_("MPI_Isend"); return 1; }
static inline int MPI_Issend(const void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *) { dkt_shim_p2p_("MPI_Issend"); return 1; }
static inline int MPI_Send(const void *, int, MPI_Datatype, int, int, MPI_Comm) { dkt_shim_p2p_("MPI_Send"); return 1; }
static inline int MPI_Irecv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *) { dkt_shim_p2p_("MPI_Irecv"); return 1; }
static inline int MPI_Recv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *) { dkt_shim_p2p_("MPI_Recv"); return 1; }
static inline int MPI_Sendrecv(const void *, int, MPI_Datatype, int, int, void *, int, MPI_Datatype, int, int, MPI_Comm,
                               MPI_Status *) { dkt_shim_p2p_("MPI_Sendrecv"); return 1; }
static inline int MPI_Wait(MPI_Request *, MPI_Status *) { return MPI_SUCCESS; }
static inline int MPI_Waitall(int, MPI_Request *, MPI_Status *) { return MPI_SUCCESS; }
static inline int MPI_Get_count(const MPI_Status *, MPI_Datatype, int *c) { *c = 0; return MPI_SUCCESS; }

static inline int MPI_Type_contiguous(int n, MPI_Datatype old, MPI_Datatype *nt) { *nt = n * old; return MPI_SUCCESS; }
static inline int MPI_Type_commit(MPI_Datatype *) { return MPI_SUCCESS; }
static inline int MPI_Type_free(MPI_Datatype *) { return MPI_SUCCESS; }
static inline int MPI_Op_create(MPI_User_function *, int, MPI_Op *op) { *op = 100; return MPI_SUCCESS; }
static inline int MPI_Op_free(MPI_Op *) { return MPI_SUCCESS; }
static inline int MPI_Comm_split(MPI_Comm c, int, int, MPI_Comm *n) { *n = c; return MPI_SUCCESS; }
static inline int MPI_Comm_dup(MPI_Comm c, MPI_Comm *n) { *n = c; return MPI_SUCCESS; }
static inline int MPI_Comm_free(MPI_Comm *) { return MPI_SUCCESS; }
static inline int MPI_Comm_group(MPI_Comm, MPI_Group *g) { *g = 1; return MPI_SUCCESS; }
static inline int MPI_Group_incl(MPI_Group, int, const int *, MPI_Group *g) { *g = 1; return MPI_SUCCESS; }
static inline int MPI_Group_free(MPI_Group *) { return MPI_SUCCESS; }
static inline int MPI_Comm_create(MPI_Comm c, MPI_Group, MPI_Comm *n) { *n = c; return MPI_SUCCESS; }

#endif
