// TEST INFRASTRUCTURE ONLY (oracle build).  Minimal MPI stand-in that lets the
// reference's hot-path sources compile and run UNCHANGED without an MPI
// installation.  Ranks are fork()ed processes connected by socketpairs
// (MINIMPI_NP=N in the environment; default 1).  Only the 16 calls the path
// uses exist (reference call sites: src/linearpart.h:195-384,431-467,
// src/tiffIO.cpp:383-426, src/d8.cpp:293,334-346, src/aread8.cpp:118-133).
#pragma once
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef long long MPI_Offset;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR, count_bytes; } MPI_Status;
#define MPI_COMM_WORLD 0
#define MPI_SUM 0
#define MPI_BSEND_OVERHEAD 96
#define MPI_BYTE 1
#define MPI_SHORT 2
#define MPI_INT 3
#define MPI_LONG 4
#define MPI_FLOAT 5
#define MPI_DOUBLE 6
#define MPI_INT16_T 7
#define MPI_INT32_T 8
#define MPI_SUCCESS 0
int MPI_Init(int*, char***);
int MPI_Finalize(void);
int MPI_Comm_rank(MPI_Comm, int*);
int MPI_Comm_size(MPI_Comm, int*);
double MPI_Wtime(void);
int MPI_Abort(MPI_Comm, int);
int MPI_Barrier(MPI_Comm);
int MPI_Allreduce(const void*, void*, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Bcast(void*, int, MPI_Datatype, int, MPI_Comm);
int MPI_Send(const void*, int, MPI_Datatype, int, int, MPI_Comm);
int MPI_Bsend(const void*, int, MPI_Datatype, int, int, MPI_Comm);
int MPI_Recv(void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status*);
int MPI_Probe(int, int, MPI_Comm, MPI_Status*);
int MPI_Get_count(const MPI_Status*, MPI_Datatype, int*);
int MPI_Buffer_attach(void*, int);
int MPI_Buffer_detach(void*, int*);
#ifdef __cplusplus
}
#endif
