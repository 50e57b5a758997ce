/* mpi.h -- SHIM, test infrastructure only: a one-process stand-in for the handful of MPI calls the reference makes
 * (src/MGSolver.cpp:51-52, 694; src/EllipticSolver.cpp; src/Impec2DSolver.cpp; src/Driver.cpp).  Rank 0 of 1. */
#ifndef IMPEC_REF_SHIM_MPI_H
#define IMPEC_REF_SHIM_MPI_H
#include <string.h>
#include <stdlib.h>
#include <stdio.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef struct MPI_Status { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_DOUBLE 8
#define MPI_INT 4
#define MPI_CHAR 1
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
static inline int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; return MPI_SUCCESS; }
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm c, int *r) { (void)c; *r = 0; return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm c, int *n) { (void)c; *n = 1; return MPI_SUCCESS; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return MPI_SUCCESS; }
static inline int MPI_Abort(MPI_Comm c, int code) { (void)c; exit(code); return MPI_SUCCESS; }
static inline double MPI_Wtime(void) { return 0.0; }
static inline int MPI_Allgather(const void *sb, int sc, MPI_Datatype st, void *rb, int rc, MPI_Datatype rt, MPI_Comm c) {
    (void)rc; (void)rt; (void)c; memcpy(rb, sb, (size_t)sc * (size_t)st); return MPI_SUCCESS;
}
static inline int MPI_Bcast(void *b, int n, MPI_Datatype t, int root, MPI_Comm c) { (void)b; (void)n; (void)t; (void)root; (void)c; return MPI_SUCCESS; }
static inline int MPI_Send(const void *b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c) {
    (void)b; (void)n; (void)t; (void)dst; (void)tag; (void)c; fprintf(stderr, "mpi shim: MPI_Send in a one-process run\n"); abort(); return 1;
}
static inline int MPI_Recv(void *b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *s) {
    (void)b; (void)n; (void)t; (void)src; (void)tag; (void)c; (void)s; fprintf(stderr, "mpi shim: MPI_Recv in a one-process run\n"); abort(); return 1;
}
#ifdef __cplusplus
}
#endif
#endif
