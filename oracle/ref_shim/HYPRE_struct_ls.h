/* HYPRE_struct_ls.h -- SHIM, test infrastructure only.
 *
 * HYPRE (hypre-2.8.0b, INSTALL:44-46 of the reference) is a third-party dependency that is neither in /root/reference nor in this
 * image.  oracle/Makefile.ref compiles the reference's UNMODIFIED src/MGSolver.cpp against this header.  The shim RECORDS what the
 * reference hands to HYPRE's struct interface -- grid extents, stencil offsets, every HYPRE_StructMatrixSetValues /
 * HYPRE_StructVectorSetValues call (src/MGSolver.cpp:153-185, 268-590, 997-1391) -- so that tests can compare the reference's
 * assembled 9-entry matrix and right-hand side entry by entry with the oracle's, and it SOLVES the recorded system with its own
 * Krylov solver (ref_hypre_shim.cpp: Dirichlet rows eliminated, PCG with a symmetric Gauss-Seidel preconditioner, BiCGStab when the
 * matrix is not symmetric) to the tolerance the reference sets (HYPRE_StructPCGSetTol, relative two-norm as :597-603 configure).
 * HYPRE's iterates, iteration counts and SMG cycle are NOT reproduced (iteration-level parity unpinned); the
 * converged solution of the reference's own matrix is.
 */
#ifndef IMPEC_REF_SHIM_HYPRE_STRUCT_LS_H
#define IMPEC_REF_SHIM_HYPRE_STRUCT_LS_H
#include "mpi.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef int HYPRE_Int;
struct refhypre_grid; struct refhypre_stencil; struct refhypre_matrix; struct refhypre_vector; struct refhypre_solver;
typedef struct refhypre_grid *HYPRE_StructGrid;
typedef struct refhypre_stencil *HYPRE_StructStencil;
typedef struct refhypre_matrix *HYPRE_StructMatrix;
typedef struct refhypre_vector *HYPRE_StructVector;
typedef struct refhypre_solver *HYPRE_StructSolver;
typedef HYPRE_Int (*HYPRE_PtrToStructSolverFcn)(HYPRE_StructSolver, HYPRE_StructMatrix, HYPRE_StructVector, HYPRE_StructVector);

HYPRE_Int HYPRE_StructGridCreate(MPI_Comm comm, HYPRE_Int ndim, HYPRE_StructGrid *grid);
HYPRE_Int HYPRE_StructGridSetExtents(HYPRE_StructGrid grid, HYPRE_Int *ilower, HYPRE_Int *iupper);
HYPRE_Int HYPRE_StructGridAssemble(HYPRE_StructGrid grid);
HYPRE_Int HYPRE_StructGridDestroy(HYPRE_StructGrid grid);
HYPRE_Int HYPRE_StructStencilCreate(HYPRE_Int ndim, HYPRE_Int size, HYPRE_StructStencil *stencil);
HYPRE_Int HYPRE_StructStencilSetElement(HYPRE_StructStencil stencil, HYPRE_Int entry, HYPRE_Int *offset);
HYPRE_Int HYPRE_StructStencilDestroy(HYPRE_StructStencil stencil);
HYPRE_Int HYPRE_StructMatrixCreate(MPI_Comm comm, HYPRE_StructGrid grid, HYPRE_StructStencil stencil, HYPRE_StructMatrix *matrix);
HYPRE_Int HYPRE_StructMatrixInitialize(HYPRE_StructMatrix matrix);
HYPRE_Int HYPRE_StructMatrixSetValues(HYPRE_StructMatrix matrix, HYPRE_Int *index, HYPRE_Int nentries, HYPRE_Int *entries, double *values);
HYPRE_Int HYPRE_StructMatrixAssemble(HYPRE_StructMatrix matrix);
HYPRE_Int HYPRE_StructMatrixDestroy(HYPRE_StructMatrix matrix);
HYPRE_Int HYPRE_StructVectorCreate(MPI_Comm comm, HYPRE_StructGrid grid, HYPRE_StructVector *vector);
HYPRE_Int HYPRE_StructVectorInitialize(HYPRE_StructVector vector);
HYPRE_Int HYPRE_StructVectorSetValues(HYPRE_StructVector vector, HYPRE_Int *index, double value);
HYPRE_Int HYPRE_StructVectorSetBoxValues(HYPRE_StructVector vector, HYPRE_Int *ilower, HYPRE_Int *iupper, double *values);
HYPRE_Int HYPRE_StructVectorGetValues(HYPRE_StructVector vector, HYPRE_Int *index, double *value);
HYPRE_Int HYPRE_StructVectorAssemble(HYPRE_StructVector vector);
HYPRE_Int HYPRE_StructVectorDestroy(HYPRE_StructVector vector);

/* Krylov solver + its (ignored) preconditioner object */
HYPRE_Int HYPRE_StructPCGCreate(MPI_Comm comm, HYPRE_StructSolver *solver);
HYPRE_Int HYPRE_StructPCGDestroy(HYPRE_StructSolver solver);
HYPRE_Int HYPRE_StructPCGSetup(HYPRE_StructSolver solver, HYPRE_StructMatrix A, HYPRE_StructVector b, HYPRE_StructVector x);
HYPRE_Int HYPRE_StructPCGSolve(HYPRE_StructSolver solver, HYPRE_StructMatrix A, HYPRE_StructVector b, HYPRE_StructVector x);
HYPRE_Int HYPRE_StructPCGSetTol(HYPRE_StructSolver solver, double tol);
HYPRE_Int HYPRE_StructPCGSetMaxIter(HYPRE_StructSolver solver, HYPRE_Int max_iter);
HYPRE_Int HYPRE_StructPCGSetTwoNorm(HYPRE_StructSolver solver, HYPRE_Int two_norm);
HYPRE_Int HYPRE_StructPCGSetRelChange(HYPRE_StructSolver solver, HYPRE_Int rel_change);
HYPRE_Int HYPRE_StructPCGSetPrecond(HYPRE_StructSolver solver, HYPRE_PtrToStructSolverFcn precond, HYPRE_PtrToStructSolverFcn precond_setup, HYPRE_StructSolver precond_solver);
HYPRE_Int HYPRE_StructPCGSetLogging(HYPRE_StructSolver solver, HYPRE_Int logging);
HYPRE_Int HYPRE_StructPCGSetPrintLevel(HYPRE_StructSolver solver, HYPRE_Int level);
HYPRE_Int HYPRE_StructPCGGetNumIterations(HYPRE_StructSolver solver, HYPRE_Int *num_iterations);
HYPRE_Int HYPRE_StructPCGGetFinalRelativeResidualNorm(HYPRE_StructSolver solver, double *norm);

HYPRE_Int HYPRE_StructSMGCreate(MPI_Comm comm, HYPRE_StructSolver *solver);
HYPRE_Int HYPRE_StructSMGDestroy(HYPRE_StructSolver solver);
HYPRE_Int HYPRE_StructSMGSetup(HYPRE_StructSolver solver, HYPRE_StructMatrix A, HYPRE_StructVector b, HYPRE_StructVector x);
HYPRE_Int HYPRE_StructSMGSolve(HYPRE_StructSolver solver, HYPRE_StructMatrix A, HYPRE_StructVector b, HYPRE_StructVector x);
HYPRE_Int HYPRE_StructSMGSetMemoryUse(HYPRE_StructSolver solver, HYPRE_Int memory_use);
HYPRE_Int HYPRE_StructSMGSetTol(HYPRE_StructSolver solver, double tol);
HYPRE_Int HYPRE_StructSMGSetMaxIter(HYPRE_StructSolver solver, HYPRE_Int max_iter);
HYPRE_Int HYPRE_StructSMGSetRelChange(HYPRE_StructSolver solver, HYPRE_Int rel_change);
HYPRE_Int HYPRE_StructSMGSetZeroGuess(HYPRE_StructSolver solver);
HYPRE_Int HYPRE_StructSMGSetNumPreRelax(HYPRE_StructSolver solver, HYPRE_Int num_pre_relax);
HYPRE_Int HYPRE_StructSMGSetNumPostRelax(HYPRE_StructSolver solver, HYPRE_Int num_post_relax);
HYPRE_Int HYPRE_StructSMGSetLogging(HYPRE_StructSolver solver, HYPRE_Int logging);
HYPRE_Int HYPRE_StructSMGGetNumIterations(HYPRE_StructSolver solver, HYPRE_Int *num_iterations);
HYPRE_Int HYPRE_StructSMGGetFinalRelativeResidualNorm(HYPRE_StructSolver solver, double *norm);

HYPRE_Int HYPRE_StructJacobiCreate(MPI_Comm comm, HYPRE_StructSolver *solver);
HYPRE_Int HYPRE_StructJacobiDestroy(HYPRE_StructSolver solver);
HYPRE_Int HYPRE_StructJacobiSetup(HYPRE_StructSolver solver, HYPRE_StructMatrix A, HYPRE_StructVector b, HYPRE_StructVector x);
HYPRE_Int HYPRE_StructJacobiSolve(HYPRE_StructSolver solver, HYPRE_StructMatrix A, HYPRE_StructVector b, HYPRE_StructVector x);
HYPRE_Int HYPRE_StructJacobiSetTol(HYPRE_StructSolver solver, double tol);
HYPRE_Int HYPRE_StructJacobiSetMaxIter(HYPRE_StructSolver solver, HYPRE_Int max_iter);
HYPRE_Int HYPRE_StructJacobiSetNonZeroGuess(HYPRE_StructSolver solver);
HYPRE_Int HYPRE_StructJacobiGetNumIterations(HYPRE_StructSolver solver, HYPRE_Int *num_iterations);
HYPRE_Int HYPRE_StructJacobiGetFinalRelativeResidualNorm(HYPRE_StructSolver solver, double *norm);

/* ---- recording interface (ours): the last system the reference assembled and solved through the calls above.
 * Matrix layout: values[(i*ny + j)*9 + e] with e in the reference's entry order C,W,SW,S,SE,E,NE,N,NW (src/MGSolver.cpp:286-294);
 * i = first HYPRE index (x cell), j = second (y cell). */
typedef struct refhypre_record {
    int nx, ny, nentries;
    int offsets[9][2];
    const double *matrix;      /* nx*ny*9 */
    const double *rhs;         /* nx*ny   */
    const double *solution;    /* nx*ny   */
    double tol;
    int iterations;            /* of the shim's solver */
    double final_relres;
    int symmetric;             /* after eliminating the Dirichlet rows */
    int nsolves;               /* systems solved since the process started */
} refhypre_record;
const refhypre_record *refhypre_last(void);
/* optional external solver for the recorded system (e.g. a sparse direct solve from the test harness); return 0 on success */
typedef int (*refhypre_solve_fn)(const refhypre_record *sys, double *x_out);
void refhypre_set_solver(refhypre_solve_fn fn);
#ifdef __cplusplus
}
#endif
#endif
