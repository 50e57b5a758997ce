/* _hypre_utilities.h -- SHIM (see HYPRE_struct_ls.h): the reference includes it (src/MGSolver.cpp:44) but uses nothing from it. */
#ifndef IMPEC_REF_SHIM_HYPRE_UTILITIES_H
#define IMPEC_REF_SHIM_HYPRE_UTILITIES_H
#include "mpi.h"
#endif
