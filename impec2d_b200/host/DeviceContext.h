// DeviceContext.h -- one impec_ctx (one GPU) per SimParams, shared by the solver objects that borrow the same
// SimParams* (the reference's solvers all borrow sp and never own it: src/EllipticSolver.cpp:59, src/HyperbolicSolver.cpp:63).
#ifndef DEVICECONTEXT_H_
#define DEVICECONTEXT_H_
#include "SimParams.h"
#include "GenSimDefs.h"
#include "../../include/impec2d_b200.h"

class DeviceContext {
public:
    // returns the context for sp, creating or re-creating it when the grid or physics in sp changed
    static impec_ctx *Get(SimParams *sp);
    static void Release(SimParams *sp);
    static impec_params ToParams(const SimParams *sp);
    // make the device copies current for an operation that reads them
    static void Use(impec_ctx *ctx, const Array2d *c, const Array2d *p, const Array3d *v, const double *q, SimParams *sp);
    static void InvalidateSource(SimParams *sp);          // call after changing q in place
    static void Check(impec_ctx *ctx, int rc, const char *what);   // negative: fatal (exit 1, upstream's convention); positive: warning
    static void SetMethod(impec_ctx *ctx, int method);            // MUSCL / upwind of the calling solver object, per call
    // Multi-GPU (src/Driver.cpp:58-87 hands myid / num_procs to Impec2DSolver): one process per GPU, rank r owns rows [r N1/P, (r+1) N1/P).
    // Call before the first Get().  The NCCL id travels through a small file (rank 0 writes it, the others wait for it):
    // IMPEC_NCCL_ID_FILE, default /tmp/impec2d_nccl_id.<MASTER_PORT or parent pid>; the device is IMPEC_DEVICE, else LOCAL_RANK, else rank.
    static void SetRanks(int rank, int nranks);
    static int Rank();
    static int NumRanks();
    // collective: the whole field assembled from all slabs into every rank's host array (for logging)
    static void GatherGlobal(impec_ctx *ctx, const Array2d *c, const Array2d *p, const Array3d *v);
};
#endif
