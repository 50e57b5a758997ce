/* impec2d_b200.h -- C ABI of the B200-native IMPES hot path for pw359/Impec2D.
 *
 * The reference has no FFI boundary: its solvers are C++ classes new-ed in
 * Impec2DSolver::Initialize (src/Impec2DSolver.cpp:131-133).  This header is the boundary a
 * maintainer binds instead: each entry point names the reference method whose loop nest it
 * replaces (paths relative to the reference tree).  The host-side mirror of the reference classes
 * (impec2d_b200/host/, same class and method names) is a thin layer over these calls; see
 * INTEGRATION.md for the binding.
 *
 * Conventions
 *   - plain C types only; every pointer is a HOST pointer unless the name ends in _dev
 *   - host field layouts are the reference's (src/Impec2DSolver.cpp:151-182):
 *       scalar c,p : (N1+2g) x (N2+2g) doubles, C order, j fastest, g = num_ghostcells
 *       velocity   : (N1+1+2g) x (N2+1+2g) x 2 doubles, component fastest (blitz Array3d)
 *       q          : N1*N2 doubles, q[i*N2+j]
 *   - one context per process / GPU; calls on one context must be serialised by the caller; all
 *     device work is enqueued on the context's stream and calls that return scalars synchronise it
 *   - return value: 0 = ok, negative = error, positive = warning, call completed (impec_last_error gives the text); never exit()
 *   - there is NO CPU fallback: creating a context without a CUDA device fails
 */
#ifndef IMPEC2D_B200_H_
#define IMPEC2D_B200_H_
#ifdef __cplusplus
extern "C" {
#endif

#define IMPEC_OK 0
#define IMPEC_WARN_NOCONV 1        /* POSITIVE = warning: a solver stopped at max_iter; the call COMPLETED with the last iterate,
                                      as upstream (src/MGSolver.cpp:612-622 never checks convergence); info->converged = 0 */
#define IMPEC_ERR_ARG (-1)
#define IMPEC_ERR_CUDA (-2)
#define IMPEC_ERR_UNSUPPORTED (-3)
#define IMPEC_ERR_NOCONV (-4)      /* the solver broke down (NaN): fatal                                              */
#define IMPEC_ERR_NCCL (-5)

/* enum values follow src/SimParams.h:13-27 */
enum { IMPEC_PARALLEL = 0, IMPEC_DIAGONAL = 1, IMPEC_HORIZONTAL = 2 };                       /* S_Type        */
enum { IMPEC_MEANCONC = 0, IMPEC_HARMONIC = 1 };                                             /* AVG_PROC      */
enum { IMPEC_ALTERNATING = 0, IMPEC_X_FIRST = 1, IMPEC_Y_FIRST = 2, IMPEC_STRANG = 3, IMPEC_SYMMETRIC_PARALLEL = 4 }; /* DIM_SPLITTING */
enum { IMPEC_MUSCL = 0, IMPEC_UPWIND = 1 };                                                  /* hyper_method  */
enum { IMPEC_EXPLICIT = 0, IMPEC_CRANKNICOLSON = 1 };                                        /* P_Type        */

/* The subset of SimParams (src/SimParams.h:29-173) the hot path reads. */
typedef struct impec_params {
    int    N1, N2, g;                 /* x_cells, y_cells, num_ghostcells (must be 2)              */
    double x_lb, x_ub, y_lb, y_ub;
    double mu_res, mu_inj;
    int    avgproc;
    double nu;
    int    is_test, testno;
    int    stype;
    double porosity, cfl, inj_fluid_conc;
    int    diffusion;
    double alpha_l, alpha_t, dt_diff_limit;
    int    ptype;
    double eps, para_eps;
    int    hyper_method;
    int    dim_splitting;
    /* solver controls (ours: the reference has HYPRE StructPCG+SMG here, src/MGSolver.cpp:594-629) */
    int    max_iter;                  /* 10000 upstream                                            */
    int    mg_max_levels;             /* 0 = as deep as the grid allows                            */
    int    mg_nmin;                   /* stop coarsening once min(N1,N2) <= mg_nmin (default 4)    */
    int    mg_ncoarse;                /* symmetric GS sweeps on the coarsest level (default 8)     */
    int    mg_w_from, mg_w_to;        /* levels visited twice per cycle (default 2..4; see impec_mg_w_range) */
} impec_params;

typedef struct impec_diag {           /* src/Impec2DSolver.h:16-25 plus the per-step extrema        */
    double total_mass;                /* GetTotalMass            src/Impec2DSolver.cpp:607-620      */
    double divergence_error;          /* CalculateDivergenceError :542-557                          */
    double boundary_velocity_error;   /* CalculateVelocityBoundaryError :562-588                    */
    double total_mass_outflux;        /* HyperbolicSolver::GetTotalMassOutflux                      */
    double c_min, c_max;
} impec_diag;

typedef struct impec_integral_diag {  /* integral diagnostics of a displacement run (north-star check 2)          */
    double produced_fraction;         /* |q|-weighted mean c over the production cells (q < 0): what the wells produce             */
    double mixing_zone_length;        /* area of {lo <= c <= hi} / injector-producer distance hypot(Lx, Ly)                        */
    double mixing_zone_cells;         /* number of cells in that band (exact integer)                                              */
    double mass_in_place;             /* sum c dx dy                                                                                */
} impec_integral_diag;

typedef struct impec_solve_info {
    int    iterations;                /* PCG iterations (MGSolver::num_iterations)                  */
    double relres;                    /* recurrence ||r||/||b||, the quantity HYPRE's PCG tests      */
    double true_relres;               /* ||b - A x||/||b|| recomputed after the solve               */
    int    converged;                 /* 1 = recurrence criterion met; 2 = stopped by the optional stagnation guard     */
} impec_solve_info;

typedef struct impec_ctx impec_ctx;

/* Communication back end for host-memory builds (the CPU-emulated test library): the product library talks NCCL itself.
 * All buffers are host pointers.  op: 0 sum, 1 max, 2 min.  Return 0 on success. */
typedef struct impec_comm_callbacks {
    void *user;
    /* send n doubles to rank-1 (send_lo) and rank+1 (send_hi), receive n from each; NULL where there is no neighbour */
    int (*sendrecv)(void *user, const double *send_lo, double *recv_lo, const double *send_hi, double *recv_hi, long long n);
    int (*allreduce)(void *user, double *buf, int n, int op);
    int (*allgather)(void *user, const double *send, double *recv, long long n_per_rank);
    int (*bcast)(void *user, double *buf, long long n, int root);
} impec_comm_callbacks;

void  impec_default_params(impec_params *p);
const char *impec_version(void);
int   impec_is_emulated(void);        /* 1 only for the CPU-emulated test build (never shipped)    */

/* rank/nranks: slab decomposition along i over the GPUs of one node (x_cells a multiple of 2 * nranks with at least 8 rows per slab, y_cells even:
 * the slab levels need one coarser level below them; anything else is refused with IMPEC_ERR_ARG and a message); nccl_unique_id = 128 bytes from
 * impec_nccl_unique_id on rank 0 (NULL when nranks == 1).  cuda_stream: a cudaStream_t to enqueue on, or
 * NULL for a context-owned stream. */
int   impec_nccl_unique_id(char id_out[128]);
/* Host placement for one-process-per-GPU runs (ours; upstream leaves process placement to mpirun, src/Driver.cpp:58-87): restricts the CALLING
 * thread (threads it creates later inherit) to the CPUs that are local to `device` -- /sys/bus/pci/devices/<bus id>/local_cpulist, intersected
 * with the affinity mask the process already has -- so that pinned host buffers allocated afterwards and the host scan of q live on the NUMA node
 * the GPU hangs off.  Call before allocating host buffers / creating the context.  Returns the number of CPUs bound to (> 0) and writes the
 * kernel's list ("0-55,112-167") to cpulist_out when given; IMPEC_ERR_UNSUPPORTED, with the mask unchanged, when the topology is not visible or
 * none of those CPUs is allowed to this process. */
int   impec_bind_host_to_device(int device, char *cpulist_out, int len);
int   impec_ctx_create(const impec_params *p, int device, int rank, int nranks, const char *nccl_unique_id,
                       void *cuda_stream, impec_ctx **out);
void  impec_ctx_destroy(impec_ctx *ctx);
const char *impec_last_error(const impec_ctx *ctx);
int   impec_synchronize(impec_ctx *ctx);
/* transport used between the ranks of this context: "single", "nccl", "nvlink-p2p+nccl" (halo rows and scalar
 * reductions through mapped peer memory, NCCL for bootstrap and large collectives; needs peer access between the
 * GPUs, IMPEC_P2P=0 disables it) or "callbacks" (test harness) */
const char *impec_comm_backend(const impec_ctx *ctx);
int   impec_ctx_set_comm_callbacks(impec_ctx *ctx, const impec_comm_callbacks *cb);
/* slab owned by this context: rows [row0, row0 + nrows) of the global grid (everything when nranks == 1) */
int   impec_local_rows(const impec_ctx *ctx, int *row0, int *nrows);

/* host <-> device, reference layouts (Impec2DSolver owns p,c,v,q: src/Impec2DSolver.h:31-46) */
int   impec_upload_c(impec_ctx *ctx, const double *c_ghosted);
int   impec_upload_p(impec_ctx *ctx, const double *p_ghosted);
int   impec_upload_q(impec_ctx *ctx, const double *q);
int   impec_upload_velocity(impec_ctx *ctx, const double *v_aos);
int   impec_download_c(impec_ctx *ctx, double *c_ghosted);
int   impec_download_p(impec_ctx *ctx, double *p_ghosted);
int   impec_download_velocity(impec_ctx *ctx, double *v_aos);
/* multi-GPU logging: the WHOLE field ("c", "p": ghosted scalar layout; "v": AoS velocity) assembled from the slabs of all ranks into
 * every rank's host array (collective).  Ghost rows and the wall face row are left untouched.  The reference keeps full copies of
 * p, c, v on every rank (src/MGSolver.cpp:670-706 MapHYPREVectorToVertex + Allgather); here that happens only when asked for. */
int   impec_gather_rows(impec_ctx *ctx, const char *field, double *host_full);
/* Multi-rank note: upload/download take the GLOBAL host arrays and move only this rank's slab (plus its ghost rows);
 * the *_rows variants take just the slab: c,p as (nrows+2g) x (N2+2g), q as nrows x N2. */
int   impec_upload_c_rows(impec_ctx *ctx, const double *c_slab_ghosted);
int   impec_upload_q_rows(impec_ctx *ctx, const double *q_slab);
int   impec_download_c_rows(impec_ctx *ctx, double *c_slab_ghosted);
/* bytes of field data this context has moved over PCIe so far (counted from the copies the library issues; q is scanned on the
 * host and only its non-zero cells are sent when it is sparse) */
int   impec_transfer_bytes(const impec_ctx *ctx, unsigned long long *h2d, unsigned long long *d2h);
/* device-resident alternatives for callers that already hold device memory (padded layout) */
int   impec_field_layout(const impec_ctx *ctx, int *ld, int *row0, int *col0, int *rows);
void *impec_field_dev(impec_ctx *ctx, const char *name);   /* "c","p","q","u","v" */
int   impec_set_q_dev_changed(impec_ctx *ctx);             /* rebuild the well list after writing q on the device */

/* a1   HyperbolicSolver::FillBoundaries            src/HyperbolicSolver.cpp:56-61 */
int   impec_fill_boundaries(impec_ctx *ctx);
/* a8-a11 EllipticSolver::Solve -> MGSolver::SolveLinearMultigrid -> Solve_Elliptic_cc_to_c
 *      src/EllipticSolver.cpp:133-142, src/MGSolver.cpp:145-149, 223-248 */
int   impec_pressure_solve(impec_ctx *ctx, double eps, impec_solve_info *info);
/* a12  EllipticSolver::CalcVelocities              src/EllipticSolver.cpp:660-683 */
int   impec_calc_velocities(impec_ctx *ctx);
/* a13  HyperbolicSolver::MaxStableTimestep         src/HyperbolicSolver.cpp:93-198 */
int   impec_max_stable_dt(impec_ctx *ctx, double *dt);
/* a15  HyperbolicSolver::SourceSplit / TRIntegrator::Integrate  src/HyperbolicSolver.cpp:201-205, src/TRIntegrator.cpp:49-71 */
int   impec_source_split(impec_ctx *ctx, double dt);
/* a16  HyperbolicSplitSolver::Advect_X / Advect_Y / AdvectSplit src/HyperbolicSplitSolver.cpp:51-159 */
int   impec_advect_x(impec_ctx *ctx, double dt);
int   impec_advect_y(impec_ctx *ctx, double dt);
int   impec_advect_split(impec_ctx *ctx, double dt);
/* CalculateIntercellFlux_F / _G (src/HyperbolicSplitSolver.h:19-22; src/MUSCLSplitSolver.cpp:102-189, src/UpwindSplitSolver.cpp:43-61):
 * the intercell fluxes of the current c and velocity for a step dt as a HOST array in the reference's shape -- dir 0: F[(N1+1)][N2+2g]
 * (index base (0,-g)), dir 1: G[N1+2g][N2+1] (index base (-g,0)).  The sweeps never materialise these arrays.  Slab contexts: every rank fills its
 * own rows (with the halo rows it shares with its neighbours) of the caller's whole-grid array. */
int   impec_intercell_flux(impec_ctx *ctx, int dir, double dt, double *out_host);
/* a19  HyperbolicSolver::DiffusionSplitCN -> MGSolver::Solve_Parabolic_cc_to_c  src/HyperbolicSolver.cpp:213-223, src/MGSolver.cpp:952-976 */
/*      alpha_l == alpha_t: symmetric links operator, PCG with a one-march symmetric Gauss-Seidel preconditioner; alpha_l != alpha_t: the full
 *      non-symmetric 9-point operator of src/MGSolver.cpp:923-948 applied matrix-free, BiCGStab.  Both to para_eps, x0 = 0. */
int   impec_diffusion_split(impec_ctx *ctx, double dt, impec_solve_info *info);
/* a14  HyperbolicSplitSolver::Solve                src/HyperbolicSplitSolver.cpp:162-198 */
int   impec_transport_step(impec_ctx *ctx, double dt);
/* a20  one iteration of the while loop in Impec2DSolver::Solve  src/Impec2DSolver.cpp:339-389.
 *      t_next_output <= t disables output-time clipping.  dt_used and info may be NULL. */
int   impec_impes_step(impec_ctx *ctx, double t, double t_next_output, double *dt_used, impec_solve_info *info);
/* a20  diagnostics of the loop body                src/Impec2DSolver.cpp:391-415, 530-620 */
int   impec_diagnostics(impec_ctx *ctx, impec_diag *out);
/* breakthrough / recovery / mixing-zone inputs as device reductions: no field is downloaded (2.1 GB per sample at 16384^2 otherwise).
 * lo, hi: the concentration band of the mixing zone (0.1, 0.9 in impec2d_b200/diagnostics.py) */
int   impec_integral_diagnostics(impec_ctx *ctx, double lo, double hi, impec_integral_diag *out);

/* 8f  Impec2DSolver::Project / ParallelProjector::FillDiagonal  src/Impec2DSolver.cpp:690-716, src/ParallelProjector.cpp:91-199
 *     field 0 = c, 1 = p (ghosts must be filled: they are after impec_impes_step); out_host: dN1 x dN2 doubles, row-major.
 *     Slab contexts: collective, every rank receives the whole projected field (chunks of diagonal rows are summed over the ranks). */
int   impec_project_dims(const impec_ctx *ctx, int *dN1, int *dN2);
int   impec_project_to_diagonal(impec_ctx *ctx, int field, double *out_host);

/* 8f  restart: the reference reads a text file of c only (init_file, src/Impec2DSolver.cpp:194-220).  A checkpoint is the exact
 *     device state of this rank (c, p, u, v with ghosts, produced mass, sweep parity, t, step): the restarted run continues bit
 *     for bit.  One file per rank. */
int   impec_checkpoint_save(impec_ctx *ctx, const char *path, double t, long long step);
int   impec_checkpoint_load(impec_ctx *ctx, const char *path, double *t, long long *step);

/* benchmark / test hooks */
int   impec_build_operator(impec_ctx *ctx);                              /* a2-a8: links from the current c */
int   impec_vcycle(impec_ctx *ctx, const double *r_host, double *z_host); /* one MG cycle; NULL/NULL = on the internal residual */
int   impec_mg_levels(const impec_ctx *ctx);
int   impec_mg_w_range(const impec_ctx *ctx, int *from, int *to);        /* levels the cycle visits twice, as in effect (8 or more ranks with 10-50 M cells each: 2..3 instead of the default 2..4) */
long long impec_kernel_launches(const impec_ctx *ctx);                   /* kernels launched by this context so far */
int   impec_last_parabolic_info(const impec_ctx *ctx, impec_solve_info *info); /* the last Crank-Nicolson solve run inside impec_transport_step / impec_impes_step (MGSolver::last_parabolic_its, src/MGSolver.cpp:972) */
long long impec_overlapped_exchanges(const impec_ctx *ctx);              /* halo exchanges issued concurrently with the interior row chunks of their consumer (multi-GPU) */
int   impec_reset_state(impec_ctx *ctx);                                 /* produced mass := 0, sweep parity := Y */
/* options: "warm_start" (x0 = previous p instead of the reference's 0), "stagnation_guard" (stop PCG at the FP64
 * residual floor instead of iterating the recurrence residual down to eps), "mg_fused", "mg_links_fp32",
 * "mg_graph", "mg_small", "mg_w_from", "mg_w_to", "mg_ncoarse", "max_iter".  Defaults reproduce the reference's semantics.
 * Kernel-generation switches (results identical, for A/B measurements and tests): "mg_march" (register-march multigrid kernels
 * on the big levels, default 1), "matvec_march" (1), "sweep_shared" (MUSCL with one reconstruction per cell, 1),
 * "sweep_y_pairs" (y sweep with column pairs per lane, 1), "mg_fused_mid" (1), "sgs_march" (one-march Crank-Nicolson preconditioner, 1),
 * "l2_prefetch" / "sweep_prefetch" (rows of L2 software prefetch ahead of the marching kernels, 4 / 4; 0 = off),
 * "halo_overlap" (multi-GPU: halo exchange concurrent with the interior row chunks of its consumer, 1).
 * Environment, read at context creation: IMPEC_MARCH_MIN_CELLS, IMPEC_TH32_BELOW, IMPEC_TH16_BELOW, IMPEC_DIST_MIN_CELLS,
 * IMPEC_L2_PREFETCH ("pre,post,matvec,sgs"), IMPEC_SWEEP_PREFETCH, IMPEC_SWEEP_Y_PAIRS, IMPEC_HALO_OVERLAP, IMPEC_P2P, IMPEC_P2P_HALO,
 * IMPEC_P2P_GATHER. */
int   impec_set_option(impec_ctx *ctx, const char *key, double value);
/* average device time of one named piece of the path ("smooth_sweep", "residual", "vcycle", "sweep_x", ...) over
 * `reps` launches, measured with CUDA events on the context stream */
int   impec_profile_kernel(impec_ctx *ctx, const char *name, int reps, double *avg_ms);

/* host-side setup helpers (no device work) */
/* Impec2DSolver::InitializeSource  src/Impec2DSolver.cpp:241-280, src/Util.cpp:444-586; well_profile 0 circle, 2 one_cell */
int   impec_init_source(const impec_params *p, double I, double well_rc, int well_profile, double *q);
/* InitHorizontalFront (src/Util.cpp:607-639; source_type "horizontal"): injection layer of round(well_width (x_ub-x_lb)/dx) rows at
 * i = 0.., production layer at the far end.  mode 0 pulsing (q_freq, q_high_multiple), 1 linear ramp tstart..tfinal, 2 constant; t = time. */
int   impec_init_horizontal_front(const impec_params *P, double I, double well_width, int mode, double q_freq, double q_high_multiple,
                                  double tstart, double tfinal, double t, double *q);
/* developed synthetic concentration for benchmarks; writes the interior of a ghosted array */
int   impec_synth_state(const impec_params *p, unsigned long long seed, double *c_ghosted);
/* the same restricted to global rows [row0, row0+nrows): q_slab is nrows x N2, c_slab_ghosted is (nrows+2g) x (N2+2g) */
int   impec_init_source_rows(const impec_params *p, double I, double well_rc, int well_profile, int row0, int nrows, double *q_slab);
int   impec_synth_state_rows(const impec_params *p, unsigned long long seed, int row0, int nrows, double *c_slab_ghosted);

#ifdef __cplusplus
}
#endif
#endif
