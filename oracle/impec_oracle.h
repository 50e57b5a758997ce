/* impec_oracle.h -- CPU ORACLE for the Impec2D IMPES hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (impec2d_b200/, include/)
 * may include, link or call this.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, as the checker.
 *
 * It is a plain-C restatement of what pw359/Impec2D computes on the path named
 * in SURVEY.md section 8(a).  Each function cites the reference file:line it
 * follows (paths relative to /root/reference/).
 *
 * PARITY PINNING -- pinned to the reference itself:
 *   * oracle/_ref/libimpec_ref.so is the reference's own translation units
 *     (its src .cpp files, compiled UNMODIFIED by oracle/Makefile against the shim headers
 *     in oracle/ref_shim/ for blitz, libconfig++, MPI and HYPRE, none of which is
 *     installed here).  tests/test_oracle_vs_ref.py runs the reference's objects
 *     and this oracle on the same inputs: ghost fills, wells, mobilities, the
 *     9-entry matrix with Neumann folding and pin, right-hand sides, Darcy
 *     velocity, time step, TR source step, MUSCL/upwind fluxes, all five
 *     splittings, the Crank-Nicolson matrix/right-hand side and the projection
 *     agree BIT FOR BIT; solves agree to 1e-10.
 *   * tests/golden/ref_vectors.npz holds outputs of that compiled reference
 *     (and of its own main() on elliptic_test.cfg / hyperbolic_test.cfg); the
 *     oracle is checked against them wherever the tests run.
 *   * the reference's analytic known answers (orders 2.00 / 1.85-1.88).
 * What stays UNPINNED: iteration-level parity with HYPRE's SMG-PCG.  HYPRE
 * (hypre-2.8.0b, INSTALL:44-46) is a third-party library outside the reference
 * tree; the PCG loop below restates its published struct-PCG stopping rule
 * (two-norm, relative, recurrence residual) with a geometric multigrid cycle of
 * our own design in place of SMG.  Any solver reaching the same tolerance on the
 * same matrix is conformant.
 *
 * Array conventions (src/Impec2DSolver.cpp:151-182, src/GenSimDefs.h:25-31):
 *   scalar fields c,p : (N1+2g) x (N2+2g) doubles, C order, j fastest, ghost g
 *   u, v (separate)   : (N1+1+2g) x (N2+1+2g) doubles; u(i,j) on the left face
 *                       of cell (i,j), v(i,j) on its bottom face
 *   q                 : N1*N2 doubles, q[i*N2+j]
 */
#ifndef IMPEC_ORACLE_H_
#define IMPEC_ORACLE_H_
#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_PARALLEL = 0, ORC_DIAGONAL = 1, ORC_HORIZONTAL = 2 };
enum { ORC_MEANCONC = 0, ORC_HARMONIC = 1 };
enum { ORC_ALTERNATING = 0, ORC_X_FIRST = 1, ORC_Y_FIRST = 2, ORC_STRANG = 3, ORC_SYMMETRIC_PARALLEL = 4 };
enum { ORC_MUSCL = 0, ORC_UPWIND = 1 };
enum { ORC_EXPLICIT = 0, ORC_CRANKNICOLSON = 1 };
enum { ORC_PROFILE_SINE = 0, ORC_PROFILE_TOPHAT = 1, ORC_PROFILE_GAUSS = 2, ORC_PROFILE_HSINE = 3 };

typedef struct orc_params {
    int    N1, N2, g;                 /* cells and ghost width (2)                 */
    double x_lb, x_ub, y_lb, y_ub;
    double mu_res, mu_inj;
    int    avgproc;                   /* ORC_MEANCONC | ORC_HARMONIC               */
    double nu;                        /* 5-point / rotated-stencil blend           */
    int    is_test, testno;           /* elliptic manufactured test selects K(x,y) */
    int    stype;                     /* ORC_PARALLEL | ORC_DIAGONAL | ORC_HORIZONTAL */
    double porosity, cfl, inj_fluid_conc;
    int    diffusion;
    double alpha_l, alpha_t, dt_diff_limit;
    int    ptype;                     /* ORC_EXPLICIT | ORC_CRANKNICOLSON          */
    double eps, para_eps;
    int    hyper_method;              /* ORC_MUSCL | ORC_UPWIND                    */
    int    dim_splitting;
    /* multigrid-PCG controls (ours; the reference has HYPRE SMG-PCG here) */
    int    max_iter;                  /* 10000 in the reference (MGSolver.cpp:89)  */
    int    mg_max_levels;             /* 0 = as many as the grid allows            */
    int    mg_nmin;                   /* stop coarsening at min(N1,N2) <= mg_nmin  */
    int    mg_ncoarse;                /* symmetric GS sweeps on the coarsest level */
    int    mg_w_from, mg_w_to;        /* levels visited twice (gamma=2), inclusive */
} orc_params;

typedef struct orc_state {            /* mutable solver state (HyperbolicSolver.h:18-19) */
    double total_mass_outflux;
    int    prev_first_sweep;          /* 0 = X_SWEEP, 1 = Y_SWEEP */
    int    last_its, last_parabolic_its;
    double last_relres, last_true_relres;
} orc_state;

void   orc_default_params(orc_params *P);
void   orc_state_init(orc_state *S);

/* ghost fills: src/Util.cpp:215-256, 283-324, 331-438 */
void   orc_fill_periodic_scalar(double *a, int N1, int N2, int g);
void   orc_fill_neumann_scalar(double *a, int N1, int N2, int g);
void   orc_fill_neumann_velocity(double *u, double *v, int N1, int N2, int g);
void   orc_fill_boundaries(const orc_params *P, double *c);          /* HyperbolicSolver.cpp:56-61 */

/* viscosity / mobility: src/EllipticSolver.cpp:418-524, 692-754 */
double orc_mu(const orc_params *P, double c);
double orc_K11(const orc_params *P, double x, double y);
double orc_K22(const orc_params *P, double x, double y);

/* pressure matrix: src/MGSolver.cpp:268-570, 710-841.  A is N1*N2*9, HYPRE entry
 * order C,W,SW,S,SE,E,NE,N,NW (MGSolver.cpp:164-166, 286-294); b = -q. */
void   orc_stencil_cell(const orc_params *P, const double *c, int i, int j, double coef[9]);
void   orc_assemble_elliptic(const orc_params *P, const double *c, const double *q, double *A, double *b);
void   orc_matvec9(int N1, int N2, const double *A, const double *x, double *y);

/* pressure solve (a8-a11): returns 0 on convergence, 1 if max_iter hit */
int    orc_solve_elliptic(const orc_params *P, orc_state *S, const double *c, const double *q, double *p);

/* generic SPD-form 9-point MG-PCG on an assembled matrix (sign = -1 elliptic, +1 parabolic) */
int    orc_mgpcg(const orc_params *P, int N1, int N2, const double *A, const double *b, int sign,
                 double tol, double *x, int *its, double *relres, double *true_relres);
/* one multigrid cycle applied to r (zero initial guess), for kernel-level parity tests */
int    orc_mg_apply(const orc_params *P, int N1, int N2, const double *A, int sign, const double *r, double *z);

/* Darcy velocity (a12): src/EllipticSolver.cpp:541-683 */
void   orc_calc_velocities(const orc_params *P, double *p, const double *c, double *u, double *v);

/* time step (a13): src/HyperbolicSolver.cpp:93-198 */
double orc_max_stable_dt(const orc_params *P, const double *u, const double *v);
double orc_max_stable_advection_dt(const orc_params *P, const double *u, const double *v);
double orc_max_stable_dt_combined(const orc_params *P, const double *u, const double *v);

/* source step (a15): src/TRIntegrator.cpp:49-71; returns the signed mass difference */
double orc_tr_integrate(const orc_params *P, double *c, const double *q, double dt);

/* fluxes and sweeps (a16-a18) */
void   orc_flux_F(const orc_params *P, int method, const double *c, const double *u, double dt, double *F);
void   orc_flux_G(const orc_params *P, int method, const double *c, const double *v, double dt, double *G);
void   orc_advect_x(const orc_params *P, double *c, const double *u, double dt);
void   orc_advect_y(const orc_params *P, double *c, const double *v, double dt);
void   orc_advect_split(const orc_params *P, orc_state *S, double *c, const double *u, const double *v, double dt);

/* dispersion (a19): src/MGSolver.cpp:846-976, 997-1371 */
void   orc_dispersion_F(const orc_params *P, const double *u, const double *v, int i, int j, double *d11, double *d12);
void   orc_dispersion_G(const orc_params *P, const double *u, const double *v, int i, int j, double *d22, double *d21);
void   orc_assemble_parabolic(const orc_params *P, const double *c, const double *u, const double *v, double dt,
                              double *A, double *b);
int    orc_solve_parabolic(const orc_params *P, orc_state *S, double *c, const double *u, const double *v, double dt);

/* transport step (a14): src/HyperbolicSplitSolver.cpp:162-198 */
int    orc_transport_step(const orc_params *P, orc_state *S, double *c, const double *u, const double *v,
                          const double *q, double dt);

/* one IMPES step (a20): src/Impec2DSolver.cpp:339-389; dt_clip<=0 means no clipping. Returns dt used. */
double orc_impes_step(const orc_params *P, orc_state *S, double *p, double *c, double *u, double *v,
                      const double *q, double t, double t_next_output);

/* wells: src/Util.cpp:444-639, src/Impec2DSolver.cpp:241-280 */
double orc_exact_vol_frac_circle(double x1, double y1, double r, double dx, double dy);
void   orc_init_source(const orc_params *P, double I, double well_rc, int well_profile /*0 circle,2 onecell*/, double *q);
int    orc_bicgstab9(int N1, int N2, const double *A, const double *b, double tol, int max_iter, double *x, int *its, double *relres);   /* non-symmetric CN matrix (alpha_l != alpha_t) */
int    orc_init_horizontal_front(const orc_params *P, double I, double well_width, int mode /*0 pulsing,1 linear,2 constant*/, double q_freq,
                                 double q_high_multiple, double tstart, double tfinal, double t, double *q);   /* src/Util.cpp:607-639 */

/* test profiles and norms: src/Util.cpp:55-147 */
double orc_profile(const orc_params *P, int profile, double x, double y);
double orc_advection_norm(const orc_params *P, const double *c, int profile, int norm /*1|2*/);

/* elliptic manufactured test: src/EllipticSolver.cpp:319-410 */
void   orc_init_elliptic_test(const orc_params *P, double *p_ex, double *q, double *u_ex, double *v_ex);

/* diagnostics: src/Impec2DSolver.cpp:530-620 */
double orc_total_mass(const orc_params *P, const double *c);
double orc_divergence_error(const orc_params *P, const double *u, const double *v, const double *q);
double orc_boundary_velocity_error(const orc_params *P, const double *u, const double *v);

/* synthetic developed state; counter-based noise, decomposition independent */
/* Impec2DSolver::Project / ParallelProjector::FillDiagonal  src/Impec2DSolver.cpp:690-716, src/ParallelProjector.cpp:91-199 */
void   orc_project_dims(const orc_params *P, int *dN1, int *dN2);
void   orc_project_parallel(const orc_params *P, const double *a_ghosted, int dN1, int dN2, double *out);
void   orc_synth_state(const orc_params *P, unsigned long long seed, double *c);

int    orc_num_threads(void);
void   orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
