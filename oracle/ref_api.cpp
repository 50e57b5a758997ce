// ref_api.cpp -- TEST INFRASTRUCTURE ONLY (part of oracle/_ref/libimpec_ref.so; never linked into or loaded by the product).
//
// A flat C interface to the REFERENCE's own objects.  oracle/Makefile.ref compiles the reference's translation units UNMODIFIED from
// /root/reference/src (SimParams, Util, TRIntegrator, TimeIntegrator, HyperbolicSolver, HyperbolicSplitSolver, MUSCLSplitSolver,
// UpwindSplitSolver, EllipticSolver, MGSolver, ...) against the shim headers in oracle/ref_shim/ (blitz, libconfig++, MPI, HYPRE) and
// links them with this file.  Every function below only constructs the reference's classes the way Impec2DSolver::Initialize does
// (src/Impec2DSolver.cpp:123-142) and calls their methods; no numerical statement of the reference is restated here.  It is what
// pins oracle/impec_oracle.c (the portable C restatement) to the reference itself: tests/test_oracle_vs_ref.py.
//
// Array layouts at this interface are those of include/impec2d_b200.h (= the reference's blitz arrays flattened in C order):
//   c, p : (N1+2g) x (N2+2g);  v : (N1+1+2g) x (N2+1+2g) x 2;  q : N1*N2.
#include "MGSolver.h"
#include "MUSCLSplitSolver.h"
#include "UpwindSplitSolver.h"
#include "Util.h"
#include "ParallelProjector.h"
#include "HYPRE_struct_ls.h"
#include <string>
#include <cstring>
#include <cstdlib>
#include <new>
#include <omp.h>

namespace {
// protected members of the reference's classes are reached through using-declarations in derived classes (the reference's
// sources stay untouched)
struct MGOpen : public MGSolver {
    explicit MGOpen(SimParams *sp) : MGSolver(sp) {}
    using EllipticSolver::mu;
    using EllipticSolver::KoverMu_11_im12j;
    using EllipticSolver::KoverMu_22_ijm12;
    using EllipticSolver::KoverMu_11_ip12jp12;
    using EllipticSolver::KoverMu_22_ip12jp12;
};
struct MusclOpen : public MUSCLSplitSolver {
    explicit MusclOpen(SimParams *sp) : MUSCLSplitSolver(sp) {}
    using HyperbolicSplitSolver::AdvectSplit;
    using HyperbolicSolver::SourceSplit;
    using HyperbolicSplitSolver::Solve;
};
struct UpwindOpen : public UpwindSplitSolver {
    explicit UpwindOpen(SimParams *sp) : UpwindSplitSolver(sp) {}
    using HyperbolicSplitSolver::AdvectSplit;
    using HyperbolicSolver::SourceSplit;
    using HyperbolicSplitSolver::Solve;
};

struct Ref {
    SimParams *sp = nullptr;
    MGOpen *es = nullptr;
    MusclOpen *hs = nullptr;
    UpwindOpen *up = nullptr;
    Array2d c, p;
    Array3d v;
    double *q = nullptr;
};

// ResizeArrays as upstream (src/Impec2DSolver.cpp:151-182)
void resize_arrays(Ref *R) {
    SimParams *sp = R->sp;
    Idx2d idx; idx[0] = -sp->p_bounds; idx[1] = -sp->p_bounds;
    R->p.resize(sp->N1 + 2 * sp->p_bounds, sp->N2 + 2 * sp->p_bounds); R->p.reindexSelf(idx); R->p = 0.;
    R->c.resize(sp->N1 + 2 * sp->p_bounds, sp->N2 + 2 * sp->p_bounds); R->c.reindexSelf(idx); R->c = 0.;
    R->v.resize((sp->N1 + 1) + 2 * sp->v_bounds, (sp->N2 + 1) + 2 * sp->v_bounds, 2);
    Idx3d i3; i3[0] = -sp->v_bounds; i3[1] = -sp->v_bounds; i3[2] = 0;
    R->v.reindexSelf(i3); R->v = 0.;
}
HyperbolicSplitSolver *pick(Ref *R, int method) { return method == 1 ? (HyperbolicSplitSolver *)R->up : (HyperbolicSplitSolver *)R->hs; }
}  // namespace

extern "C" {

void *ref_open(const char *cfg_path) {
    Ref *R = new Ref();
    // SimParams leaves members uninitialised that it later reads (see ref_zero_new.cpp): construct it in zero-filled memory.  (The
    // zero-filling operator new of ref_zero_new.cpp covers the reference's own `new SimParams` in its executable; inside a host
    // process that loaded another C++ runtime first, the dynamic linker may bind operator new elsewhere, hence placement new here.)
    void *mem = calloc(1, sizeof(SimParams));
    R->sp = new (mem) SimParams(std::string(cfg_path));    // exits the process on a bad file, as upstream
    R->es = new MGOpen(R->sp);
    R->hs = new MusclOpen(R->sp);
    R->up = new UpwindOpen(R->sp);
    R->hs->AssignMGSolver(R->es);
    R->up->AssignMGSolver(R->es);
    resize_arrays(R);
    R->q = new double[(size_t)R->sp->N1 * R->sp->N2]();
    return R;
}
void ref_close(void *h) {
    Ref *R = (Ref *)h;
    if (!R) return;
    delete R->hs; delete R->up; delete R->es; delete[] R->q; R->sp->~SimParams(); free(R->sp); delete R;
}
void ref_dims(void *h, int *N1, int *N2, int *g) { Ref *R = (Ref *)h; *N1 = R->sp->N1; *N2 = R->sp->N2; *g = R->sp->c_bounds; }
// the derived quantities SimParams computes (src/SimParams.cpp:449-456) plus a few parsed values, for cross-checking the cfg reader
void ref_params(void *h, double *out) {
    SimParams *sp = ((Ref *)h)->sp;
    const double v[] = {sp->dx, sp->dy, sp->x_lb, sp->x_ub, sp->y_lb, sp->y_ub, sp->mu_res, sp->mu_inj, sp->nu, sp->cfl, sp->porosity, sp->eps, sp->para_eps,
                        sp->alpha_l, sp->alpha_t, sp->dt_diff_limit, sp->inj_fluid_conc, sp->I, sp->well_rc, sp->tstart, sp->tfinal,
                        (double)sp->stype, (double)sp->avgproc, (double)sp->dim_splitting, (double)sp->diffusion, (double)sp->ptype, (double)sp->is_test,
                        (double)sp->well_profile, (double)sp->hmtype};
    memcpy(out, v, sizeof(v));
}
int ref_nparams(void) { return 29; }
void ref_set_threads(int n) { omp_set_num_threads(n); }

#define SCALAR_IO(NAME, ARR)                                                                                  \
    void ref_set_##NAME(void *h, const double *a) { Ref *R = (Ref *)h; const int g = R->sp->c_bounds, W = R->sp->N2 + 2 * g;      \
        for (int i = -g; i < R->sp->N1 + g; i++) for (int j = -g; j < R->sp->N2 + g; j++) R->ARR(i, j) = a[(size_t)(i + g) * W + (j + g)]; } \
    void ref_get_##NAME(void *h, double *a) { Ref *R = (Ref *)h; const int g = R->sp->c_bounds, W = R->sp->N2 + 2 * g;            \
        for (int i = -g; i < R->sp->N1 + g; i++) for (int j = -g; j < R->sp->N2 + g; j++) a[(size_t)(i + g) * W + (j + g)] = R->ARR(i, j); }
SCALAR_IO(c, c)
SCALAR_IO(p, p)
void ref_set_v(void *h, const double *a) {
    Ref *R = (Ref *)h; const int g = R->sp->v_bounds, W = R->sp->N2 + 1 + 2 * g;
    for (int i = -g; i < R->sp->N1 + 1 + g; i++) for (int j = -g; j < R->sp->N2 + 1 + g; j++) for (int k = 0; k < 2; k++) R->v(i, j, k) = a[((size_t)(i + g) * W + (j + g)) * 2 + k];
}
void ref_get_v(void *h, double *a) {
    Ref *R = (Ref *)h; const int g = R->sp->v_bounds, W = R->sp->N2 + 1 + 2 * g;
    for (int i = -g; i < R->sp->N1 + 1 + g; i++) for (int j = -g; j < R->sp->N2 + 1 + g; j++) for (int k = 0; k < 2; k++) a[((size_t)(i + g) * W + (j + g)) * 2 + k] = R->v(i, j, k);
}
void ref_set_q(void *h, const double *q) { Ref *R = (Ref *)h; memcpy(R->q, q, sizeof(double) * (size_t)R->sp->N1 * R->sp->N2); }
void ref_get_q(void *h, double *q) { Ref *R = (Ref *)h; memcpy(q, R->q, sizeof(double) * (size_t)R->sp->N1 * R->sp->N2); }

// the dispatch of Impec2DSolver::InitializeSource (src/Impec2DSolver.cpp:241-280) over the reference's Init* functions (src/Util.cpp:444-639)
int ref_init_source(void *h, double t) {
    Ref *R = (Ref *)h; SimParams *sp = R->sp;
    const double Inj_Rate = sp->I, r = sp->well_rc * (sp->x_ub - sp->x_lb);
    const double circval = Inj_Rate * 4. / (M_PI * r * r), qval = Inj_Rate / (sp->dx * sp->dy);
    memset(R->q, 0, sizeof(double) * (size_t)sp->N1 * sp->N2);
    if (sp->stype == PARALLEL) {
        if (sp->well_profile == CIRCULAR) InitQuarterCircleParallel(sp, R->q, r, circval);
        else if (sp->well_profile == ONECELL) InitOneCellParallel(sp, R->q, qval);
        else return 1;
    } else if (sp->stype == DIAGONAL) {
        if (sp->well_profile == CIRCULAR) InitQuarterCircleDiagonal(sp, R->q, r, circval);
        else if (sp->well_profile == ONECELL) InitOneCellDiagonal(sp, R->q, qval);
        else return 1;
    } else if (sp->stype == HORIZONTAL) InitHorizontalFront(sp, R->q, t);
    else return 1;
    return 0;
}

// ---- a1
void ref_fill_boundaries(void *h) { Ref *R = (Ref *)h; R->hs->FillBoundaries(R->c); }
void ref_fill_neumann_velocity(void *h) { Ref *R = (Ref *)h; FillGhostCellsNeumannVelocity(R->v, R->sp->v_bounds); }
// ---- a2-a4
double ref_mu(void *h, double c) { return ((Ref *)h)->es->mu(c); }
double ref_face_mob_x(void *h, double c_im1j, double c_ij, int i, int j) { return ((Ref *)h)->es->KoverMu_11_im12j(c_im1j, c_ij, i, j); }
double ref_face_mob_y(void *h, double c_ijm1, double c_ij, int i, int j) { return ((Ref *)h)->es->KoverMu_22_ijm12(c_ijm1, c_ij, i, j); }
double ref_vertex_mob_11(void *h, int i, int j, double c_ij, double c_ip1j, double c_ijp1, double c_ip1jp1) { return ((Ref *)h)->es->KoverMu_11_ip12jp12(i, j, c_ij, c_ip1j, c_ijp1, c_ip1jp1); }
double ref_vertex_mob_22(void *h, int i, int j, double c_ij, double c_ip1j, double c_ijp1, double c_ip1jp1) { return ((Ref *)h)->es->KoverMu_22_ip12jp12(i, j, c_ij, c_ip1j, c_ijp1, c_ip1jp1); }
// ---- a5-a7: the 9 weighted coefficients of one cell in the order W,NW,N,NE,E,SE,S,SW,C (before Neumann folding)
void ref_cell_coefficients(void *h, double nu, int i, int j, double *out9) {
    Ref *R = (Ref *)h;
    R->es->GetWeightedMatrixCoefficientsForCell(nu, i, j, out9[0], out9[1], out9[2], out9[3], out9[4], out9[5], out9[6], out9[7], out9[8], R->c);
}
// ---- a8-a11: EllipticSolver::Solve (assembly with folding and pin, right-hand side, solve, copy-out); the assembled system is
// available afterwards through refhypre_last()
int ref_pressure_solve(void *h) { Ref *R = (Ref *)h; R->es->Solve(R->p, R->c, R->q); return R->es->GetLastIts(); }
// ---- a12
void ref_calc_velocities(void *h) { Ref *R = (Ref *)h; R->es->CalcVelocities(R->p, R->v, R->c); }
// ---- a13
double ref_max_stable_dt(void *h) { Ref *R = (Ref *)h; return R->hs->MaxStableTimestep(R->c, R->v); }
double ref_max_stable_dt_combined(void *h) { Ref *R = (Ref *)h; return R->hs->MaxStableTimestepCombined(R->c, R->v); }
double ref_max_stable_dt_advection(void *h) { Ref *R = (Ref *)h; return R->hs->MaxStableAdvectionTimestep(R->c, R->v); }
// ---- a14-a18 (method 0 = MUSCLSplitSolver, 1 = UpwindSplitSolver)
int ref_transport_step(void *h, double dt, int method) {
    Ref *R = (Ref *)h;
    return method == 1 ? (int)R->up->Solve(R->c, R->v, R->q, dt) : (int)R->hs->Solve(R->c, R->v, R->q, dt);
}
int ref_advect_split(void *h, double dt, int method) {
    Ref *R = (Ref *)h;
    return method == 1 ? (int)R->up->AdvectSplit(R->c, R->v, R->q, dt) : (int)R->hs->AdvectSplit(R->c, R->v, R->q, dt);
}
int ref_advect_x(void *h, double dt, int method) { Ref *R = (Ref *)h; return (int)pick(R, method)->Advect_X(R->c, R->v, dt); }
int ref_advect_y(void *h, double dt, int method) { Ref *R = (Ref *)h; return (int)pick(R, method)->Advect_Y(R->c, R->v, dt); }
int ref_source_split(void *h, double dt, int method) {
    Ref *R = (Ref *)h;
    return method == 1 ? (int)R->up->SourceSplit(R->c, R->q, dt) : (int)R->hs->SourceSplit(R->c, R->q, dt);
}
double ref_total_mass_outflux(void *h, int method) { Ref *R = (Ref *)h; return pick(R, method)->GetTotalMassOutflux(); }
// intercell fluxes F (x faces: (N1+1) x (N2+2g)) and G (y faces: (N1+2g) x (N2+1)), flattened in C order
void ref_flux_F(void *h, double dt, int method, double *out) {
    Ref *R = (Ref *)h; SimParams *sp = R->sp; const int g = sp->c_bounds;
    Array2d F;
    F.resize(sp->N1 + 1, sp->N2 + g * 2);
    Idx2d idx; idx[0] = 0; idx[1] = -g; F.reindexSelf(idx);           // as Advect_X does (src/HyperbolicSplitSolver.cpp:114-120)
    pick(R, method)->CalculateIntercellFlux_F(F, R->c, R->v, dt);
    for (int i = 0; i <= sp->N1; i++) for (int j = -g; j < sp->N2 + g; j++) out[(size_t)i * (sp->N2 + 2 * g) + (j + g)] = F(i, j);
}
void ref_flux_G(void *h, double dt, int method, double *out) {
    Ref *R = (Ref *)h; SimParams *sp = R->sp; const int g = sp->c_bounds;
    Array2d G;
    G.resize(sp->N1 + g * 2, sp->N2 + 1);
    Idx2d idx; idx[0] = -g; idx[1] = 0; G.reindexSelf(idx);           // as Advect_Y does (:140-145)
    pick(R, method)->CalculateIntercellFlux_G(G, R->c, R->v, dt);
    for (int i = -g; i < sp->N1 + g; i++) for (int j = 0; j <= sp->N2; j++) out[(size_t)(i + g) * (sp->N2 + 1) + j] = G(i, j);
}
// ---- a19
int ref_diffusion_split(void *h, double dt) { Ref *R = (Ref *)h; R->hs->DiffusionSplit(R->c, R->v, dt); return R->es->GetLastParabolicIts(); }
// ---- a20: the body of the while loop of Impec2DSolver::Solve (src/Impec2DSolver.cpp:343-389) on the reference's own objects
double ref_impes_step(void *h, double t, double t_next_output, int *its) {
    Ref *R = (Ref *)h;
    R->hs->FillBoundaries(R->c);
    R->es->Solve(R->p, R->c, R->q);
    R->es->CalcVelocities(R->p, R->v, R->c);
    double dt = R->hs->MaxStableTimestep(R->c, R->v);
    if (t_next_output > t && t + dt >= t_next_output) dt = t_next_output - t;
    R->hs->Solve(R->c, R->v, R->q, dt);
    if (its) *its = R->es->GetLastIts();
    return dt;
}
// ---- test profiles (src/Util.cpp:55-114) and their norms (:117-147)
double ref_profile(void *h, int which, double x, double y) {
    SimParams *sp = ((Ref *)h)->sp;
    switch (which) {
    case 0: return Sinewave(sp, x, y);
    case 1: return Tophat(sp, x, y);
    case 2: return Gaussbell(sp, x, y);
    case 3: return HorizontalSine(sp, x, y);
    case 4: return VerticalSine(sp, x, y);
    case 5: return DiagonalSinewave(sp, x, y);
    default: return 0.;
    }
}
double ref_exact_vol_frac_circle(double x1, double y1, double r, double dx, double dy) { return ExactVolFracCircle(x1, y1, r, dx, dy); }
// ---- 8f-3: ParallelProjector on the concentration (field 0) or pressure (1); out: dN1 x dN2 (src/Impec2DSolver.cpp:690-716)
void ref_project(void *h, int field, int dN1, int dN2, double *out) {
    Ref *R = (Ref *)h; SimParams *sp = R->sp;
    Array2d arrd;
    const int bounds = sp->c_bounds;
    arrd.resize(dN1 + 2 * bounds, dN2 + 2 * bounds);
    Idx2d idx; idx[0] = -bounds; idx[1] = -bounds; arrd.reindexSelf(idx);
    arrd = 0.;
    ParallelProjector pp(sp->x_lb, sp->y_lb, sp->x_ub, sp->y_ub, dN1, dN2, field == 0 ? &R->c : &R->p, bounds);
    pp.FillDiagonal(arrd, bounds);
    for (int i = 0; i < dN1; i++) for (int j = 0; j < dN2; j++) out[(size_t)i * dN2 + j] = arrd(i, j);
}

}  // extern "C"
