// Solvers.cpp -- EllipticSolver/MGSolver, HyperbolicSolver family, TRIntegrator: thin forwards to the C ABI.
// Every method keeps the reference's argument meaning; arrays are made current on the device lazily
// (DeviceContext::Use) and marked device-newer after an operation that writes them.
#include "EllipticSolver.h"
#include "HyperbolicSolver.h"
#include "TimeIntegrator.h"
#include "DeviceContext.h"
#include "Util.h"
#include <iostream>
#include <fstream>
#include <cmath>
#include <cstdlib>
#include <stdexcept>
using namespace std;

// ------------------------------------------------------------------ EllipticSolver (src/EllipticSolver.cpp:59-142)
EllipticSolver::EllipticSolver(SimParams *sp) : sp(sp) { SetNu(sp->nu); last_parabolic_its = 0; last_relres = last_true_relres = 0; Initialize(); }
EllipticSolver::~EllipticSolver() {}
void EllipticSolver::Initialize() { last_its = 0; total_its = 0; calls = 0; }
void EllipticSolver::SetNu(double nu_) {
    if (nu_ >= 0. && nu_ <= 1.) { nu = nu_; sp->nu = nu_; }
    else { cerr << "EllipticSolver::SetNu> Invalid value for nu: " << nu_ << endl; exit(1); }
}
bool EllipticSolver::Solve(Array2d &p, Array2d &c, double *q) {
    calls++;
    const bool ret = SolveLinearMultigrid(c, sp->c_bounds, p, sp->p_bounds, q);
    total_its += last_its;
    return ret;
}
// src/EllipticSolver.cpp:660-683
void EllipticSolver::CalcVelocities(Array2d &p, Array3d &v, Array2d &c) {
    impec_ctx *ctx = DeviceContext::Get(sp);
    DeviceContext::Use(ctx, &c, &p, nullptr, nullptr, sp);
    if (!v.BoundTo(ctx)) v.Bind(ctx, FIELD_V);
    DeviceContext::Check(ctx, impec_calc_velocities(ctx), "impec_calc_velocities");
    v.MarkDeviceNewer(); p.MarkDeviceNewer();          // p's ghost ring is filled on the device (:663)
}

// ------------------------------------------------------------------ MGSolver (src/MGSolver.cpp:50-149, 223-248, 952-976)
MGSolver::MGSolver(SimParams *sp) : EllipticSolver(sp) {
    // Dirichlet cells (src/MGSolver.cpp:109-140): cell (1,1) for parallel / diagonal wells, the rows i = 1 (pressure) and i = N1-2 (Crank-
    // Nicolson) for the horizontal front -- both handled inside the device library from sp->stype
    if (sp->stype != PARALLEL && sp->stype != DIAGONAL && sp->stype != HORIZONTAL) { cerr << "ERROR: Unknown source type!" << endl; exit(1); }
}
MGSolver::~MGSolver() {}
bool MGSolver::SolveLinearMultigrid(Array2d &c_c, int c_bounds, Array2d &p_c, int p_bounds, double const *const q_v) {
    return Solve_Elliptic_cc_to_c(c_c, c_bounds, p_c, p_bounds, q_v);
}
bool MGSolver::Solve_Elliptic_cc_to_c(const Array2d &c_c, int c_bounds, Array2d &p_c, int p_bounds, double const *const q_c) {
    if (c_bounds <= 0) { cout << "MGSolver::Error> At least one layer of ghost cells is required for the concentration aray!" << endl; return false; }
    (void)p_bounds;
    impec_ctx *ctx = DeviceContext::Get(sp);
    DeviceContext::Use(ctx, &c_c, nullptr, nullptr, q_c, sp);
    if (!p_c.BoundTo(ctx)) p_c.Bind(ctx, FIELD_P);
    impec_solve_info info;
    const int rc = impec_pressure_solve(ctx, sp->eps, &info);
    last_its = info.iterations; last_relres = info.relres; last_true_relres = info.true_relres;
    p_c.MarkDeviceNewer();
    DeviceContext::Check(ctx, rc, "impec_pressure_solve");   // IMPEC_WARN_NOCONV is only reported: upstream never detects non-convergence
    return true;
}
bool MGSolver::Solve_Parabolic_cc_to_c(Array2d &c_c, int c_bounds, const Array3d &v, int v_bounds, Array2d const *const ls, int ls_bounds, const double dt) {
    (void)c_bounds; (void)v_bounds; (void)ls; (void)ls_bounds;
    impec_ctx *ctx = DeviceContext::Get(sp);
    DeviceContext::Use(ctx, &c_c, nullptr, &v, nullptr, sp);
    impec_solve_info info;
    DeviceContext::Check(ctx, impec_diffusion_split(ctx, dt, &info), "impec_diffusion_split");
    last_parabolic_its = info.iterations;
    c_c.MarkDeviceNewer();
    return true;
}

// manufactured-solution test p = cos x cos y (src/EllipticSolver.cpp:155-470): right-hand sides and exact fields
static double K11t(int testno, double x, double y) { switch (testno) { case 2: return exp(-0.3 * x); case 3: case 4: return 2. + sin(2. * x) * sin(y); default: return 1.; } }
static double K22t(int testno, double x, double y) { switch (testno) { case 2: return exp(-0.3 * y); case 3: return 2. + sin(2. * x) * sin(y); case 4: return 2. + sin(x) * sin(2. * y); default: return 1.; } }
static double qTest(int testno, double mu_res, double x, double y) {
    const double t = 2.;
    double a1, a2, a1_x, a2_y;
    switch (testno) {
    case 1: return 2. / mu_res * cos(x) * cos(y);
    case 2: a1 = exp(-0.3 * x); a2 = exp(-0.3 * y); a1_x = -0.3 * exp(-0.3 * x); a2_y = -0.3 * exp(-0.3 * y); break;
    case 3: a1 = a2 = 2. + sin(t * x) * sin(y); a1_x = t * cos(t * x) * sin(y); a2_y = cos(y) * sin(t * x); break;
    case 4: a1 = 2. + sin(t * x) * sin(y); a2 = 2. + sin(x) * sin(t * y); a1_x = t * cos(t * x) * sin(y); a2_y = t * cos(t * y) * sin(x); break;
    default: cerr << "EllipticSolver::TestNeumann::Initialization> Unknown test no.!" << endl; exit(1);
    }
    return 1. / mu_res * ((a1 + a2) * cos(x) * cos(y) + a1_x * sin(x) * cos(y) + a2_y * cos(x) * sin(y));
}

bool EllipticSolver::TestNeumann(GnuPlotter *gp) {
    (void)gp;
    sp->btype = NEUMANN;
    bool stat = true;
    sp->N1 = sp->init_x_cells; sp->N2 = sp->init_y_cells;
    // Several processes (one slab each): every rank runs the test, the norms are taken over the gathered fields, rank 0 writes the files.
    const bool root = DeviceContext::Rank() == 0, slabs = DeviceContext::NumRanks() > 1;
    ofstream norm_stream;
    if (root) {
        norm_stream.open(sp->norm_file.c_str());
        if (!norm_stream.is_open()) { cerr << "Cannot write norm-file! Convergence test aborted..." << endl; return false; }
        norm_stream.precision(30);
    }
    for (int k = 0; k < sp->levels; k++) {
        if (root) cout << "-> (N1,N2):\t(" << sp->N1 << " , " << sp->N2 << ")";
        sp->dx = (sp->x_ub - sp->x_lb) / sp->N1; sp->dy = (sp->y_ub - sp->y_lb) / sp->N2; sp->M = sp->N1 * sp->N2;
        EllipticSolver *mg_test = new MGSolver(sp);
        const int g = sp->p_bounds, N1 = sp->N1, N2 = sp->N2;
        Array2d p, p_ex, c; Array3d v, v_ex;
        Idx2d idx; idx[0] = -g; idx[1] = -g;
        p.resize(N1 + 2 * g, N2 + 2 * g); p.reindexSelf(idx);
        p_ex.resize(N1 + 2 * g, N2 + 2 * g); p_ex.reindexSelf(idx);
        c.resize(N1 + 2 * g, N2 + 2 * g); c.reindexSelf(idx);
        Idx3d i3; i3[0] = -sp->v_bounds; i3[1] = -sp->v_bounds; i3[2] = 0;
        v.resize(N1 + 1 + 2 * sp->v_bounds, N2 + 1 + 2 * sp->v_bounds, 2); v.reindexSelf(i3);
        v_ex.resize(N1 + 1 + 2 * sp->v_bounds, N2 + 1 + 2 * sp->v_bounds, 2); v_ex.reindexSelf(i3);
        double *q = new double[sp->M];
        for (int i = 0; i <= N1; i++) for (int j = 0; j < N2; j++) { const double x = sp->x_lb + i * sp->dx, y = sp->y_lb + (j + 0.5) * sp->dy; v_ex(i, j, 0) = K11t(sp->testno, x, y) / sp->mu_res * sin(x) * cos(y); }
        for (int i = 0; i < N1; i++) for (int j = 0; j <= N2; j++) { const double x = sp->x_lb + (i + 0.5) * sp->dx, y = sp->y_lb + j * sp->dy; v_ex(i, j, 1) = K22t(sp->testno, x, y) / sp->mu_res * cos(x) * sin(y); }
        for (int i = 0; i < N1; i++) for (int j = 0; j < N2; j++) {
            const double x = sp->x_lb + (i + 0.5) * sp->dx, y = sp->y_lb + (j + 0.5) * sp->dy;
            p_ex(i, j) = cos(x) * cos(y); q[i * N2 + j] = qTest(sp->testno, sp->mu_res, x, y);
        }
        // c = 0 everywhere, so its Neumann ghost fill (:231) is a no-op
        stat = stat && mg_test->Solve(p, c, q);
        if (root) cout << "\tITS:\t" << mg_test->GetLastIts() << endl;
        mg_test->CalcVelocities(p, v, c);
        if (slabs) DeviceContext::GatherGlobal(DeviceContext::Get(sp), NULL, &p, &v);
        double p_sum = 0, u_sum = 0, v_sum = 0;
        for (int i = 0; i < N1; i++) for (int j = 0; j < N2; j++) p_sum += fabs(p(i, j) - p_ex(i, j));
        for (int i = 0; i <= N1; i++) for (int j = 0; j < N2; j++) u_sum += fabs(v_ex(i, j, 0) - v(i, j, 0));
        for (int i = 0; i < N1; i++) for (int j = 0; j <= N2; j++) v_sum += fabs(v_ex(i, j, 1) - v(i, j, 1));
        p_sum *= sp->dx * sp->dy; u_sum *= sp->dx * sp->dy; v_sum *= sp->dx * sp->dy;
        if (k == 0 && root) {
            PrintDToPlotFile(sp->num_result_file, p, sp->x_lb, sp->y_lb, sp->dx, sp->dy, sp->p_bounds);
            PrintDToPlotFile(sp->ex_result_file, p_ex, sp->x_lb, sp->y_lb, sp->dx, sp->dy, sp->p_bounds);
        }
        if (root) norm_stream << log(sp->dx) << "\t" << log(p_sum) << "\t" << log(u_sum) << "\t" << log(v_sum) << endl;
        p.Unbind(); c.Unbind(); v.Unbind();
        delete[] q; delete mg_test;
        sp->N1 *= 2; sp->N2 *= 2;
    }
    if (root) norm_stream.close();
    return stat;
}

// ------------------------------------------------------------------ TRIntegrator (src/TRIntegrator.cpp:49-71)
int TRIntegrator::Integrate(Array2d &c, double const *const q, const double dt) {
    impec_ctx *ctx = DeviceContext::Get(sp);
    DeviceContext::Use(ctx, &c, nullptr, nullptr, q, sp);
    impec_diag d0, d1;
    DeviceContext::Check(ctx, impec_diagnostics(ctx, &d0), "impec_diagnostics");
    DeviceContext::Check(ctx, impec_source_split(ctx, dt), "impec_source_split");
    DeviceContext::Check(ctx, impec_diagnostics(ctx, &d1), "impec_diagnostics");
    total_mass_outflux = d1.total_mass_outflux - d0.total_mass_outflux;     // |sum of mass differences| of this call
    c.MarkDeviceNewer();
    return 0;
}

// ------------------------------------------------------------------ HyperbolicSolver (src/HyperbolicSolver.cpp:56-223)
HyperbolicSolver::HyperbolicSolver(SimParams *sp) : sp(sp) { mg = NULL; total_mass_outflux = 0; prev_first_sweep = Y_SWEEP; ti = new TRIntegrator(sp); }
HyperbolicSolver::~HyperbolicSolver() { delete ti; }
void HyperbolicSolver::FillBoundaries(Array2d &c) {
    impec_ctx *ctx = DeviceContext::Get(sp);
    DeviceContext::Use(ctx, &c, nullptr, nullptr, nullptr, sp);
    DeviceContext::Check(ctx, impec_fill_boundaries(ctx), "impec_fill_boundaries");
    c.MarkDeviceNewer();
}
double HyperbolicSolver::GetTotalMassOutflux() {
    impec_ctx *ctx = DeviceContext::Get(sp);
    impec_diag d; DeviceContext::Check(ctx, impec_diagnostics(ctx, &d), "impec_diagnostics");
    return d.total_mass_outflux;
}
double HyperbolicSolver::MaxStableTimestep(const Array2d &c, const Array3d &v) {
    impec_ctx *ctx = DeviceContext::Get(sp);
    DeviceContext::Use(ctx, &c, nullptr, &v, nullptr, sp);
    double dt = 0; DeviceContext::Check(ctx, impec_max_stable_dt(ctx, &dt), "impec_max_stable_dt");
    return dt;
}
double HyperbolicSolver::MaxStableAdvectionTimestep(const Array2d &c, const Array3d &v) { return MaxStableTimestep(c, v); }   // is_test selects the rule (:93-98)
double HyperbolicSolver::MaxStableTimestepCombined(const Array2d &c, const Array3d &v) { return MaxStableTimestep(c, v); }
bool HyperbolicSolver::SourceSplit(Array2d &c, double const *const q, const double dt) {
    impec_ctx *ctx = DeviceContext::Get(sp);
    DeviceContext::Use(ctx, &c, nullptr, nullptr, q, sp);
    DeviceContext::Check(ctx, impec_source_split(ctx, dt), "impec_source_split");     // accumulates |mass outflux| on the device
    c.MarkDeviceNewer();
    return true;
}
bool HyperbolicSolver::DiffusionSplit(Array2d &c, const Array3d &v, const double dt) { return DiffusionSplitCN(c, v, dt); }
bool HyperbolicSolver::DiffusionSplitCN(Array2d &c, const Array3d &v, const double dt) {
    if (mg == NULL) { cout << "HyperbolicSolver:: Error> A Multigrid-Solver is needed for Crank-Nicolson!" << endl; exit(1); }
    return mg->Solve_Parabolic_cc_to_c(c, sp->c_bounds, v, sp->v_bounds, NULL, sp->ls_bounds, dt);
}

// ------------------------------------------------------------------ HyperbolicSplitSolver (src/HyperbolicSplitSolver.cpp:51-198)
bool HyperbolicSplitSolver::Advect_X(Array2d &c, const Array3d &v, const double dt) {
    impec_ctx *ctx = DeviceContext::Get(sp);
    DeviceContext::SetMethod(ctx, MethodId());
    DeviceContext::Use(ctx, &c, nullptr, &v, nullptr, sp);
    DeviceContext::Check(ctx, impec_advect_x(ctx, dt), "impec_advect_x");
    c.MarkDeviceNewer();
    return true;
}
bool HyperbolicSplitSolver::Advect_Y(Array2d &c, const Array3d &v, const double dt) {
    impec_ctx *ctx = DeviceContext::Get(sp);
    DeviceContext::SetMethod(ctx, MethodId());
    DeviceContext::Use(ctx, &c, nullptr, &v, nullptr, sp);
    DeviceContext::Check(ctx, impec_advect_y(ctx, dt), "impec_advect_y");
    c.MarkDeviceNewer();
    return true;
}
bool HyperbolicSplitSolver::AdvectSplit(Array2d &c, const Array3d &v, double const *const q, const double dt) {
    (void)q;
    impec_ctx *ctx = DeviceContext::Get(sp);
    DeviceContext::SetMethod(ctx, MethodId());
    DeviceContext::Use(ctx, &c, nullptr, &v, nullptr, sp);
    DeviceContext::Check(ctx, impec_advect_split(ctx, dt), "impec_advect_split");
    c.MarkDeviceNewer();
    return true;
}
// S(dt/2) [D(dt/2)] A(dt) S(dt/2) [D(dt/2)] in one call (src/HyperbolicSplitSolver.cpp:162-198)
bool HyperbolicSplitSolver::Solve(Array2d &c, const Array3d &v, double const *const q, const double dt) {
    impec_ctx *ctx = DeviceContext::Get(sp);
    DeviceContext::SetMethod(ctx, MethodId());
    DeviceContext::Use(ctx, &c, nullptr, &v, q, sp);
    DeviceContext::Check(ctx, impec_transport_step(ctx, dt), "impec_transport_step");
    if (sp->diffusion) { impec_solve_info pi; if (impec_last_parabolic_info(ctx, &pi) == IMPEC_OK) ReportParabolicIts(pi.iterations); }
    c.MarkDeviceNewer();
    return true;
}
// CalculateIntercellFlux_F / _G (src/MUSCLSplitSolver.cpp:102-189, src/UpwindSplitSolver.cpp:43-61): the sweeps evaluate their fluxes on the
// fly, so the flux FIELD is produced on request by its own kernel (impec_intercell_flux) and copied into the caller's array, which
// keeps the reference's shape and index base (F: (N1+1) x (N2+2g) from (0,-g); G: (N1+2g) x (N2+1) from (-g,0)).
static void intercell_flux(SimParams *sp, int method, int dir, Array2d &F, const Array2d &c, const Array3d &v, const double dt) {
    impec_ctx *ctx = DeviceContext::Get(sp);
    DeviceContext::SetMethod(ctx, method);
    DeviceContext::Use(ctx, &c, nullptr, &v, nullptr, sp);
    const int g = sp->c_bounds;
    const int rows = dir == 0 ? sp->N1 + 1 : sp->N1 + 2 * g, cols = dir == 0 ? sp->N2 + 2 * g : sp->N2 + 1;
    const int bi = dir == 0 ? 0 : -g, bj = dir == 0 ? -g : 0;
    if (F.extent(0) != rows || F.extent(1) != cols) { F.resize(rows, cols); Idx2d idx; idx[0] = bi; idx[1] = bj; F.reindexSelf(idx); }
    DeviceContext::Check(ctx, impec_intercell_flux(ctx, dir, dt, F.data()), "impec_intercell_flux");      // row-major, same shape
}
void HyperbolicSplitSolver::CalculateIntercellFlux_F(Array2d &F, const Array2d &c, const Array3d &v, const double dt) { intercell_flux(sp, MethodId(), 0, F, c, v, dt); }
void HyperbolicSplitSolver::CalculateIntercellFlux_G(Array2d &G, const Array2d &c, const Array3d &v, const double dt) { intercell_flux(sp, MethodId(), 1, G, c, v, dt); }
MUSCLSplitSolver::MUSCLSplitSolver(SimParams *sp) : HyperbolicSplitSolver(sp) {}
int MUSCLSplitSolver::MethodId() const { return IMPEC_MUSCL; }
UpwindSplitSolver::UpwindSplitSolver(SimParams *sp) : HyperbolicSplitSolver(sp) {}
int UpwindSplitSolver::MethodId() const { return IMPEC_UPWIND; }
