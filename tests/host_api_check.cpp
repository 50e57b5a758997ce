// host_api_check.cpp -- drives the C++ class surface of impec2d_b200/host directly (no driver executable): the calls a maintainer of the
// reference makes on EllipticSolver / HyperbolicSplitSolver objects.  Built by tests/test_host_layer.py against the CPU-emulated kernel
// library and by tests/test_gpu_parity.py against the product library (libimpec2d_host.so + libimpec2d_b200.so) on the B200.
//   usage: host_api_check <parallel.cfg variant>
#include "SimParams.h"
#include "GenSimDefs.h"
#include "EllipticSolver.h"
#include "HyperbolicSolver.h"
#include "DeviceContext.h"
#include "Util.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

static int fails = 0;
#define CHECK(cond, ...) do { if (!(cond)) { fails++; std::printf("FAIL %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); } } while (0)

static double noise(unsigned &s) { s = s * 1664525u + 1013904223u; return (double)(s >> 8) / 16777216.0; }

template <class Solver>
static void check_fluxes(SimParams *sp, const char *name) {
    const int g = sp->c_bounds, N1 = sp->N1, N2 = sp->N2;
    Array2d c; Idx2d i2; i2[0] = -g; i2[1] = -g;
    c.resize(N1 + 2 * g, N2 + 2 * g); c.reindexSelf(i2);
    Array3d v; Idx3d i3; i3[0] = -sp->v_bounds; i3[1] = -sp->v_bounds; i3[2] = 0;
    v.resize(N1 + 1 + 2 * sp->v_bounds, N2 + 1 + 2 * sp->v_bounds, 2); v.reindexSelf(i3);
    unsigned seed = 12345u;
    for (int i = 0; i < N1; i++)
        for (int j = 0; j < N2; j++) c(i, j) = 0.5 + 0.4 * std::sin(0.37 * i) * std::cos(0.23 * j) + 0.05 * noise(seed);
    for (int i = -sp->v_bounds; i < N1 + 1 + sp->v_bounds; i++)
        for (int j = -sp->v_bounds; j < N2 + 1 + sp->v_bounds; j++) { v(i, j, 0) = 0.7 - 1.1 * noise(seed); v(i, j, 1) = -0.4 + 0.9 * noise(seed); }
    Solver hs(sp);
    hs.FillBoundaries(c);
    const double dt = 0.3 * sp->dx;
    // CalculateIntercellFlux_F / Advect_X (src/HyperbolicSplitSolver.cpp:113-133): c += dt/dx (F(i,j) - F(i+1,j))
    Array2d F, G;
    hs.CalculateIntercellFlux_F(F, c, v, dt);
    CHECK(F.extent(0) == N1 + 1 && F.extent(1) == N2 + 2 * g && F.base()[0] == 0 && F.base()[1] == -g, "%s: shape of F", name);
    Array2d cx = c.copy();
    hs.Advect_X(cx, v, dt);
    double worst = 0., fmax = 0.;
    for (int i = 0; i < N1; i++)
        for (int j = -g; j < N2 + g; j++) {
            const double expect = c(i, j) + dt / sp->dx * (F(i, j) - F(i + 1, j));
            worst = std::fmax(worst, std::fabs(expect - cx(i, j))); fmax = std::fmax(fmax, std::fabs(F(i, j)));
        }
    CHECK(worst == 0.0, "%s: Advect_X differs from c + dt/dx (F_i - F_i+1) by %.3e", name, worst);
    CHECK(fmax > 0.0, "%s: F is identically zero", name);
    hs.CalculateIntercellFlux_G(G, c, v, dt);
    CHECK(G.extent(0) == N1 + 2 * g && G.extent(1) == N2 + 1 && G.base()[0] == -g && G.base()[1] == 0, "%s: shape of G", name);
    Array2d cy = c.copy();
    hs.Advect_Y(cy, v, dt);
    worst = 0.;
    for (int i = -g; i < N1 + g; i++)
        for (int j = 0; j < N2; j++) worst = std::fmax(worst, std::fabs(c(i, j) + dt / sp->dy * (G(i, j) - G(i, j + 1)) - cy(i, j)));
    CHECK(worst == 0.0, "%s: Advect_Y differs from c + dt/dy (G_j - G_j+1) by %.3e", name, worst);
    std::printf("%s: fluxes consistent with the sweeps (max |F| %.4f)\n", name, fmax);
}

// EllipticSolver::Solve + CalcVelocities + MaxStableTimestep + HyperbolicSolver::Solve, as Impec2DSolver::Solve's loop body calls them
static void check_step(SimParams *sp) {
    const int g = sp->c_bounds, N1 = sp->N1, N2 = sp->N2;
    Array2d c, p; Idx2d i2; i2[0] = -g; i2[1] = -g;
    c.resize(N1 + 2 * g, N2 + 2 * g); c.reindexSelf(i2);
    p.resize(N1 + 2 * g, N2 + 2 * g); p.reindexSelf(i2);
    Array3d v; Idx3d i3; i3[0] = -sp->v_bounds; i3[1] = -sp->v_bounds; i3[2] = 0;
    v.resize(N1 + 1 + 2 * sp->v_bounds, N2 + 1 + 2 * sp->v_bounds, 2); v.reindexSelf(i3);
    std::vector<double> q((size_t)N1 * N2, 0.0);
    InitSource(sp, q.data());
    MGSolver es(sp);
    MUSCLSplitSolver hs(sp);
    hs.AssignMGSolver(&es);
    double mass0 = 0.;
    for (int s = 0; s < 2; s++) {
        hs.FillBoundaries(c);
        CHECK(es.Solve(p, c, q.data()), "EllipticSolver::Solve failed");
        es.CalcVelocities(p, v, c);
        const double dt = hs.MaxStableTimestep(c, v);
        CHECK(dt > 0. && std::isfinite(dt), "MaxStableTimestep returned %g", dt);
        HyperbolicSolver *h = &hs;                              // (the reference calls Solve through the base class: Impec2DSolver::hs is a HyperbolicSolver*)
        CHECK(h->Solve(c, v, q.data(), dt), "HyperbolicSolver::Solve failed");
        double mass = 0., cmin = 1e300, cmax = -1e300;
        for (int i = 0; i < N1; i++)
            for (int j = 0; j < N2; j++) { mass += c(i, j) * sp->dx * sp->dy; cmin = std::fmin(cmin, c(i, j)); cmax = std::fmax(cmax, c(i, j)); }
        std::printf("step %d: its %d relres %.2e dt %.4e mass %.6e c in [%.3e, %.3e]\n", s + 1, es.GetLastIts(), es.GetLastRelativeResidual(), dt, mass, cmin, cmax);
        CHECK(es.GetLastIts() > 0 && es.GetLastRelativeResidual() < sp->eps * 10, "pressure solve did not converge");
        CHECK(mass > mass0 && cmin > -1e-12 && cmax < 1. + 1e-12, "concentration out of range or mass not growing");
        mass0 = mass;
    }
}

int main(int argc, char **argv) {
    if (argc < 2) { std::printf("usage: host_api_check <cfg>\n"); return 2; }
    SimParams sp(argv[1]);
    check_fluxes<MUSCLSplitSolver>(&sp, "MUSCLSplitSolver");
    check_fluxes<UpwindSplitSolver>(&sp, "UpwindSplitSolver");
    check_step(&sp);
    DeviceContext::Release(&sp);
    if (fails) { std::printf("host_api_check: %d failure(s)\n", fails); return 1; }
    std::printf("host_api_check ok\n");
    return 0;
}
