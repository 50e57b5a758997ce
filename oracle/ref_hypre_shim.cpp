// ref_hypre_shim.cpp -- TEST INFRASTRUCTURE ONLY (part of oracle/_ref/libimpec_ref.so, never of the product).
// Implementation of oracle/ref_shim/HYPRE_struct_ls.h: records the struct-interface calls of the reference's src/MGSolver.cpp and
// solves the recorded 9-point system with a small Krylov solver of its own.  See the header for what is and is not reproduced.
#include "ref_shim/HYPRE_struct_ls.h"
#include <vector>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <algorithm>

struct refhypre_grid { int ilo[2], ihi[2]; };
struct refhypre_stencil { int n; int off[16][2]; };
struct refhypre_matrix { refhypre_grid g; refhypre_stencil s; std::vector<double> v; };
struct refhypre_vector { refhypre_grid g; std::vector<double> v; };
struct refhypre_solver { double tol = 1e-6; int max_iter = 1000; int its = 0; double relres = 0.; };

namespace {
struct Record {
    refhypre_record r;
    std::vector<double> matrix, rhs, solution;
} g_rec;
refhypre_solve_fn g_solver = nullptr;
int g_nsolves = 0;

inline int gnx(const refhypre_grid &g) { return g.ihi[0] - g.ilo[0] + 1; }
inline int gny(const refhypre_grid &g) { return g.ihi[1] - g.ilo[1] + 1; }
inline long cell(const refhypre_grid &g, const int *idx) { return (long)(idx[0] - g.ilo[0]) * gny(g) + (idx[1] - g.ilo[1]); }

// y = A x on the cells that are not Dirichlet rows ("free"), x of Dirichlet cells taken as 0
struct Sys {
    int nx, ny, ne; const int (*off)[2]; const double *a; std::vector<char> fixed;
    void apply(const std::vector<double> &x, std::vector<double> &y) const {
        #pragma omp parallel for schedule(static)
        for (int i = 0; i < nx; i++)
            for (int j = 0; j < ny; j++) {
                const long k = (long)i * ny + j;
                if (fixed[k]) { y[k] = 0.; continue; }
                double s = 0.;
                for (int e = 0; e < ne; e++) {
                    const double c = a[k * ne + e];
                    if (c == 0.) continue;
                    const int ii = i + off[e][0], jj = j + off[e][1];
                    if (ii < 0 || ii >= nx || jj < 0 || jj >= ny) continue;      // the reference zeroes (folds) such entries
                    const long kk = (long)ii * ny + jj;
                    if (!fixed[kk]) s += c * x[kk];
                }
                y[k] = s;
            }
    }
    // z = M^-1 r, M = symmetric Gauss-Seidel (forward then backward lexicographic sweep from zero)
    void sgs(const std::vector<double> &r, std::vector<double> &z) const {
        std::fill(z.begin(), z.end(), 0.);
        for (int pass = 0; pass < 2; pass++)
            for (long t = 0; t < (long)nx * ny; t++) {
                const long k = pass == 0 ? t : (long)nx * ny - 1 - t;
                if (fixed[k]) continue;
                const int i = (int)(k / ny), j = (int)(k % ny);
                double s = r[k], d = 0.;
                for (int e = 0; e < ne; e++) {
                    const double c = a[k * ne + e];
                    if (off[e][0] == 0 && off[e][1] == 0) { d = c; continue; }
                    if (c == 0.) continue;
                    const int ii = i + off[e][0], jj = j + off[e][1];
                    if (ii < 0 || ii >= nx || jj < 0 || jj >= ny) continue;
                    const long kk = (long)ii * ny + jj;
                    if (!fixed[kk]) s -= c * z[kk];
                }
                z[k] = s / d;
            }
    }
};
double dot(const std::vector<double> &a, const std::vector<double> &b) {
    double s = 0.;
    #pragma omp parallel for reduction(+ : s) schedule(static)
    for (long k = 0; k < (long)a.size(); k++) s += a[k] * b[k];
    return s;
}
}  // namespace

extern "C" {
const refhypre_record *refhypre_last(void) { g_rec.r.matrix = g_rec.matrix.data(); g_rec.r.rhs = g_rec.rhs.data(); g_rec.r.solution = g_rec.solution.data(); g_rec.r.nsolves = g_nsolves; return &g_rec.r; }
void refhypre_set_solver(refhypre_solve_fn fn) { g_solver = fn; }

HYPRE_Int HYPRE_StructGridCreate(MPI_Comm, HYPRE_Int ndim, HYPRE_StructGrid *grid) { if (ndim != 2) { fprintf(stderr, "hypre shim: 2-D grids only\n"); abort(); } *grid = new refhypre_grid(); return 0; }
HYPRE_Int HYPRE_StructGridSetExtents(HYPRE_StructGrid g, HYPRE_Int *lo, HYPRE_Int *hi) { g->ilo[0] = lo[0]; g->ilo[1] = lo[1]; g->ihi[0] = hi[0]; g->ihi[1] = hi[1]; return 0; }
HYPRE_Int HYPRE_StructGridAssemble(HYPRE_StructGrid) { return 0; }
HYPRE_Int HYPRE_StructGridDestroy(HYPRE_StructGrid g) { delete g; return 0; }
HYPRE_Int HYPRE_StructStencilCreate(HYPRE_Int ndim, HYPRE_Int size, HYPRE_StructStencil *s) { if (ndim != 2 || size > 16) abort(); *s = new refhypre_stencil(); (*s)->n = size; return 0; }
HYPRE_Int HYPRE_StructStencilSetElement(HYPRE_StructStencil s, HYPRE_Int e, HYPRE_Int *off) { s->off[e][0] = off[0]; s->off[e][1] = off[1]; return 0; }
HYPRE_Int HYPRE_StructStencilDestroy(HYPRE_StructStencil s) { delete s; return 0; }
HYPRE_Int HYPRE_StructMatrixCreate(MPI_Comm, HYPRE_StructGrid g, HYPRE_StructStencil s, HYPRE_StructMatrix *m) { *m = new refhypre_matrix(); (*m)->g = *g; (*m)->s = *s; return 0; }
HYPRE_Int HYPRE_StructMatrixInitialize(HYPRE_StructMatrix m) { m->v.assign((size_t)gnx(m->g) * gny(m->g) * m->s.n, 0.); return 0; }
HYPRE_Int HYPRE_StructMatrixSetValues(HYPRE_StructMatrix m, HYPRE_Int *idx, HYPRE_Int n, HYPRE_Int *entries, double *values) {
    const long k = cell(m->g, idx);
    for (int e = 0; e < n; e++) m->v[(size_t)k * m->s.n + entries[e]] = values[e];
    return 0;
}
HYPRE_Int HYPRE_StructMatrixAssemble(HYPRE_StructMatrix) { return 0; }
HYPRE_Int HYPRE_StructMatrixDestroy(HYPRE_StructMatrix m) { delete m; return 0; }
HYPRE_Int HYPRE_StructVectorCreate(MPI_Comm, HYPRE_StructGrid g, HYPRE_StructVector *v) { *v = new refhypre_vector(); (*v)->g = *g; return 0; }
HYPRE_Int HYPRE_StructVectorInitialize(HYPRE_StructVector v) { v->v.assign((size_t)gnx(v->g) * gny(v->g), 0.); return 0; }
HYPRE_Int HYPRE_StructVectorSetValues(HYPRE_StructVector v, HYPRE_Int *idx, double value) { v->v[cell(v->g, idx)] = value; return 0; }
HYPRE_Int HYPRE_StructVectorSetBoxValues(HYPRE_StructVector v, HYPRE_Int *lo, HYPRE_Int *hi, double *values) {
    // HYPRE's box values run with the FIRST index fastest
    long t = 0;
    for (int j = lo[1]; j <= hi[1]; j++) for (int i = lo[0]; i <= hi[0]; i++) { const int idx[2] = {i, j}; v->v[cell(v->g, idx)] = values[t++]; }
    return 0;
}
HYPRE_Int HYPRE_StructVectorGetValues(HYPRE_StructVector v, HYPRE_Int *idx, double *value) { *value = v->v[cell(v->g, idx)]; return 0; }
HYPRE_Int HYPRE_StructVectorAssemble(HYPRE_StructVector) { return 0; }
HYPRE_Int HYPRE_StructVectorDestroy(HYPRE_StructVector v) { delete v; return 0; }

static HYPRE_Int solve_recorded(HYPRE_StructSolver s, HYPRE_StructMatrix A, HYPRE_StructVector b, HYPRE_StructVector x) {
    const int nx = gnx(A->g), ny = gny(A->g), ne = A->s.n;
    const long n = (long)nx * ny;
    g_rec.r.nx = nx; g_rec.r.ny = ny; g_rec.r.nentries = ne; g_rec.r.tol = s->tol;
    for (int e = 0; e < ne && e < 9; e++) { g_rec.r.offsets[e][0] = A->s.off[e][0]; g_rec.r.offsets[e][1] = A->s.off[e][1]; }
    g_rec.matrix = A->v; g_rec.rhs = b->v;
    g_nsolves++;
    if (g_solver) {
        g_rec.solution.assign(n, 0.);
        refhypre_last();
        if (g_solver(&g_rec.r, g_rec.solution.data()) != 0) { fprintf(stderr, "hypre shim: external solver failed\n"); abort(); }
        x->v = g_rec.solution; s->its = 1; s->relres = 0.; g_rec.r.iterations = 1; g_rec.r.final_relres = 0.; g_rec.r.symmetric = -1;
        return 0;
    }
    Sys S; S.nx = nx; S.ny = ny; S.ne = ne; S.off = A->s.off; S.a = A->v.data(); S.fixed.assign(n, 0);
    int ec = -1;
    for (int e = 0; e < ne; e++) if (A->s.off[e][0] == 0 && A->s.off[e][1] == 0) ec = e;
    if (ec < 0) { fprintf(stderr, "hypre shim: stencil without a centre entry\n"); abort(); }
    // Dirichlet rows: only the diagonal is set (src/MGSolver.cpp:517-567).  x_k = b_k / a_kk; move their columns to the right-hand side.
    std::vector<double> sol(n, 0.), rhs(b->v);
    for (long k = 0; k < n; k++) {
        bool only_diag = true;
        for (int e = 0; e < ne; e++) if (e != ec && A->v[(size_t)k * ne + e] != 0.) { only_diag = false; break; }
        if (only_diag) { S.fixed[k] = 1; sol[k] = b->v[k] / A->v[(size_t)k * ne + ec]; }
    }
    for (int i = 0; i < nx; i++)
        for (int j = 0; j < ny; j++) {
            const long k = (long)i * ny + j;
            if (S.fixed[k]) { rhs[k] = 0.; continue; }
            for (int e = 0; e < ne; e++) {
                const int ii = i + A->s.off[e][0], jj = j + A->s.off[e][1];
                if (ii < 0 || ii >= nx || jj < 0 || jj >= ny) continue;
                const long kk = (long)ii * ny + jj;
                if (S.fixed[kk]) rhs[k] -= A->v[(size_t)k * ne + e] * sol[kk];
            }
        }
    // symmetry of the reduced operator
    bool sym = true; double amax = 0.;
    for (size_t t = 0; t < A->v.size(); t++) amax = std::max(amax, fabs(A->v[t]));
    for (int i = 0; i < nx && sym; i++)
        for (int j = 0; j < ny && sym; j++) {
            const long k = (long)i * ny + j;
            if (S.fixed[k]) continue;
            for (int e = 0; e < ne; e++) {
                const int ii = i + A->s.off[e][0], jj = j + A->s.off[e][1];
                if (ii < 0 || ii >= nx || jj < 0 || jj >= ny) continue;
                const long kk = (long)ii * ny + jj;
                if (S.fixed[kk]) continue;
                int er = -1;
                for (int f = 0; f < ne; f++) if (A->s.off[f][0] == -A->s.off[e][0] && A->s.off[f][1] == -A->s.off[e][1]) er = f;
                if (er < 0 || fabs(A->v[(size_t)k * ne + e] - A->v[(size_t)kk * ne + er]) > 1e-12 * amax) { sym = false; break; }
            }
        }
    g_rec.r.symmetric = sym ? 1 : 0;
    // the diagonal of the pressure operator is negative (C = -sum): SGS handles either sign
    const double bnorm2 = dot(b->v, b->v);
    std::vector<double> xx(n, 0.), r(rhs), z(n), p(n), q(n);
    int it = 0; double rr = dot(r, r);
    const double stop = s->tol * s->tol * (bnorm2 > 0. ? bnorm2 : 1.);
    if (sym) {
        S.sgs(r, z); p = z; double rz = dot(r, z);
        while (rr >= stop && it < s->max_iter) {
            S.apply(p, q);
            const double alpha = rz / dot(p, q);
            for (long k = 0; k < n; k++) { xx[k] += alpha * p[k]; r[k] -= alpha * q[k]; }
            rr = dot(r, r); it++;
            if (rr < stop) break;
            S.sgs(r, z);
            const double rz1 = dot(r, z), beta = rz1 / rz; rz = rz1;
            for (long k = 0; k < n; k++) p[k] = z[k] + beta * p[k];
        }
    } else {                                            // right-preconditioned BiCGStab
        std::vector<double> r0(r), v(n, 0.), ph(n), sh(n), sv(n), t(n);
        double rho = 1., alpha = 1., om = 1.;
        std::fill(p.begin(), p.end(), 0.);
        while (rr >= stop && it < s->max_iter) {
            const double rho1 = dot(r0, r), beta = (rho1 / rho) * (alpha / om); rho = rho1;
            for (long k = 0; k < n; k++) p[k] = r[k] + beta * (p[k] - om * v[k]);
            S.sgs(p, ph); S.apply(ph, v);
            alpha = rho / dot(r0, v);
            for (long k = 0; k < n; k++) sv[k] = r[k] - alpha * v[k];
            S.sgs(sv, sh); S.apply(sh, t);
            om = dot(t, sv) / dot(t, t);
            for (long k = 0; k < n; k++) { xx[k] += alpha * ph[k] + om * sh[k]; r[k] = sv[k] - om * t[k]; }
            rr = dot(r, r); it++;
        }
    }
    for (long k = 0; k < n; k++) if (!S.fixed[k]) sol[k] = xx[k];
    x->v = sol; g_rec.solution = sol;
    s->its = it; s->relres = bnorm2 > 0. ? sqrt(rr / bnorm2) : 0.;
    g_rec.r.iterations = it; g_rec.r.final_relres = s->relres;
    return 0;
}

#define SOLVER_COMMON(NAME) \
    HYPRE_Int HYPRE_Struct##NAME##Create(MPI_Comm, HYPRE_StructSolver *s) { *s = new refhypre_solver(); return 0; } \
    HYPRE_Int HYPRE_Struct##NAME##Destroy(HYPRE_StructSolver s) { delete s; return 0; } \
    HYPRE_Int HYPRE_Struct##NAME##Setup(HYPRE_StructSolver, HYPRE_StructMatrix, HYPRE_StructVector, HYPRE_StructVector) { return 0; } \
    HYPRE_Int HYPRE_Struct##NAME##Solve(HYPRE_StructSolver s, HYPRE_StructMatrix A, HYPRE_StructVector b, HYPRE_StructVector x) { return solve_recorded(s, A, b, x); } \
    HYPRE_Int HYPRE_Struct##NAME##SetTol(HYPRE_StructSolver s, double tol) { s->tol = tol; return 0; } \
    HYPRE_Int HYPRE_Struct##NAME##SetMaxIter(HYPRE_StructSolver s, HYPRE_Int m) { s->max_iter = m; return 0; } \
    HYPRE_Int HYPRE_Struct##NAME##GetNumIterations(HYPRE_StructSolver s, HYPRE_Int *n) { *n = s->its; return 0; } \
    HYPRE_Int HYPRE_Struct##NAME##GetFinalRelativeResidualNorm(HYPRE_StructSolver s, double *r) { *r = s->relres; return 0; }
SOLVER_COMMON(PCG)
SOLVER_COMMON(SMG)
SOLVER_COMMON(Jacobi)
HYPRE_Int HYPRE_StructPCGSetTwoNorm(HYPRE_StructSolver, HYPRE_Int) { return 0; }
HYPRE_Int HYPRE_StructPCGSetRelChange(HYPRE_StructSolver, HYPRE_Int) { return 0; }
HYPRE_Int HYPRE_StructPCGSetPrecond(HYPRE_StructSolver, HYPRE_PtrToStructSolverFcn, HYPRE_PtrToStructSolverFcn, HYPRE_StructSolver) { return 0; }
HYPRE_Int HYPRE_StructPCGSetLogging(HYPRE_StructSolver, HYPRE_Int) { return 0; }
HYPRE_Int HYPRE_StructPCGSetPrintLevel(HYPRE_StructSolver, HYPRE_Int) { return 0; }
HYPRE_Int HYPRE_StructSMGSetMemoryUse(HYPRE_StructSolver, HYPRE_Int) { return 0; }
HYPRE_Int HYPRE_StructSMGSetRelChange(HYPRE_StructSolver, HYPRE_Int) { return 0; }
HYPRE_Int HYPRE_StructSMGSetZeroGuess(HYPRE_StructSolver) { return 0; }
HYPRE_Int HYPRE_StructSMGSetNumPreRelax(HYPRE_StructSolver, HYPRE_Int) { return 0; }
HYPRE_Int HYPRE_StructSMGSetNumPostRelax(HYPRE_StructSolver, HYPRE_Int) { return 0; }
HYPRE_Int HYPRE_StructSMGSetLogging(HYPRE_StructSolver, HYPRE_Int) { return 0; }
HYPRE_Int HYPRE_StructJacobiSetNonZeroGuess(HYPRE_StructSolver) { return 0; }
}  // extern "C"
