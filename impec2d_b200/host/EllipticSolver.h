// EllipticSolver / MGSolver -- pressure interface, same surface as src/EllipticSolver.h:58-69 and
// src/MGSolver.h:62-66.  "MGSolver" upstream wraps HYPRE StructPCG+SMG; here it drives the B200 multigrid-PCG
// (impec_pressure_solve) and the Crank-Nicolson dispersion solve (impec_diffusion_split).
#ifndef ELLIPTICSOLVER_H_
#define ELLIPTICSOLVER_H_
#include "GenSimDefs.h"
#include "SimParams.h"
class GnuPlotter;
class EllipticSolver {
    int total_its, calls;
    double nu;
protected:
    int last_its, last_parabolic_its;
    double last_relres, last_true_relres;
    SimParams *sp;
public:
    EllipticSolver(SimParams *sp);
    virtual ~EllipticSolver();
    virtual bool SolveLinearMultigrid(Array2d &c_c, int c_bounds, Array2d &p_c, int p_bounds, double const *const q_v) = 0;
    bool Solve(Array2d &p, Array2d &c, double *q);
    void CalcVelocities(Array2d &p, Array3d &v, Array2d &c);
    void SetNu(double nu);
    int GetLastIts() { return last_its; }
    int GetLastParabolicIts() { return last_parabolic_its; }
    void SetLastParabolicIts(int n) { last_parabolic_its = n; }        // extra: the CN solves run inside the fused transport step on the device
    double GetAvgIts() { return calls > 0 ? ((double)total_its) / calls : 0; }
    double GetLastRelativeResidual() { return last_relres; }          // extra: what HYPRE's GetFinalRelativeResidualNorm returned
    double GetLastTrueRelativeResidual() { return last_true_relres; } // extra: ||b - Ax||/||b|| recomputed
    bool TestNeumann(GnuPlotter *gp);
protected:
    void Initialize();
};
class MGSolver : public EllipticSolver {
public:
    MGSolver(SimParams *sp);
    virtual ~MGSolver();
    bool Solve_Elliptic_cc_to_c(const Array2d &c_c, int c_bounds, Array2d &p_c, int p_bounds, double const *const q_c);
    bool Solve_Parabolic_cc_to_c(Array2d &c_c, int c_bounds, const Array3d &v, int v_bounds, Array2d const *const ls, int ls_bounds, const double dt);
    virtual bool SolveLinearMultigrid(Array2d &c_c, int c_bounds, Array2d &p_c, int p_bounds, double const *const q_v);
};
#endif
