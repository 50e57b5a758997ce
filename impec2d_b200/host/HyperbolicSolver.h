// HyperbolicSolver / HyperbolicSplitSolver / MUSCLSplitSolver / UpwindSplitSolver -- transport interface, same
// surface as src/HyperbolicSolver.h:27-39, src/HyperbolicSplitSolver.h:19-22, src/MUSCLSplitSolver.h,
// src/UpwindSplitSolver.h.  The flux evaluation is fused into the sweep kernels, so Advect_X / Advect_Y never
// materialise F or G.
#ifndef HYPERBOLICSOLVER_H_
#define HYPERBOLICSOLVER_H_
#include "GenSimDefs.h"
#include "SimParams.h"
#include "TimeIntegrator.h"
#include "EllipticSolver.h"
enum SWEEP { X_SWEEP, Y_SWEEP };
class HyperbolicSolver {
private:
    TimeIntegrator *ti;
    MGSolver *mg;
protected:
    SimParams *sp;
    double total_mass_outflux;
    SWEEP prev_first_sweep;
    bool SourceSplit(Array2d &c, double const *const q, const double dt);
    double Flux(double vel, double conc) { return vel * conc / sp->porosity; }
    void ReportParabolicIts(int n) { if (mg) mg->SetLastParabolicIts(n); }   // CN solves run inside the fused transport step on the device
public:
    virtual bool Solve(Array2d &c, const Array3d &v, double const *const q, const double dt) = 0;
    HyperbolicSolver(SimParams *sp);
    virtual ~HyperbolicSolver();
    double GetTotalMassOutflux();
    void FillBoundaries(Array2d &c);
    double MaxStableTimestepCombined(const Array2d &c, const Array3d &v);
    double MaxStableAdvectionTimestep(const Array2d &c, const Array3d &v);
    double MaxStableTimestep(const Array2d &c, const Array3d &v);
    bool DiffusionSplit(Array2d &c, const Array3d &v, const double dt);
    bool DiffusionSplitCN(Array2d &c, const Array3d &v, const double dt);
    void AssignMGSolver(MGSolver *mg) { this->mg = mg; }
};
class HyperbolicSplitSolver : public HyperbolicSolver {
protected:
    bool AdvectSplit(Array2d &c, const Array3d &v, double const *const q, const double dt);
    virtual bool Solve(Array2d &c, const Array3d &v, double const *const q, const double dt);
    virtual int MethodId() const = 0;                    // IMPEC_MUSCL | IMPEC_UPWIND
public:
    // The reference exposes the raw intercell fluxes (src/HyperbolicSplitSolver.h:19-22); the sweep kernels evaluate theirs on the fly, so
    // these two run a dedicated flux kernel and fill F / G in the reference's shape.
    virtual void CalculateIntercellFlux_F(Array2d &F, const Array2d &c, const Array3d &v, const double dt);
    virtual void CalculateIntercellFlux_G(Array2d &G, const Array2d &c, const Array3d &v, const double dt);
    bool Advect_X(Array2d &c, const Array3d &v, const double dt);
    bool Advect_Y(Array2d &c, const Array3d &v, const double dt);
    HyperbolicSplitSolver(SimParams *sp) : HyperbolicSolver(sp) {}
    virtual ~HyperbolicSplitSolver() {}
};
class MUSCLSplitSolver : public HyperbolicSplitSolver {
protected:
    virtual int MethodId() const;
public:
    MUSCLSplitSolver(SimParams *sp);
    virtual ~MUSCLSplitSolver() {}
};
class UpwindSplitSolver : public HyperbolicSplitSolver {
protected:
    virtual int MethodId() const;
public:
    UpwindSplitSolver(SimParams *sp);
    virtual ~UpwindSplitSolver() {}
    double CalculateIntercellFlux(double SPEED, double UL, double UR) { return (SPEED > 0) ? Flux(SPEED, UL) : Flux(SPEED, UR); }
};
#endif
