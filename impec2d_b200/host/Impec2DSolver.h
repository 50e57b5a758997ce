// Impec2DSolver -- the driver class, same surface as src/Impec2DSolver.h:54-57.
#ifndef IMPEC2DSOLVER_H_
#define IMPEC2DSOLVER_H_
#include <string>
#include <vector>
#include "EllipticSolver.h"
#include "HyperbolicSolver.h"
#include "GenSimDefs.h"
#include "SimParams.h"
struct DiagState { int ctr; int its; double t, dt, mass_err, diff_err, vel_err, fillperdt; };
class Impec2DSolver {
    Array2d p, c;
    Array3d v;
    EllipticSolver *es;
    HyperbolicSolver *hs;
    double *q;
    int myid, num_procs;
public:
    SimParams *sp;
    Impec2DSolver(const std::string &configfile, int myid, int num_procs);
    virtual ~Impec2DSolver();
    bool Run();
    // extras for embedding / tests
    const Array2d &Concentration() const { return c; }
    const Array2d &Pressure() const { return p; }
    const std::vector<DiagState> &Diagnostics() const { return diagstates; }
    double ElapsedMs() const { return elapsed_ms; }
    int max_steps = -1;                               // stop after this many steps (-1: run to endtime)
    // binary restart (impec_checkpoint_save / _load: exact device state, the restarted run continues bit for bit)
    std::string checkpoint_file, restart_file;        // write to / start from this file (empty: off)
    int checkpoint_every = 0;                         // also write every N steps (0: only at the end of the run)
private:
    std::vector<DiagState> diagstates;
    double elapsed_ms = 0;
    void Initialize();
    void ResizeArrays();
    void InitializeConcentration();
    void InitializeSource();
    bool Solve();
    bool TestComponent();
    bool TestAdvection();
    double Advection();
    void OutputState(const DiagState &ds);
    void WriteDiagnostics();
    std::string RankFile(const std::string &f) const;
    void LogResults(double postfix, bool logTime);
    void Project(Array2d &arrd, Array2d const *const arrp, int bounds, int &d_N1, int &d_N2, double &d_dx, double &d_dy);   // src/Impec2DSolver.h
};
#endif
