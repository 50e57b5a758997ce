// Impec2DSolver.cpp -- driver: owns p, c, v, q; time loop of src/Impec2DSolver.cpp:293-466; self tests
// (TestComponent :719-744, TestAdvection :773-846, Advection :848-991).  Fields live on the GPU for the whole loop;
// the host copies are touched only for logging.
#include "Impec2DSolver.h"
#include "DeviceContext.h"
#include "Util.h"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
using namespace std;

Impec2DSolver::Impec2DSolver(const string &configfile, int myid, int num_procs) : es(NULL), hs(NULL), q(NULL), myid(myid), num_procs(num_procs) {
    sp = new SimParams(configfile);          // every rank parses the file itself (no token ring needed: read-only)
    DeviceContext::SetRanks(myid, num_procs);                          // one process per GPU, slabs of rows (src/Driver.cpp:58-87 passes the MPI rank here)
    if (sp->gnuplot && myid == 0) cout << "(gnuplot live view is not part of this build; results are written to files)" << endl;
    Initialize();
}
Impec2DSolver::~Impec2DSolver() {
    p.Unbind(); c.Unbind(); v.Unbind();
    delete[] q; delete es; delete hs;
    DeviceContext::Release(sp);
    delete sp;
}
void Impec2DSolver::Initialize() {
    es = new MGSolver(sp);                   // hard-wired upstream as well (src/Impec2DSolver.cpp:131-133)
    hs = new MUSCLSplitSolver(sp);
    hs->AssignMGSolver((MGSolver *)es);
    ResizeArrays();
    InitializeConcentration();
    InitializeSource();
}
void Impec2DSolver::ResizeArrays() {
    Idx2d idx; idx[0] = -sp->p_bounds; idx[1] = -sp->p_bounds;
    p.resize(sp->N1 + 2 * sp->p_bounds, sp->N2 + 2 * sp->p_bounds); p.reindexSelf(idx);
    c.resize(sp->N1 + 2 * sp->p_bounds, sp->N2 + 2 * sp->p_bounds); c.reindexSelf(idx);
    Idx3d i3; i3[0] = -sp->v_bounds; i3[1] = -sp->v_bounds; i3[2] = 0;
    v.resize(sp->N1 + 1 + 2 * sp->v_bounds, sp->N2 + 1 + 2 * sp->v_bounds, 2); v.reindexSelf(i3);
    delete[] q; q = new double[sp->M];
}
// "x y c" triples in i-major order (src/Impec2DSolver.cpp:194-220)
void Impec2DSolver::InitializeConcentration() {
    c = 0.;
    if (sp->init_file.empty()) return;
    ifstream f(sp->init_file.c_str());
    if (!f.is_open()) { cout << "Unable to open file" << sp->init_file << "; Initializing to zero." << endl; return; }
    string line; int ctr = 0; double x, y, conc;
    while (getline(f, line)) {
        if (line.empty()) continue;
        stringstream ss(line);
        if (!(ss >> x >> y >> conc)) continue;
        if (ctr < sp->M) c(ctr / sp->N2, ctr % sp->N2) = conc;
        ctr++;
    }
    if (ctr != sp->N1 * sp->N2) { cerr << "Concentration Input FILE:: wrong dimensions!" << endl; exit(1); }
}
void Impec2DSolver::InitializeSource() { InitSource(sp, q); }

// one binary restart file per rank (each holds that rank's slab)
string Impec2DSolver::RankFile(const string &f) const { if (num_procs <= 1) return f; stringstream ss; ss << f << ".rank" << myid; return ss.str(); }

bool Impec2DSolver::Run() { return sp->is_test ? TestComponent() : Solve(); }

bool Impec2DSolver::Solve() {
    bool state = true;
    const auto t_start = chrono::steady_clock::now();
    double t = sp->tstart, dt = 0;
    size_t tidx = 0;
    int counter = 0;
    if (myid == 0) cout << "Final time: " << sp->tfinal << endl;
    if (find(sp->times.begin(), sp->times.end(), sp->tfinal) == sp->times.end()) sp->times.push_back(sp->tfinal);
    sort(sp->times.begin(), sp->times.end());
    impec_ctx *ctx = DeviceContext::Get(sp);
    double maxc = 0, minc = 10e10, avg_fill = 0;
    diagstates.clear();
    if (!restart_file.empty()) {
        DeviceContext::Use(ctx, &c, &p, &v, q, sp);                    // bind the arrays, upload q; then replace the device state
        long long step0 = 0;
        DeviceContext::Check(ctx, impec_checkpoint_load(ctx, RankFile(restart_file).c_str(), &t, &step0), "impec_checkpoint_load");
        c.MarkDeviceNewer(); p.MarkDeviceNewer(); v.MarkDeviceNewer();
        counter = (int)step0;
        while (tidx < sp->times.size() && sp->times[tidx] <= t) tidx++;
        if (myid == 0) cout << "Restarted from " << restart_file << " at step " << counter << ", t = " << t << endl;
    }
    const int first_step = counter;
    while (state && t < sp->tfinal && (max_steps < 0 || counter - first_step < max_steps)) {
        counter++;
        hs->FillBoundaries(c);                                         // :343
        state = state && es->Solve(p, c, q);                           // :348
        if (!state) { cerr << "ERROR::EllipticSolver: [" << counter << "]\t" << "t:\t" << t << endl; break; }
        es->CalcVelocities(p, v, c);                                   // :358
        dt = hs->MaxStableTimestep(c, v);                              // :370
        if (tidx < sp->times.size() && t + dt >= sp->times[tidx]) { dt = sp->times[tidx] - t; tidx++; }   // :375-378
        state = state && hs->Solve(c, v, q, dt);                       // :383
        if (!state) { cerr << "ERROR::HyperbolicSolver: [" << counter << "]\t" << "t:\t" << "\tdt:\t" << dt << endl; break; }
        t += dt;
        DiagState ds;
        ds.ctr = counter; ds.t = t; ds.dt = dt; ds.its = es->GetLastIts();
        ds.fillperdt = dt * sp->I / (sp->dx * sp->dy); avg_fill += ds.fillperdt;
        impec_diag d;                                                   // one fused reduction pass instead of :530-620's four loops
        DeviceContext::Check(ctx, impec_diagnostics(ctx, &d), "impec_diagnostics");
        double influx = (t - sp->tstart) * sp->I / (sp->dx * sp->dy);
        if (sp->stype == PARALLEL) influx *= 2.;
        ds.mass_err = sp->check_mass_conservation ? fabs((d.total_mass + d.total_mass_outflux - influx) / influx) : -1;
        ds.vel_err = sp->check_noflow_boundary ? d.boundary_velocity_error : -1;
        ds.diff_err = sp->check_divergence ? d.divergence_error : -1;
        diagstates.push_back(ds);
        if (myid == 0) OutputState(ds);
        if (sp->log_to_file) {                                          // (every rank takes the same decisions: LogResults gathers collectively)
            if (find(sp->times.begin(), sp->times.end(), t) != sp->times.end()) LogResults(t, true);
            const bool listed = find(sp->counters.begin(), sp->counters.end(), counter) != sp->counters.end();
            if (listed || (sp->log_skip_steps != 0 && (counter == 1 || counter % sp->log_skip_steps == 0))) LogResults(counter, false);
        }
        maxc = max(maxc, d.c_max); minc = min(minc, d.c_min);
        if (sp->stype == HORIZONTAL && sp->im != CONSTANT) InitHorizontalFront(sp, q, t);      // time-dependent injection (:437-438)
        if (!checkpoint_file.empty() && checkpoint_every > 0 && counter % checkpoint_every == 0)
            DeviceContext::Check(ctx, impec_checkpoint_save(ctx, RankFile(checkpoint_file).c_str(), t, counter), "impec_checkpoint_save");
    }
    if (!checkpoint_file.empty()) DeviceContext::Check(ctx, impec_checkpoint_save(ctx, RankFile(checkpoint_file).c_str(), t, counter), "impec_checkpoint_save");
    if (counter - first_step) avg_fill /= (counter - first_step);
    elapsed_ms = chrono::duration<double, milli>(chrono::steady_clock::now() - t_start).count();
    if (myid == 0) {
        cout << "overall maximum: " << maxc << endl << "overall minimum: " << minc << endl << "average filled cells per dt: " << avg_fill << endl;
        cout << "---------------------------------------------------------" << endl;
        if (counter > 0) cout << "Average no. of Iterations: " << es->GetAvgIts() << endl;
        cout << "Elapsed time [ms]: " << (long)(elapsed_ms + 0.5) << endl;
        WriteDiagnostics();
    }
    LogResults(-1, false);
    return state;
}

void Impec2DSolver::OutputState(const DiagState &ds) {
    stringstream ss;
    ss.setf(ios_base::fixed, ios_base::floatfield); ss.precision(6);
    ss << "[" << ds.ctr << "]\t" << "ITS_e: " << ds.its << "\tITS_p: " << es->GetLastParabolicIts() << "\t t: " << ds.t << "\tdt: " << ds.dt << "\tcells/dt: " << ds.fillperdt;
    if (sp->check_mass_conservation) ss << "\tE_m: " << ds.mass_err * 100;
    if (sp->check_noflow_boundary) ss << "\tE_v: " << ds.vel_err;
    if (sp->check_divergence) ss << "\tE_div: " << ds.diff_err;
    cout << ss.str() << endl;
}
void Impec2DSolver::WriteDiagnostics() {
    const string fname = sp->results_dir + "/diagnostics.dat";
    ofstream file(fname.c_str());
    if (!file.is_open()) { cerr << "Unable to write to diagnostics-file: " << fname << endl; return; }
    const string SEP = "    ";
    for (const DiagState &ds : diagstates) {
        file << ds.ctr << SEP << ds.its << SEP << ds.t << SEP << ds.dt << SEP << ds.fillperdt << SEP;
        if (sp->check_mass_conservation) file << ds.mass_err << SEP;
        if (sp->check_noflow_boundary) file << ds.vel_err << SEP;
        if (sp->check_divergence) file << ds.diff_err << SEP;
        file << endl;
    }
}
// Impec2DSolver::Project (src/Impec2DSolver.cpp:690-716): the parallel-layout field arrp (the concentration or the pressure of
// this solver) interpolated onto the diagonal grid of round(N/sqrt 2) cells by the device kernel behind
// impec_project_to_diagonal (ParallelProjector::FillDiagonal, src/ParallelProjector.cpp:176-199).  arrd gets `bounds` zero ghosts.
void Impec2DSolver::Project(Array2d &arrd, Array2d const *const arrp, int bounds, int &d_N1, int &d_N2, double &d_dx, double &d_dy) {
    impec_ctx *ctx = DeviceContext::Get(sp);
    if (arrp != &c && arrp != &p) { cerr << "Impec2DSolver::Project: only the solver's own concentration / pressure can be projected" << endl; exit(1); }
    DeviceContext::Use(ctx, &c, &p, NULL, NULL, sp);
    DeviceContext::Check(ctx, impec_project_dims(ctx, &d_N1, &d_N2), "impec_project_dims");
    vector<double> buf((size_t)d_N1 * d_N2);
    DeviceContext::Check(ctx, impec_project_to_diagonal(ctx, arrp == &p ? 1 : 0, buf.data()), "impec_project_to_diagonal");
    arrd.resize(d_N1 + 2 * bounds, d_N2 + 2 * bounds);
    Idx2d idx; idx[0] = -bounds; idx[1] = -bounds;
    arrd.reindexSelf(idx);
    arrd = 0;
    for (int i = 0; i < d_N1; i++) for (int j = 0; j < d_N2; j++) arrd(i, j) = buf[(size_t)i * d_N2 + j];
    d_dx = (sp->x_ub - sp->x_lb) / (sqrt(2.) * d_N1);
    d_dy = (sp->y_ub - sp->y_lb) / (sqrt(2.) * d_N2);
}

// concentration (and, at the end, pressure and velocity) in the reference's text format (src/Impec2DSolver.cpp:622-686).
// As upstream, a PARALLEL run writes c and p projected onto the diagonal grid (:650-667) and the velocity on the native grid.
void Impec2DSolver::LogResults(double postfix, bool logTime) {
    string dir = sp->log_dir, post = ".dat";
    if (postfix < 0) dir = sp->results_dir;
    else { stringstream ss; ss << "_"; if (logTime) ss << "t_"; ss << postfix; post = ss.str() + ".dat"; }
    if ((size_t)sp->N1 * sp->N2 > (size_t)4096 * 4096) { if (postfix < 0 && myid == 0) cout << "(text output skipped above 4096^2)" << endl; return; }
    if (num_procs > 1) {
        // Slabs: every call below is collective (each rank contributes its rows), rank 0 writes.  A PARALLEL run projects c and p onto
        // the diagonal grid as on one GPU (impec_project_to_diagonal sums the ranks' parts of every chunk of diagonal rows).
        impec_ctx *ctx = DeviceContext::Get(sp);
        if (sp->stype == PARALLEL) {
            int d_N1, d_N2; double d_dx, d_dy;
            Array2d diag_c, diag_p;
            Project(diag_c, &c, sp->c_bounds, d_N1, d_N2, d_dx, d_dy);
            if (postfix < 0) Project(diag_p, &p, sp->p_bounds, d_N1, d_N2, d_dx, d_dy);
            if (postfix < 0) DeviceContext::GatherGlobal(ctx, NULL, NULL, &v);
            if (myid != 0) return;
            PrintDToPlotFile(dir + "/concentration" + post, diag_c, sp->x_lb, sp->y_lb, d_dx, d_dy, sp->c_bounds);
            if (postfix < 0) {
                PrintDToPlotFile(dir + "/pressure" + post, diag_p, sp->x_lb, sp->y_lb, d_dx, d_dy, sp->p_bounds);
                PrintVelToPlotFile(dir + "/velocity_u" + post, dir + "/velocity_v" + post, v, sp->x_lb, sp->y_lb, sp->dx, sp->dy, sp->v_bounds, sp->N1, sp->N2);
            }
            return;
        }
        DeviceContext::GatherGlobal(ctx, &c, postfix < 0 ? &p : NULL, postfix < 0 ? &v : NULL);
        if (myid != 0) return;
        PrintDToPlotFile(dir + "/concentration" + post, c, sp->x_lb, sp->y_lb, sp->dx, sp->dy, sp->c_bounds);
        if (postfix < 0) {
            PrintDToPlotFile(dir + "/pressure" + post, p, sp->x_lb, sp->y_lb, sp->dx, sp->dy, sp->p_bounds);
            PrintVelToPlotFile(dir + "/velocity_u" + post, dir + "/velocity_v" + post, v, sp->x_lb, sp->y_lb, sp->dx, sp->dy, sp->v_bounds, sp->N1, sp->N2);
        }
        return;
    }
    if (sp->stype == PARALLEL) {
        int d_N1, d_N2; double d_dx, d_dy;
        Array2d diag_c, diag_p;
        Project(diag_c, &c, sp->c_bounds, d_N1, d_N2, d_dx, d_dy);
        PrintDToPlotFile(dir + "/concentration" + post, diag_c, sp->x_lb, sp->y_lb, d_dx, d_dy, sp->c_bounds);
        if (postfix < 0) {
            Project(diag_p, &p, sp->p_bounds, d_N1, d_N2, d_dx, d_dy);
            PrintDToPlotFile(dir + "/pressure" + post, diag_p, sp->x_lb, sp->y_lb, d_dx, d_dy, sp->p_bounds);
        }
    } else {
        PrintDToPlotFile(dir + "/concentration" + post, c, sp->x_lb, sp->y_lb, sp->dx, sp->dy, sp->c_bounds);
        if (postfix < 0) PrintDToPlotFile(dir + "/pressure" + post, p, sp->x_lb, sp->y_lb, sp->dx, sp->dy, sp->p_bounds);
    }
    if (postfix < 0)
        PrintVelToPlotFile(dir + "/velocity_u" + post, dir + "/velocity_v" + post, v, sp->x_lb, sp->y_lb, sp->dx, sp->dy, sp->v_bounds, sp->N1, sp->N2);
}

bool Impec2DSolver::TestComponent() {
    if (myid == 0) cout << "Executing test no. " << sp->testno << " for the ";
    if (sp->test_component == ELLIPTIC) { if (myid == 0) cout << "ELLIPTIC Solver " << endl; return es->TestNeumann(NULL); }
    if (myid == 0) cout << "Hyperbolic Solver " << endl;
    if (sp->testno == 1) return TestAdvection();
    cerr << "Hyperbolic TEST:: Unknown test number: " << sp->testno << endl;
    return false;
}
bool Impec2DSolver::TestAdvection() {
    sp->N1 = sp->init_x_cells; sp->N2 = sp->init_y_cells;
    // several processes: every rank advects its slab, the norm is taken over the gathered field, rank 0 reports
    ofstream norm_stream;
    if (myid == 0) {
        norm_stream.open(sp->norm_file.c_str());
        if (!norm_stream.is_open()) { cerr << "Cannot write norm-file! Convergence test aborted..." << endl; return false; }
        norm_stream.precision(30);
    }
    for (int k = 0; k < sp->levels; k++) {
        sp->dx = (sp->x_ub - sp->x_lb) / sp->N1; sp->dy = (sp->y_ub - sp->y_lb) / sp->N2; sp->M = sp->N1 * sp->N2;
        p.Unbind(); c.Unbind(); v.Unbind();
        ResizeArrays();
        for (int i = 0; i < sp->M; i++) q[i] = 0.;
        if (myid == 0) cout << "->Resolution:\t" << sp->N1 << "  x  " << sp->N2 << flush;
        const double lpnorm = Advection();
        if (myid == 0) {
            cout << "\t\tL" << sp->hyp_norm + 1 << ":\t" << lpnorm << endl;
            norm_stream << log(sp->dx) << "\t" << log(lpnorm) << endl;
        }
        sp->N1 *= 2; sp->N2 *= 2;
    }
    return true;
}
// one period of diagonal periodic advection (src/Impec2DSolver.cpp:848-1015)
double Impec2DSolver::Advection() {
    double (*fptr)(SimParams const *const, double, double) = NULL;
    switch (sp->hyp_test) { case SINE: fptr = Sinewave; break; case TOPHAT: fptr = Tophat; break; case GAUSS: fptr = Gaussbell; break; case HSINE: fptr = VerticalSine; break; }
    for (int i = 0; i < sp->N1; i++) for (int j = 0; j < sp->N2; j++) c(i, j) = fptr(sp, sp->x_lb + (i + 0.5) * sp->dx, sp->y_lb + (j + 0.5) * sp->dy);
    const double ux = -(sp->x_ub - sp->x_lb), uy = (sp->y_ub - sp->y_lb);
    for (int i = -sp->v_bounds; i <= sp->N1 + sp->v_bounds; i++) for (int j = -sp->v_bounds; j <= sp->N2 + sp->v_bounds; j++) { v(i, j, 0) = ux; v(i, j, 1) = uy; }
    const double dt_max = hs->MaxStableAdvectionTimestep(c, v);
    double t = sp->tstart, dt = dt_max;
    size_t tidx = 0;
    if (find(sp->times.begin(), sp->times.end(), sp->tfinal) == sp->times.end()) sp->times.push_back(sp->tfinal);
    sort(sp->times.begin(), sp->times.end());
    const auto t0 = chrono::steady_clock::now();
    while (t < sp->tfinal) {
        if (tidx < sp->times.size() && t + dt >= sp->times[tidx]) { dt = sp->times[tidx] - t; tidx++; } else dt = dt_max;
        if (!hs->Solve(c, v, q, dt)) break;
        t += dt;
    }
    if (myid == 0) cout << "\ttime [ms]: " << (long)(chrono::duration<double, milli>(chrono::steady_clock::now() - t0).count() + 0.5) << flush;
    if (num_procs > 1) DeviceContext::GatherGlobal(DeviceContext::Get(sp), &c, NULL, NULL);
    return sp->hyp_norm == L2 ? CalcAdvectionL2Norm(sp, c, t, fptr) : CalcAdvectionL1Norm(sp, c, t, fptr);
}
