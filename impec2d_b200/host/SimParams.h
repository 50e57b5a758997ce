// SimParams.h -- all simulation parameters, same public field names and enums as the reference's SimParams
// (src/SimParams.h:13-173), read from the same examples/*.cfg files (keys: src/SimParams.cpp:57-432).
#ifndef SIMPARAMS_H_
#define SIMPARAMS_H_
#include <string>
#include <vector>
#include "ConfigFile.h"
using namespace std;

enum M_Type { DIRECT, CG, CG_CUDA, MG };
enum HM_Type { UPWINDSPLIT, WAFSPLIT, WAF, CTU, MUSCL };
enum S_Type { PARALLEL, DIAGONAL, HORIZONTAL };
enum Solver_Test { ELLIPTIC, HYPERBOLIC };
enum Hyp_Test { SINE, TOPHAT, GAUSS, HSINE };
enum Hyp_Norm { L1, L2 };
enum B_Type { DIRICHLET, SHIFTEDNEUMANN, NEUMANN };
enum DIM_SPLITTING { ALTERNATING, X_FIRST, Y_FIRST, STRANG, SYMMETRIC_PARALLEL };
enum AVG_PROC { MEANCONC, HARMONIC };
enum TIME_INTEGRATOR { TR, HEUN, FE, HEUNFE };
enum WELL { CIRCULAR, GAUSSIAN, ONECELL };
enum P_Type { EXPLICIT, CRANKNICOLSON };
enum SYNCMODE { SYNC, ASYNC };
enum INJ_MODE { PULSING, LINEAR, CONSTANT };

class SimParams {
public:
    ConfigFile cfg;
    double tfinal, tstart;
    int c_bounds, p_bounds, v_bounds;
    int N1, N2, M;
    double x_lb, y_lb, x_ub, y_ub, dx, dy;
    M_Type mtype; P_Type ptype; HM_Type hmtype; S_Type stype;
    double I, porosity;
    bool log_to_file = false, time_matching = false; int log_skip_steps = 0; string log_dir; bool log_delete_content_first = false; string results_dir;
    vector<int> counters; vector<double> times;
    bool log_mass_conservation = false, check_mass_conservation = false, check_noflow_boundary = false; double max_mass_frac_error = 0;
    double inj_fluid_conc, cfl, b_val; B_Type btype; double eps, para_eps, mu_res, mu_inj;
    Solver_Test test_component = ELLIPTIC; int testno = 0;
    string num_result_file, ex_result_file, norm_file, ex_u_file, num_u_file, ex_v_file, num_v_file;
    int levels = 0, init_x_cells = 0, init_y_cells = 0;
    bool is_test = false; Hyp_Test hyp_test = SINE; Hyp_Norm hyp_norm = L1; bool gnuplot = false;
    DIM_SPLITTING dim_splitting; bool check_divergence = false; AVG_PROC avgproc; string init_file;
    TIME_INTEGRATOR ti; double ti_eps; double nu; bool diffusion; double alpha_l, alpha_t;
    bool levelset = false; int ls_bounds = 0, ls_flag_neigh = 0; bool ls_flagged_diff = false; int ls_iter = 0;
    double well_rc; WELL well_profile; double q_freq, q_rand_max; int front_const_width, front_rand_width; double q_high_multiple; INJ_MODE im; double well_width;
    double dt_diff_limit; bool ls_mod_diff = false; double sigma1 = 0, sigma2 = 0, sigma0 = 0; SYNCMODE sm;

    SimParams(const string &configfile);
    virtual ~SimParams();
};
#endif
