// SimParams.cpp -- reads the five shipped configuration files unchanged.  Keys, required/optional status and
// error behaviour (message on cerr, exit(1)) follow src/SimParams.cpp:44-470.
#include "SimParams.h"
#include <iostream>
#include <cstdlib>

static void die(const char *msg) { std::cerr << msg << std::endl; exit(1); }

SimParams::SimParams(const string &configfile) {
    try {
        cfg.readFile(configfile);
        const CfgSetting &app = cfg.getRoot()["application"];
        const CfgSetting &gen = app["general"], &ell = app["elliptic"], &hyp = app["hyperbolic"];
        string sstype, ssync;
        bool ok = gen.lookupValue("x_cells", N1) && gen.lookupValue("y_cells", N2) && gen.lookupValue("num_ghostcells", p_bounds)
                  && gen.lookupValue("num_ghostcells", c_bounds) && gen.lookupValue("x_lb", x_lb) && gen.lookupValue("x_ub", x_ub)
                  && gen.lookupValue("y_ub", y_ub) && gen.lookupValue("y_lb", y_lb) && gen.lookupValue("starttime", tstart)
                  && gen.lookupValue("endtime", tfinal) && gen.lookupValue("I", I) && gen.lookupValue("log_to_file", log_to_file)
                  && gen.lookupValue("results_dir", results_dir) && gen.lookupValue("check_mass_conservation", check_mass_conservation)
                  && gen.lookupValue("check_noflow_boundary", check_noflow_boundary) && gen.lookupValue("is_test", is_test)
                  && gen.lookupValue("source_type", sstype) && gen.lookupValue("gnuplot", gnuplot)
                  && gen.lookupValue("check_divergence", check_divergence) && gen.lookupValue("levelset", levelset)
                  && gen.lookupValue("syncmode", ssync);
        if (!ok) die("Could not find one or more items in section <<general>>!");
        if (ssync == "sync") sm = SYNC; else if (ssync == "async") sm = ASYNC; else die("Config::Synchronization: protocoll is not implemented!");
        if (check_mass_conservation && !gen.lookupValue("max_mass_frac_error", max_mass_frac_error))
            die("Could not find one or more options in section <<general>>! You need to specify <<max_mass_frac_error>>");
        if (log_to_file && !(gen.lookupValue("log_skip_steps", log_skip_steps) && gen.lookupValue("time_matching", time_matching)
                             && gen.lookupValue("log_mass_conservation", log_mass_conservation) && gen.lookupValue("log_dir", log_dir)))
            die("Could not find one or more logging-options in section <<general>>!");
        if (sstype == "parallel") stype = PARALLEL; else if (sstype == "diagonal") stype = DIAGONAL; else if (sstype == "horizontal") stype = HORIZONTAL;
        else die("Config::General: source type is not valid!");

        string smethod, sbtype, savg;
        ok = ell.lookupValue("ell_method", smethod) && ell.lookupValue("bound_type", sbtype) && ell.lookupValue("eps", eps) && ell.lookupValue("mu_res", mu_res)
             && ell.lookupValue("mu_inj", mu_inj) && ell.lookupValue("const_bound_val", b_val) && ell.lookupValue("average_procedure", savg);
        if (!ok) die("Could not find one or more items in <<elliptic>> section!");
        if (sbtype == "neumann") btype = NEUMANN; else die("Config::Elliptic: Boundary type is not supported at the moment!");
        if (smethod == "direct") mtype = DIRECT; else if (smethod == "cg") mtype = CG; else if (smethod == "mg") mtype = MG; else if (smethod == "cg_cuda") mtype = CG_CUDA;
        else die("Config::Elliptic: Method is unknown!");
        if (savg == "harmonic") avgproc = HARMONIC; else if (savg == "mean_concentration") avgproc = MEANCONC; else die("Config::Elliptic: Averaging procedure is unknown!");

        string shm, sds;
        ok = hyp.lookupValue("hyper_method", shm) && hyp.lookupValue("porosity", porosity) && hyp.lookupValue("cfl", cfl) && hyp.lookupValue("inj_fluid_conc", inj_fluid_conc)
             && hyp.lookupValue("nu", nu) && hyp.lookupValue("diffusion", diffusion) && hyp.lookupValue("alpha_l", alpha_l) && hyp.lookupValue("alpha_t", alpha_t)
             && hyp.lookupValue("dt_diff_limiter", dt_diff_limit) && hyp.lookupValue("dimensional_splitting", sds);
        if (!ok) die("Could not find one or more items in <<hyperbolic>> section!");
        if (shm == "waf_split") hmtype = WAFSPLIT; else if (shm == "waf") hmtype = WAF; else if (shm == "ctu") hmtype = CTU; else if (shm == "muscl") hmtype = MUSCL;
        else if (shm == "upwind_split") hmtype = UPWINDSPLIT; else die("Config::Hyperbolic: Method is unknown!");
        if (sds == "alternating_sweeps") dim_splitting = ALTERNATING; else if (sds == "x_sweep_first") dim_splitting = X_FIRST; else if (sds == "y_sweep_first") dim_splitting = Y_FIRST;
        else if (sds == "strang") dim_splitting = STRANG; else if (sds == "parallel") dim_splitting = SYMMETRIC_PARALLEL; else die("Config::Hyperbolic: Dimenstional Splitting specification is unknown!");

        const CfgSetting &cl = gen["log_counter_list"]; for (int i = 0; i < cl.getLength(); ++i) counters.push_back((int)cl[i]);
        const CfgSetting &tl = gen["log_time_list"]; for (int i = 0; i < tl.getLength(); ++i) times.push_back((double)tl[i]);

        if (is_test) {
            const CfgSetting &ts = app["test"];
            string comp, sprofile; int ihyp = 1;
            if (!(ts.lookupValue("component", comp) && ts.lookupValue("testno", testno) && ts.lookupValue("num_result_file", num_result_file)
                  && ts.lookupValue("ex_result_file", ex_result_file) && ts.lookupValue("norm_file", norm_file) && ts.lookupValue("initial_x_cells", init_x_cells)
                  && ts.lookupValue("initial_y_cells", init_y_cells) && ts.lookupValue("num_u_file", num_u_file) && ts.lookupValue("ex_u_file", ex_u_file)
                  && ts.lookupValue("num_v_file", num_v_file) && ts.lookupValue("ex_v_file", ex_v_file) && ts.lookupValue("levels", levels)
                  && ts.lookupValue("profile", sprofile) && ts.lookupValue("hyp_norm", ihyp)))
                std::cerr << "Could not find one or more items in section <<test>>!" << std::endl;
            if (comp == "elliptic") test_component = ELLIPTIC; else if (comp == "hyperbolic") test_component = HYPERBOLIC; else die("TestComponent> Config::General: source type is not valid!");
            if (sprofile == "sine") hyp_test = SINE; else if (sprofile == "tophat") hyp_test = TOPHAT; else if (sprofile == "gaussbell") hyp_test = GAUSS;
            else if (sprofile == "horizontal_sine") hyp_test = HSINE; else die("TestComponent> Config:: Test profile is not valid!");
            if (ihyp == 1) hyp_norm = L1; else if (ihyp == 2) hyp_norm = L2; else die("TestComponent> Config:: Norm is not supported!");
        }
        if (!gen.lookupValue("init_file", init_file)) init_file = "";
        else cout << "Reading initial data from " << init_file << endl;
        string sti;
        if (!(hyp.lookupValue("time_integration", sti) && hyp.lookupValue("ti_eps", ti_eps))) die("Could not find property <<time_integration>> or <<ti_eps>>!");
        ti = TR;

        string swp, smode;
        const CfgSetting &ws = app["wells"];
        if (!(ws.lookupValue("well_profile", swp) && ws.lookupValue("well_rc", well_rc) && ws.lookupValue("q_freq", q_freq) && ws.lookupValue("q_rand_max", q_rand_max)
              && ws.lookupValue("front_const_width", front_const_width) && ws.lookupValue("well_width", well_width) && ws.lookupValue("front_rand_width", front_rand_width)
              && ws.lookupValue("q_high_multiple", q_high_multiple) && ws.lookupValue("mode", smode)))
            die("Could not find one or more items in section <<wells>>!");
        if (smode == "pulsing") im = PULSING; else if (smode == "linear") im = LINEAR; else if (smode == "constant") im = CONSTANT; else die("Unknown injection mode");
        if (swp == "circle") well_profile = CIRCULAR; else if (swp == "gauss") well_profile = GAUSSIAN; else if (swp == "one_cell") well_profile = ONECELL; else die("Unknown profile <<wells>>!");

        string spm;
        const CfgSetting &ps = app["parabolic"];
        if (!(ps.lookupValue("para_method", spm) && ps.lookupValue("eps", para_eps))) die("Could not find one or more items in section <<parabolic>>!");
        if (spm == "crank-nicolson") ptype = CRANKNICOLSON; else if (spm == "explicit") ptype = EXPLICIT; else die("Unknown parabolic solver!");
        if (ptype == CRANKNICOLSON && mtype != MG) die("SimParams::InvalidOptions> The Crank-Nicolson scheme requires a Multigrid-Solver. Change ell_method to MG!");
    } catch (const std::exception &ex) {
        std::cerr << "Parse error: " << ex.what() << std::endl;
        exit(1);
    }
    v_bounds = p_bounds;
    dx = (x_ub - x_lb) / N1;
    dy = (y_ub - y_lb) / N2;
    M = N1 * N2;
    if (stype == PARALLEL && (dx != dy || N1 != N2)) {
        cout << "ERROR: This resolution is not supported for the parallel configuration. The module ParallelProjector assumes square grid blocks and equal resolution in both dimensions." << endl;
        exit(1);
    }
}
SimParams::~SimParams() {}
