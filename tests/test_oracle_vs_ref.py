"""Pins the portable CPU oracle (oracle/impec_oracle.c) to the REFERENCE ITSELF.

oracle/_ref/libimpec_ref.so holds the reference's own translation units (src/*.cpp of pw359/Impec2D, compiled UNMODIFIED against
the shim headers in oracle/ref_shim/) behind oracle/ref_api.cpp.  Every test feeds the same seeded inputs to the reference's objects
(MGSolver, MUSCLSplitSolver, UpwindSplitSolver, TRIntegrator, Util) and to the oracle and compares:

  * explicit path (ghost fills, wells, mobilities, 9-entry matrix with Neumann folding and pin, right-hand side, Darcy velocity,
    CFL / diffusive time step, TR source step, MUSCL / upwind fluxes, all five splittings, CN matrix and right-hand side,
    parallel->diagonal projection): BIT-EXACT (np.array_equal)
  * solves (pressure, Crank-Nicolson): the reference's HYPRE call is served by the shim's own Krylov solver on the reference's
    recorded matrix, the oracle uses its MG-PCG on its own assembly of the same matrix: relative L2 <= 1e-10
  * multi-step IMPES runs: max|dc| <= 1e-9 over the first steps (round-off of two different solvers amplified by the fingering
    instability afterwards)

The tests are skipped only where the compiled reference is absent (no /root/reference and no prebuilt oracle/_ref).
"""
import ctypes as C
import os
import numpy as np
import pytest
import oracle as O
from oracle import ref as R

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built (needs /root/reference)")
lib = O.lib()
SQRT2 = 1.41421356237
STYPE = {O.PARALLEL: "parallel", O.DIAGONAL: "diagonal", O.HORIZONTAL: "horizontal"}
AVG = {O.HARMONIC: "harmonic", O.MEANCONC: "mean_concentration"}
SPLIT = {O.ALTERNATING: "alternating_sweeps", O.X_FIRST: "x_sweep_first", O.Y_FIRST: "y_sweep_first", O.STRANG: "strang", O.SYMMETRIC_PARALLEL: "parallel"}


def pair(N1, N2, stype=O.PARALLEL, nu=2.0 / 3.0, avg=O.HARMONIC, x_ub=None, y_ub=None, mu_res=41.0, mu_inj=1.0, diffusion=0, alpha_l=1.0,
         alpha_t=1.0, ptype=O.CRANKNICOLSON, split=O.SYMMETRIC_PARALLEL, porosity=1.0, cfl=0.5, is_test=0, testno=4, eps=1e-14,
         well_profile="circle", inj=1.0, dt_diff_limit=0.9, wells=None, endtime=None):
    """The same configuration as oracle parameters and as a set of reference objects (through a generated cfg file)."""
    L = SQRT2 if stype == O.PARALLEL else 1.0
    x_ub = L if x_ub is None else x_ub
    y_ub = L if y_ub is None else y_ub
    Po = O.default_params(N1=N1, N2=N2, x_ub=x_ub, y_ub=y_ub, stype=stype, nu=nu, avgproc=avg, mu_res=mu_res, mu_inj=mu_inj, diffusion=diffusion,
                          alpha_l=alpha_l, alpha_t=alpha_t, ptype=ptype, dim_splitting=split, porosity=porosity, cfl=cfl, is_test=is_test,
                          testno=testno, eps=eps, para_eps=1e-14, inj_fluid_conc=inj, dt_diff_limit=dt_diff_limit)
    ref = R.Ref(general=dict(x_cells=N1, y_cells=N2, x_ub=x_ub, y_ub=y_ub, source_type=STYPE[stype], is_test=bool(is_test)),
                elliptic=dict(mu_res=mu_res, mu_inj=mu_inj, average_procedure=AVG[avg], eps=eps),
                hyperbolic=dict(nu=nu, diffusion=bool(diffusion), alpha_l=alpha_l, alpha_t=alpha_t, dimensional_splitting=SPLIT[split],
                                porosity=porosity, cfl=cfl, inj_fluid_conc=inj, dt_diff_limiter=dt_diff_limit),
                parabolic=dict(para_method="crank-nicolson" if ptype == O.CRANKNICOLSON else "explicit"),
                wells=dict(well_profile=well_profile, **(wells or {})),
                test=dict(component="elliptic", testno=testno, initial_x_cells=N1, initial_y_cells=N2))
    return Po, ref


def synth(Po, seed=20260917):
    c = O.scalar_field(Po)
    lib.orc_synth_state(C.byref(Po), seed, c)
    return c


def rand_velocity(Po, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    u = O.velocity_field(Po); v = O.velocity_field(Po)
    u[:] = scale * (rng.random(u.shape) - 0.5); v[:] = scale * (rng.random(v.shape) - 0.5)
    return u, v


# ----------------------------------------------------------------------------------------------------------------- setup data
def test_cfg_values_reach_the_reference():
    Po, ref = pair(48, 32, stype=O.DIAGONAL, nu=0.5, x_ub=1.0, y_ub=0.75)
    p = ref.params()
    assert p["dx"] == Po.dx and p["dy"] == Po.dy and p["nu"] == 0.5 and p["stype"] == 1 and p["avgproc"] == 1      # src/SimParams.h:16,21
    ref.close()


@pytest.mark.parametrize("stype,profile,code", [(O.PARALLEL, "circle", 0), (O.DIAGONAL, "circle", 0), (O.PARALLEL, "one_cell", 2), (O.DIAGONAL, "one_cell", 2)])
def test_wells_bitexact(stype, profile, code):
    """InitQuarterCircle{Parallel,Diagonal}, InitOneCell* (src/Util.cpp:444-586) through Impec2DSolver::InitializeSource's dispatch."""
    Po, ref = pair(40, 40, stype=stype, well_profile=profile)
    assert ref.L.ref_init_source(ref.h, 0.0) == 0
    q = np.zeros(40 * 40)
    lib.orc_init_source(C.byref(Po), 1.0, 0.05, code, q)
    assert np.array_equal(q, ref.get("q")) and np.abs(q).max() > 0
    ref.close()


def test_exact_vol_frac_circle_bitexact():
    rng = np.random.default_rng(3)
    L = R.lib()
    for _ in range(2000):
        x, y = rng.random(2) * 0.2
        r = 0.02 + rng.random() * 0.15
        dx, dy = 0.005 + rng.random(2) * 0.03
        assert lib.orc_exact_vol_frac_circle(x, y, r, dx, dy) == L.ref_exact_vol_frac_circle(x, y, r, dx, dy)


@pytest.mark.parametrize("is_test", [0, 1])
def test_ghost_fill_bitexact(is_test):
    """FillGhostCellsNeumannScalar / ...PeriodicScalar (src/Util.cpp:215-324) via HyperbolicSolver::FillBoundaries."""
    Po, ref = pair(20, 28, stype=O.DIAGONAL, is_test=is_test)
    c = np.random.default_rng(5).random(O.scalar_field(Po).shape)
    ref.set("c", c); ref.L.ref_fill_boundaries(ref.h)
    co = c.copy(); lib.orc_fill_boundaries(C.byref(Po), co)
    assert np.array_equal(co, ref.get("c"))
    ref.close()


def test_viscosity_bitexact():
    Po, ref = pair(16, 16)
    for c in np.linspace(-0.05, 1.05, 301):
        assert lib.orc_mu(C.byref(Po), c) == ref.L.ref_mu(ref.h, c)
    ref.close()


def test_profiles_bitexact():
    """Sinewave / Tophat / Gaussbell / VerticalSine (src/Util.cpp:55-114), the initial data of hyperbolic_test.cfg; profile
    "horizontal_sine" selects VerticalSine upstream (src/Impec2DSolver.cpp:866-867)."""
    Po, ref = pair(16, 16, stype=O.DIAGONAL, is_test=1)
    rng = np.random.default_rng(11)
    for which, ref_fn in ((O.SINE, 0), (O.TOPHAT, 1), (O.GAUSS, 2), (O.HSINE, 4)):
        for _ in range(200):
            x, y = rng.random(2)
            assert lib.orc_profile(C.byref(Po), which, x, y) == ref.L.ref_profile(ref.h, ref_fn, x, y)
    ref.close()


# ----------------------------------------------------------------------------------------------------------------- a2-a11
ELLIPTIC_CASES = [
    # N1, N2, stype, nu, avg, x_ub, y_ub
    (32, 32, O.PARALLEL, 2.0 / 3.0, O.HARMONIC, None, None),        # parallel.cfg
    (32, 32, O.DIAGONAL, 2.0 / 3.0, O.MEANCONC, None, None),
    (40, 24, O.DIAGONAL, 2.0 / 3.0, O.HARMONIC, 1.0, 1.0),          # dx != dy: rotated-frame cross terms (a6)
    (24, 40, O.DIAGONAL, 0.5, O.MEANCONC, 1.0, 0.8),
    (32, 32, O.DIAGONAL, 1.0, O.HARMONIC, None, None),              # pure 5-point
    (32, 32, O.DIAGONAL, 0.25, O.HARMONIC, None, None),
]


@pytest.mark.parametrize("N1,N2,stype,nu,avg,x_ub,y_ub", ELLIPTIC_CASES)
def test_elliptic_matrix_rhs_solution(N1, N2, stype, nu, avg, x_ub, y_ub):
    """SetMatrixCoefficients_Elliptic + SetRHS_Elliptic (src/MGSolver.cpp:268-590): all 9 entries of every row incl. folded edges,
    corners and the pin, bit for bit; converged pressure to solver tolerance."""
    Po, ref = pair(N1, N2, stype=stype, nu=nu, avg=avg, x_ub=x_ub, y_ub=y_ub)
    c = synth(Po); lib.orc_fill_boundaries(C.byref(Po), c)
    q = np.zeros(N1 * N2); lib.orc_init_source(C.byref(Po), 1.0, 0.05, 0, q)
    ref.set("c", c); ref.set("q", q)
    ref.L.ref_pressure_solve(ref.h)
    A, b, x, rec = ref.last_system()
    assert rec.symmetric == 1 and rec.final_relres < 1e-13
    Ao = np.zeros(N1 * N2 * 9); bo = np.zeros(N1 * N2)
    lib.orc_assemble_elliptic(C.byref(Po), c, q, Ao, bo)
    assert np.array_equal(Ao.reshape(N1, N2, 9), A)
    assert np.array_equal(bo.reshape(N1, N2), b)
    S = O.new_state(); p = O.scalar_field(Po)
    assert lib.orc_solve_elliptic(C.byref(Po), C.byref(S), c, q, p) == 0
    pr = ref.get("p")
    assert np.linalg.norm(O.interior(Po, p) - O.interior(Po, pr)) <= 1e-10 * np.linalg.norm(O.interior(Po, pr))
    ref.close()


@pytest.mark.parametrize("testno,nu", [(1, 1.0), (2, 2.0 / 3.0), (3, 0.5), (4, 2.0 / 3.0)])
def test_elliptic_manufactured_matrix(testno, nu):
    """elliptic_test.cfg operators: K_11, K_22 fields of tests 1-4 (src/EllipticSolver.cpp:418-470) and the test-mode pin
    coefficient kappa = -dA q/p_exact (src/MGSolver.cpp:524-545)."""
    N1, N2 = 20, 30
    Po, ref = pair(N1, N2, stype=O.DIAGONAL, nu=nu, avg=O.MEANCONC, x_ub=2 * np.pi, y_ub=2 * np.pi, mu_res=1.0, mu_inj=1.0, is_test=1, testno=testno, eps=1e-12)
    pex = O.scalar_field(Po); uex = O.velocity_field(Po); vex = O.velocity_field(Po); q = np.zeros(N1 * N2)
    lib.orc_init_elliptic_test(C.byref(Po), pex, q, uex, vex)
    c = O.scalar_field(Po)
    ref.set("c", c); ref.set("q", q)
    ref.L.ref_pressure_solve(ref.h)
    A, b, x, rec = ref.last_system()
    Ao = np.zeros(N1 * N2 * 9); bo = np.zeros(N1 * N2)
    lib.orc_assemble_elliptic(C.byref(Po), c, q, Ao, bo)
    assert np.array_equal(Ao.reshape(N1, N2, 9), A)
    S = O.new_state(); p = O.scalar_field(Po)
    assert lib.orc_solve_elliptic(C.byref(Po), C.byref(S), c, q, p) == 0
    pr = ref.get("p")
    assert np.linalg.norm(O.interior(Po, p) - O.interior(Po, pr)) <= 1e-9 * np.linalg.norm(O.interior(Po, pr))
    # velocities of the manufactured problem from the same pressure
    ref.set("p", p); ref.L.ref_calc_velocities(ref.h)
    u = O.velocity_field(Po); v = O.velocity_field(Po); p2 = p.copy()
    lib.orc_calc_velocities(C.byref(Po), p2, c, u, v)
    ur, vr = ref.get_uv()
    assert np.array_equal(u, ur) and np.array_equal(v, vr)
    ref.close()


# ----------------------------------------------------------------------------------------------------------------- a12-a13
@pytest.mark.parametrize("N1,N2,stype,nu,avg,x_ub,y_ub", ELLIPTIC_CASES)
def test_velocity_and_dt_bitexact(N1, N2, stype, nu, avg, x_ub, y_ub):
    """CalcVelocities incl. the ghost fills of p and v (src/EllipticSolver.cpp:541-683, src/Util.cpp:331-438) and
    MaxStableTimestepCombined (src/HyperbolicSolver.cpp:123-198) on an arbitrary (seeded random) pressure."""
    Po, ref = pair(N1, N2, stype=stype, nu=nu, avg=avg, x_ub=x_ub, y_ub=y_ub)
    c = synth(Po); lib.orc_fill_boundaries(C.byref(Po), c)
    p = np.random.default_rng(9).random(c.shape)
    ref.set("c", c); ref.set("p", p)
    ref.L.ref_calc_velocities(ref.h)
    u = O.velocity_field(Po); v = O.velocity_field(Po); po = p.copy()
    lib.orc_calc_velocities(C.byref(Po), po, c, u, v)
    ur, vr = ref.get_uv()
    assert np.array_equal(u, ur) and np.array_equal(v, vr)
    assert np.array_equal(po, ref.get("p"))                       # p's ghost ring is filled in place (:663)
    assert lib.orc_max_stable_dt(C.byref(Po), u, v) == ref.L.ref_max_stable_dt(ref.h)
    ref.close()


@pytest.mark.parametrize("ptype", [O.EXPLICIT, O.CRANKNICOLSON])
def test_diffusive_time_step_bitexact(ptype):
    """dt_D = lambda min dx^2 / (2 (d11 + d22)) and its combination with dt_A (src/HyperbolicSolver.cpp:157-196): para_method
    explicit returns the minimum, crank-nicolson the advective limit."""
    Po, ref = pair(36, 36, stype=O.DIAGONAL, diffusion=1, ptype=ptype, alpha_l=1.0, alpha_t=0.3, dt_diff_limit=0.7)
    u, v = rand_velocity(Po, 17, scale=4.0)
    ref.set_uv(u, v)
    dto = lib.orc_max_stable_dt(C.byref(Po), u, v)
    assert dto == ref.L.ref_max_stable_dt(ref.h)
    assert lib.orc_max_stable_dt_combined(C.byref(Po), u, v) == ref.L.ref_max_stable_dt_combined(ref.h)
    ref.close()


def test_test_mode_time_step_bitexact():
    Po, ref = pair(30, 30, stype=O.DIAGONAL, is_test=1, cfl=0.95)
    u, v = rand_velocity(Po, 21, scale=2.0)
    ref.set_uv(u, v)
    assert lib.orc_max_stable_dt(C.byref(Po), u, v) == ref.L.ref_max_stable_dt(ref.h)
    assert lib.orc_max_stable_advection_dt(C.byref(Po), u, v) == ref.L.ref_max_stable_dt_advection(ref.h)
    ref.close()


# ----------------------------------------------------------------------------------------------------------------- a14-a18
def developed_velocity(Po, ref, c, q):
    """A physical velocity field: the reference's own pressure + CalcVelocities for this c."""
    ref.set("c", c); ref.set("q", q)
    ref.L.ref_pressure_solve(ref.h); ref.L.ref_calc_velocities(ref.h)
    return ref.get_uv()


def test_source_step_bitexact():
    """TRIntegrator::Integrate + the produced-mass bookkeeping of SourceSplit (src/TRIntegrator.cpp:49-71, src/HyperbolicSolver.cpp:201-205)."""
    Po, ref = pair(32, 32, inj=0.9)
    c = synth(Po); q = np.zeros(32 * 32); lib.orc_init_source(C.byref(Po), 1.0, 0.05, 0, q)
    ref.set("c", c); ref.set("q", q)
    co = c.copy(); tot = 0.0
    for dt in (1e-3, 2.5e-3):
        ref.L.ref_source_split(ref.h, dt, 0)
        tot += abs(lib.orc_tr_integrate(C.byref(Po), co, q, dt))
        assert np.array_equal(co, ref.get("c"))
    assert tot == ref.L.ref_total_mass_outflux(ref.h, 0) and tot > 0
    ref.close()


@pytest.mark.parametrize("method", [O.MUSCL, O.UPWIND])
@pytest.mark.parametrize("porosity", [1.0, 0.35])
def test_intercell_fluxes_bitexact(method, porosity):
    """CalculateIntercellFlux_F/G of MUSCLSplitSolver (src/MUSCLSplitSolver.cpp:102-189) and UpwindSplitSolver (src/UpwindSplitSolver.cpp:43-65)."""
    N1, N2 = 28, 36
    Po, ref = pair(N1, N2, stype=O.DIAGONAL, porosity=porosity)
    Po.hyper_method = method
    c = np.random.default_rng(2).random(O.scalar_field(Po).shape)
    lib.orc_fill_boundaries(C.byref(Po), c)
    u, v = rand_velocity(Po, 4, scale=3.0)
    ref.set("c", c); ref.set_uv(u, v)
    dt = 0.4 * Po.dx / 1.5
    g = Po.g
    F = np.zeros((N1 + 1, N2 + 2 * g)); Fr = np.zeros_like(F)
    G = np.zeros((N1 + 2 * g, N2 + 1)); Gr = np.zeros_like(G)
    lib.orc_flux_F(C.byref(Po), method, c, u, dt, F); ref.L.ref_flux_F(ref.h, dt, method, Fr)
    lib.orc_flux_G(C.byref(Po), method, c, v, dt, G); ref.L.ref_flux_G(ref.h, dt, method, Gr)
    assert np.array_equal(F, Fr) and np.array_equal(G, Gr) and np.abs(F).max() > 0
    ref.close()


@pytest.mark.parametrize("split", [O.SYMMETRIC_PARALLEL, O.ALTERNATING, O.X_FIRST, O.Y_FIRST, O.STRANG])
@pytest.mark.parametrize("method", [O.MUSCL, O.UPWIND])
def test_advect_split_bitexact(split, method):
    """AdvectSplit in all five dimensional splittings, two consecutive calls (ALTERNATING flips its first sweep), on a physical
    velocity field (src/HyperbolicSplitSolver.cpp:51-159)."""
    N = 32
    Po, ref = pair(N, N, stype=O.DIAGONAL, split=split)
    Po.hyper_method = method
    c = synth(Po); lib.orc_fill_boundaries(C.byref(Po), c)
    q = np.zeros(N * N); lib.orc_init_source(C.byref(Po), 1.0, 0.05, 0, q)
    u, v = developed_velocity(Po, ref, c, q)
    dt = lib.orc_max_stable_dt(C.byref(Po), u, v)
    S = O.new_state(); co = c.copy()
    for _ in range(2):
        ref.L.ref_advect_split(ref.h, dt, method)
        lib.orc_advect_split(C.byref(Po), C.byref(S), co, u, v, dt)
        assert np.array_equal(co, ref.get("c"))
    assert np.abs(co - c).max() > 1e-6
    ref.close()


def test_periodic_advection_bitexact():
    """hyperbolic_test.cfg: Sinewave cell values, (u, v) = (-1, +1), periodic ghosts, CFL 0.95, 20 transport steps
    (src/Impec2DSolver.cpp:773-1015)."""
    N = 40
    Po, ref = pair(N, N, stype=O.DIAGONAL, is_test=1, cfl=0.95, x_ub=1.0, y_ub=1.0)
    c = O.scalar_field(Po); g = Po.g
    for i in range(N):
        for j in range(N):
            c[g + i, g + j] = lib.orc_profile(C.byref(Po), O.SINE, (i + 0.5) * Po.dx, (j + 0.5) * Po.dy)
    u = O.velocity_field(Po, -1.0); v = O.velocity_field(Po, 1.0)
    q = np.zeros(N * N)
    ref.set("c", c); ref.set_uv(u, v); ref.set("q", q)
    dt = lib.orc_max_stable_dt(C.byref(Po), u, v)
    assert dt == ref.L.ref_max_stable_dt(ref.h) == 0.95 * Po.dx
    S = O.new_state(); co = c.copy()
    for _ in range(20):
        lib.orc_transport_step(C.byref(Po), C.byref(S), co, u, v, q, dt)
        ref.L.ref_transport_step(ref.h, dt, 0)
    assert np.array_equal(co, ref.get("c"))
    ref.close()


def test_transport_step_bitexact():
    """HyperbolicSplitSolver::Solve = S(dt/2) A(dt) S(dt/2) without dispersion (src/HyperbolicSplitSolver.cpp:162-198), 3 steps."""
    N = 48
    Po, ref = pair(N, N)
    c = synth(Po); lib.orc_fill_boundaries(C.byref(Po), c)
    q = np.zeros(N * N); lib.orc_init_source(C.byref(Po), 1.0, 0.05, 0, q)
    u, v = developed_velocity(Po, ref, c, q)
    dt = lib.orc_max_stable_dt(C.byref(Po), u, v)
    S = O.new_state(); co = c.copy()
    for _ in range(3):
        lib.orc_transport_step(C.byref(Po), C.byref(S), co, u, v, q, dt)
        ref.L.ref_transport_step(ref.h, dt, 0)
        assert np.array_equal(co, ref.get("c"))
    assert S.total_mass_outflux == ref.L.ref_total_mass_outflux(ref.h, 0)
    ref.close()


# ----------------------------------------------------------------------------------------------------------------- a19
@pytest.mark.parametrize("N1,N2,x_ub,y_ub,al,at", [(32, 32, None, None, 1.0, 1.0), (40, 24, 1.0, 1.0, 1.0, 1.0), (32, 32, None, None, 1.0, 0.2)])
def test_crank_nicolson_matrix_rhs_solution(N1, N2, x_ub, y_ub, al, at):
    """SetupDiffusionTensorF/G, GetCoefficients_Parabolic, SetMatrixCoefficients_Parabolic, SetRHS_Parabolic
    (src/MGSolver.cpp:846-948, 997-1371): CN matrix and right-hand side bit for bit (alpha_l != alpha_t included: the 9-point,
    non-symmetric case), solution to solver tolerance."""
    Po, ref = pair(N1, N2, stype=O.DIAGONAL, diffusion=1, x_ub=x_ub, y_ub=y_ub, alpha_l=al, alpha_t=at)
    c = synth(Po); lib.orc_fill_boundaries(C.byref(Po), c)
    q = np.zeros(N1 * N2); lib.orc_init_source(C.byref(Po), 1.0, 0.05, 0, q)
    u, v = developed_velocity(Po, ref, c, q)
    dt = 0.5 * lib.orc_max_stable_dt(C.byref(Po), u, v)
    ref.set("c", c)
    ref.L.ref_diffusion_split(ref.h, dt)
    A, b, x, rec = ref.last_system()
    Ao = np.zeros(N1 * N2 * 9); bo = np.zeros(N1 * N2)
    lib.orc_assemble_parabolic(C.byref(Po), c, u, v, dt, Ao, bo)
    assert np.array_equal(Ao.reshape(N1, N2, 9), A)
    assert np.array_equal(bo.reshape(N1, N2), b)
    assert rec.final_relres < 1e-13
    # (alpha_l != alpha_t: the oracle solves the non-symmetric system with BiCGStab; the reference build applies its PCG regardless)
    S = O.new_state(); co = c.copy()
    assert lib.orc_solve_parabolic(C.byref(Po), C.byref(S), co, u, v, dt) == 0
    cr = ref.get("c")
    assert np.abs(O.interior(Po, co) - O.interior(Po, cr)).max() <= 1e-11
    assert np.abs(O.interior(Po, co) - O.interior(Po, c)).max() > 1e-6           # the half step did something
    ref.close()


# ----------------------------------------------------------------------------------------------------------------- a20
@pytest.mark.parametrize("stype,diffusion,N", [(O.PARALLEL, 0, 48), (O.DIAGONAL, 1, 40)])
def test_impes_steps_match_reference(stype, diffusion, N):
    """The body of the while loop of Impec2DSolver::Solve (src/Impec2DSolver.cpp:343-389) for 4 steps from c = 0: parallel.cfg and
    diagonal_dispersion.cfg operators.  Pressure <= 1e-9 relative, concentration max|dc| <= 1e-9, dt to 1e-10 relative."""
    Po, ref = pair(N, N, stype=stype, diffusion=diffusion)
    q = np.zeros(N * N); lib.orc_init_source(C.byref(Po), 1.0, 0.05, 0, q)
    ref.set("q", q)
    S = O.new_state()
    p = O.scalar_field(Po); c = O.scalar_field(Po); u = O.velocity_field(Po); v = O.velocity_field(Po)
    t = 0.0
    for step in range(4):
        its = C.c_int()
        dtr = ref.L.ref_impes_step(ref.h, t, 0.0, C.byref(its))
        dto = lib.orc_impes_step(C.byref(Po), C.byref(S), p, c, u, v, q, t, 0.0)
        assert abs(dto - dtr) <= 1e-10 * dtr
        pr = ref.get("p"); cr = ref.get("c")
        assert np.linalg.norm(O.interior(Po, p) - O.interior(Po, pr)) <= 1e-9 * np.linalg.norm(O.interior(Po, pr))
        assert np.abs(c - cr).max() <= 1e-9
        t += dtr
    assert c.max() > 0.5
    assert abs(S.total_mass_outflux - ref.L.ref_total_mass_outflux(ref.h, 0)) <= 1e-9 * max(1.0, S.total_mass_outflux)
    ref.close()


# ----------------------------------------------------------------------------------------------------------------- 8f-4: source_type "horizontal"
MODES = {"pulsing": 0, "linear": 1, "constant": 2}


@pytest.mark.parametrize("mode", ["constant", "pulsing", "linear"])
def test_horizontal_front_source_bitexact(mode):
    """InitHorizontalFront (src/Util.cpp:607-639) through Impec2DSolver::InitializeSource's dispatch (:273-274) at several times: layer
    width, base rate, pulsing (second half of every period at q_high_multiple times the rate) and the linear ramp."""
    Po, ref = pair(40, 24, stype=O.HORIZONTAL, x_ub=2.0, y_ub=1.0, wells=dict(mode=mode, well_width=0.1, q_freq=2.0, q_high_multiple=3.0))
    P = ref.params()
    for t in (0.0, 0.13, 0.26, 0.41, 0.9):
        assert ref.L.ref_init_source(ref.h, t) == 0
        q = np.zeros(40 * 24)
        assert lib.orc_init_horizontal_front(C.byref(Po), 1.0, 0.1, MODES[mode], 2.0, 3.0, P["tstart"], P["tfinal"], t, q) == 4
        assert np.array_equal(q, ref.get("q")) and q[0] > 0 and q[-1] < 0, (mode, t)
    ref.close()


@pytest.mark.parametrize("N1,N2,nu", [(32, 32, 2.0 / 3.0), (40, 24, 1.0)])
def test_horizontal_elliptic_matrix_rhs_solution(N1, N2, nu):
    """source_type horizontal: the Dirichlet ROW i = 1 (src/MGSolver.cpp:119-129, 517-567) in the assembled matrix, bit for bit, and the
    converged pressure."""
    Po, ref = pair(N1, N2, stype=O.HORIZONTAL, nu=nu, x_ub=1.0, y_ub=1.0)
    c = synth(Po); lib.orc_fill_boundaries(C.byref(Po), c)
    q = np.zeros(N1 * N2)
    assert lib.orc_init_horizontal_front(C.byref(Po), 1.0, 0.05, 2, 20.0, 50.0, 0.0, 1.0, 0.0, q) >= 1
    ref.set("c", c); ref.set("q", q)
    ref.L.ref_pressure_solve(ref.h)
    A, b, x, rec = ref.last_system()
    Ao = np.zeros(N1 * N2 * 9); bo = np.zeros(N1 * N2)
    lib.orc_assemble_elliptic(C.byref(Po), c, q, Ao, bo)
    assert np.array_equal(Ao.reshape(N1, N2, 9), A) and np.array_equal(bo.reshape(N1, N2), b)
    assert np.count_nonzero(A[1, :, 1:]) == 0 and np.all(A[1, :, 0] > 0)           # row 1: diagonal only
    S = O.new_state(); p = O.scalar_field(Po)
    assert lib.orc_solve_elliptic(C.byref(Po), C.byref(S), c, q, p) == 0
    pr = ref.get("p")
    assert np.linalg.norm(O.interior(Po, p) - O.interior(Po, pr)) <= 1e-10 * np.linalg.norm(O.interior(Po, pr))
    ref.close()


@pytest.mark.parametrize("diffusion", [0, 1])
def test_horizontal_impes_steps_match_reference(diffusion):
    """Three steps of the loop body with the horizontal front (with dispersion: the Dirichlet row i = N1-2 of the Crank-Nicolson problem,
    src/MGSolver.cpp:130-139, 1292-1315, 1357-1364)."""
    N1, N2 = 40, 32
    Po, ref = pair(N1, N2, stype=O.HORIZONTAL, diffusion=diffusion, x_ub=1.0, y_ub=1.0)
    q = np.zeros(N1 * N2)
    assert lib.orc_init_horizontal_front(C.byref(Po), 1.0, 0.05, 2, 20.0, 50.0, 0.0, 1.0, 0.0, q) >= 1
    ref.set("q", q)
    S = O.new_state()
    # Started from the perturbed synthetic state, not from c = 0: with c = 0 and constant injection the flow is exactly one-dimensional, the
    # transverse velocity is zero up to the round-off of the pressure solve, and wherever it comes out EXACTLY zero on a wall face (u = 0
    # there) the reference's dispersion tensor is 0/0 = NaN (src/MGSolver.cpp:868-869 divides by sqrt(u^2 + v^2) unguarded) -- which of the
    # two solvers hits an exact zero depends on the summation order of its reductions, i.e. on the thread count of the run.
    c = synth(Po); ref.set("c", c)
    p = O.scalar_field(Po); u = O.velocity_field(Po); v = O.velocity_field(Po)
    t = 0.0
    for step in range(3):
        its = C.c_int()
        dtr = ref.L.ref_impes_step(ref.h, t, 0.0, C.byref(its))
        dto = lib.orc_impes_step(C.byref(Po), C.byref(S), p, c, u, v, q, t, 0.0)
        assert abs(dto - dtr) <= 1e-10 * dtr
        pr = ref.get("p"); cr = ref.get("c")
        assert np.linalg.norm(O.interior(Po, p) - O.interior(Po, pr)) <= 1e-9 * np.linalg.norm(O.interior(Po, pr))
        assert np.abs(c - cr).max() <= 1e-9
        t += dtr
    assert c.max() > 0.1
    ref.close()


# ----------------------------------------------------------------------------------------------------------------- 8f-3
def test_projection_bitexact():
    """ParallelProjector::FillDiagonal as Impec2DSolver::Project drives it (src/Impec2DSolver.cpp:690-716, src/ParallelProjector.cpp:91-199)."""
    N = 50
    Po, ref = pair(N, N)
    c = synth(Po); lib.orc_fill_boundaries(C.byref(Po), c)
    ref.set("c", c)
    n1 = C.c_int(); n2 = C.c_int()
    lib.orc_project_dims(C.byref(Po), C.byref(n1), C.byref(n2))
    assert n1.value == int(round(N / np.sqrt(2)))
    out = np.zeros((n1.value, n2.value)); outr = np.zeros_like(out)
    lib.orc_project_parallel(C.byref(Po), c, n1.value, n2.value, out)
    ref.L.ref_project(ref.h, 0, n1.value, n2.value, outr)
    assert np.array_equal(out, outr) and out.max() > 0.5
    ref.close()


# ----------------------------------------------------------------------------------------------------------------- whole program
def test_reference_executable_elliptic_test_cfg(tmp_path):
    """BASELINE config 1: the reference's OWN main() (src/Driver.cpp) on examples/elliptic_test.cfg (testno 4, 20x30 doubling, its
    shipped values; 3 levels here) against the oracle's errors of the same manufactured problem.  The norm file holds
    ln(dx), ln(err_p), ln(err_u), ln(err_v) per level (src/EllipticSolver.cpp:242-300)."""
    from test_oracle import elliptic_run
    txt = R.cfg_text(general=dict(x_ub=6.283185307179586476, y_ub=6.283185307179586476, source_type="diagonal", x_cells=50, y_cells=50, is_test=True,
                                  results_dir=str(tmp_path), log_dir=str(tmp_path)),
                     elliptic=dict(eps=1e-12, mu_res=1., mu_inj=1., average_procedure="mean_concentration"),
                     hyperbolic=dict(nu=1., diffusion=True),
                     test=dict(component="elliptic", testno=4, initial_x_cells=20, initial_y_cells=30, profile="tophat", hyp_norm=2, levels=3,
                               norm_file=str(tmp_path / "norm.out"), num_result_file=str(tmp_path / "num.out"), ex_result_file=str(tmp_path / "ex.out"),
                               num_u_file=str(tmp_path / "u.out"), num_v_file=str(tmp_path / "v.out"), ex_u_file=str(tmp_path / "exu.out"),
                               ex_v_file=str(tmp_path / "exv.out")))
    out = R.run_executable(txt, cwd=str(tmp_path))
    assert "SUCCESSFUL" in out
    rows = np.loadtxt(tmp_path / "norm.out")
    assert rows.shape == (3, 4)
    for k in range(3):
        dx, ep, eu, ev, S = elliptic_run(20 * 2 ** k, 30 * 2 ** k, 4, 1.0)
        got = np.exp(rows[k])
        # errors of O(1e-2..1e-4) computed from pressures that agree to ~1e-10 relative
        assert got[1] == pytest.approx(ep, rel=1e-6) and got[2] == pytest.approx(eu, rel=1e-6) and got[3] == pytest.approx(ev, rel=1e-6)
    order = (rows[0, 1] - rows[2, 1]) / (rows[0, 0] - rows[2, 0])
    assert 1.9 < order < 2.05                     # published 2.00 (examples/elliptic.eps:1043)
