"""B200 parity tests proper: the real sm_100a library, called through the C ABI, against the CPU oracle on the
same seeded inputs (sizes the oracle finishes in seconds) plus size-independent properties at larger sizes."""
import ctypes as C
import os
import numpy as np
import pytest
import oracle as O
import impec2d_b200 as I
import parity_cases as pc
from helpers import production_case, advection_case, aos

pytestmark = pytest.mark.gpu


def test_library_is_cuda():
    assert I.load(False).impec_is_emulated() == 0


def test_ghost_fills(): pc.check_ghost_fills(False)
def test_velocity_roundtrip(): pc.check_velocity_roundtrip(False)
def test_single_sweeps_bitexact(): pc.check_single_sweeps_bitexact(False, N=96)
def test_single_sweeps_multitile_bitexact(): pc.check_single_sweeps_bitexact(False, N=700)
def test_intercell_fluxes_bitexact(): pc.check_intercell_fluxes(False, N=200)
def test_sweep_kernels_agree(): pc.check_sweep_variants(False, N=1030)
def test_advection_muscl_bitexact(): pc.check_advection_bitexact(False, N=200, steps=10)
def test_advection_muscl_smooth_bitexact(): pc.check_advection_bitexact(False, N=128, steps=20, perturb=False)
def test_advection_upwind_bitexact(): pc.check_advection_bitexact(False, N=128, steps=5, method=O.UPWIND)
def test_advection_porosity_bitexact(): pc.check_advection_bitexact(False, N=64, steps=3, porosity=0.37)


@pytest.mark.parametrize("split", [O.ALTERNATING, O.X_FIRST, O.Y_FIRST, O.STRANG])
def test_advection_other_splittings_bitexact(split): pc.check_advection_bitexact(False, N=64, steps=3, split=split)


def test_projection_parallel_to_diagonal(): pc.check_projection(False, N=700)
def test_checkpoint_restart(tmp_path): pc.check_checkpoint_restart(False, tmp_path, N=256)
def test_source_step(): pc.check_source_step(False, N=128)
def test_scattered_sources(): pc.check_scattered_sources(False)
def test_fused_vs_unfused_cycle(): pc.check_fused_vs_unfused(False, 1024, 1536)
def test_matvec_march_vs_tiles(): pc.check_matvec_variants(False, 1152, 2064)
def test_matvec_march_odd_columns(): pc.check_matvec_variants(False, 20, 31, eps=1e-9)
def test_march_vs_tile_kernels(): pc.check_march_vs_tiles(False, 1024, 1536)
def test_sgs_march_vs_sweeps(): pc.check_sgs_march(False, 1024, 1536)
def test_fast_initial_residual(): pc.check_fast_initial_residual(False, 1024, 1536)
def test_small_vs_tile_kernels(): pc.check_small_vs_tiles(False, 1024, 1536)


@pytest.mark.parametrize("N1,N2,stype", [(64, 64, O.PARALLEL), (128, 64, O.DIAGONAL), (100, 50, O.DIAGONAL), (141, 141, O.PARALLEL)])
def test_vcycle(N1, N2, stype): pc.check_vcycle(False, N1, N2, stype=stype)


@pytest.mark.parametrize("N1,N2,stype,nu,avg", [
    (128, 128, O.PARALLEL, 2.0 / 3.0, O.HARMONIC),        # parallel.cfg operator
    (128, 64, O.DIAGONAL, 2.0 / 3.0, O.HARMONIC),         # diagonal_dispersion.cfg aspect (dx != dy)
    (96, 96, O.DIAGONAL, 1.0, O.MEANCONC),                # pure 5-point
    (256, 256, O.PARALLEL, 2.0 / 3.0, O.HARMONIC),
])
def test_impes_step(N1, N2, stype, nu, avg): pc.check_impes_step(False, N1, N2, stype=stype, nu=nu, avg=avg, steps=2)


@pytest.mark.parametrize("N1,N2,stype", [(128, 128, O.PARALLEL), (100, 50, O.DIAGONAL), (256, 128, O.DIAGONAL)])
def test_dispersion(N1, N2, stype): pc.check_dispersion(False, N1, N2, stype=stype, steps=2)


def test_elliptic_manufactured_convergence():
    """elliptic_test.cfg on the GPU: testno 4, nu = 1, 20x30 doubling; order 2.00 +- 0.03 on the finest pair."""
    res = [pc.check_elliptic_manufactured(False, 20 * 2 ** k, 30 * 2 ** k, testno=4, nu=1.0) for k in range(5)]
    assert abs(np.log2(res[-2][1] / res[-1][1]) - 2.0) < 0.03


@pytest.mark.parametrize("testno,nu", [(1, 2.0 / 3.0), (2, 2.0 / 3.0), (3, 0.5), (4, 2.0 / 3.0)])
def test_elliptic_manufactured_variants(testno, nu): pc.check_elliptic_manufactured(False, 80, 120, testno=testno, nu=nu)


def test_hyperbolic_convergence_order_on_gpu():
    """hyperbolic_test.cfg on the GPU (50..400 cells): the order is the published 1.85-1.88 and the 50^2 norm
    equals the oracle's."""
    lib = O.lib()
    norms = []
    for k in range(4):
        N = 50 * 2 ** k
        base, Po, c, u, v = advection_case(N)
        Pi = I.default_params(**base)
        ctx = I.Context(Pi)
        ctx.upload_c(c); ctx.upload_q(np.zeros(N * N)); ctx.upload_velocity(aos(u, v))
        dtmax = ctx.max_stable_dt()
        t = 0.0
        while t < 1.0:
            dt = dtmax
            if t + dt >= 1.0: dt = 1.0 - t
            ctx.transport_step(dt); t += dt
        cg = ctx.download_c()
        norms.append((Po.dx, lib.orc_advection_norm(C.byref(Po), cg, O.SINE, 1)))
        ctx.close()
    r = np.array(norms)
    order = np.polyfit(np.log(r[:, 0]), np.log(r[:, 1]), 1)[0]
    assert 1.82 < order < 1.90, order
    assert r[0, 1] == pytest.approx(1.851837e-03, rel=1e-5)      # oracle value at 50^2


def test_large_grid_properties():
    """4096^2 (BASELINE config 2 size): properties that need no oracle run -- mass conservation of the
    advection step with closed walls, residual reduction of the pressure solve, no-flow walls."""
    N = 4096
    base, Po, c, q = production_case(N, N, eps=1e-8)
    Pi = I.default_params(**base)
    ctx = I.Context(Pi)
    ctx.upload_c(c); ctx.upload_q(q)
    dt, info = ctx.impes_step()
    assert info.converged == 1 and info.relres < 1e-8 and info.true_relres < 1e-7
    d0 = ctx.diagnostics()
    assert d0.boundary_velocity_error <= 1e-9 * N
    assert d0.divergence_error < 1e-4
    m0 = d0.total_mass
    ctx.advect_split(dt)
    d1 = ctx.diagnostics()
    assert abs(d1.total_mass - m0) <= 1e-11 * m0            # Neumann walls, zero normal velocity
    assert -1e-12 <= d1.c_min and d1.c_max <= 1.0 + 1e-3
    ctx.close()


def test_headline_size_properties():
    """16384^2 (the size BASELINE's metric is quoted on), oracle-free properties: the register-march multigrid kernels and the
    tile kernels give the SAME cycle bit for bit at full size (every chunk / warp / tile boundary of the production
    launch geometry is exercised), the marching mat-vec solve converges with HYPRE's stop rule, and the advection step conserves
    mass with closed walls."""
    N = 16384
    P = I.default_params(N1=N, N2=N, x_ub=1.41421356237, y_ub=1.41421356237, stype=O.PARALLEL, nu=2.0 / 3.0, eps=1e-10)
    c = I.synth_state(P); q = I.init_source(P)
    ctx = I.Context(P)
    ctx.upload_c(c); ctx.upload_q(q); ctx.fill_boundaries(); ctx.build_operator()
    r = np.random.default_rng(3).standard_normal(N * N)
    ctx.set_option("mg_march", 1); z1 = ctx.vcycle(r)
    ctx.set_option("mg_march", 0); z0 = ctx.vcycle(r)
    ctx.set_option("mg_march", 1)
    assert np.array_equal(z0, z1), "march vs tiles at 16384^2: max diff %g" % np.abs(z0 - z1).max()
    del z0, z1, r
    dt, info = ctx.impes_step()
    assert info.converged == 1 and info.relres < 1e-10 and info.iterations <= 20
    d0 = ctx.diagnostics()
    ctx.advect_split(dt)
    d1 = ctx.diagnostics()
    assert abs(d1.total_mass - d0.total_mass) <= 1e-11 * d0.total_mass
    ctx.close()


@pytest.mark.parametrize("N", [64, 192])
def test_full_fingering_run_integral_diagnostics(N):
    """North-star check 2: a full quarter-five-spot displacement (parallel.cfg operator, M = 41, 64^2 and 192^2, to t = 0.4).
    Pointwise agreement with the oracle over early times (max|dc| <= 1e-8 for t <= 0.05), afterwards the integral
    diagnostics computed by the ONE shared post-processor (impec2d_b200/diagnostics.py):
      breakthrough time +-1 %, produced (recovered) mass +-0.5 %, mass in place +-0.5 %, mixing-zone length +-2 %."""
    from impec2d_b200.diagnostics import RunDiagnostics
    lib = O.lib()
    base, Po, c, q = production_case(N, N, eps=1e-10)
    c[...] = 0.0                                              # the shipped run starts from c = 0 (src/Impec2DSolver.cpp:222-230)
    Pi = I.default_params(**base)
    ctx = I.Context(Pi)
    ctx.upload_c(c); ctx.upload_q(q)
    So = O.new_state()
    p = O.scalar_field(Po); u = O.velocity_field(Po); v = O.velocity_field(Po)
    co = c.copy()
    L = Po.x_ub - Po.x_lb
    dg = RunDiagnostics(q, Po.dx, Po.dy, L, L); do = RunDiagnostics(q, Po.dx, Po.dy, L, L)
    t, tfinal, g = 0.0, 0.4, Po.g
    early_checked = False
    while t < tfinal:
        dto = lib.orc_impes_step(C.byref(Po), C.byref(So), p, co, u, v, q, t, tfinal)
        dtg, info = ctx.impes_step(t, tfinal)
        assert abs(dtg - dto) <= 1e-6 * dto
        t += dto
        cg = ctx.download_c()
        if t <= 0.05:
            assert np.abs(cg - co).max() <= 1e-8, (t, np.abs(cg - co).max())
            early_checked = True
        dg.record(t, cg[g:-g, g:-g]); do.record(t, co[g:-g, g:-g])
    assert early_checked
    sg, so = dg.summary(), do.summary()
    assert so["breakthrough_time"] is not None and 0.05 < so["breakthrough_time"] < 0.4
    assert abs(sg["breakthrough_time"] - so["breakthrough_time"]) <= 0.01 * so["breakthrough_time"]
    assert abs(sg["mixing_zone_length"] - so["mixing_zone_length"]) <= 0.02 * so["mixing_zone_length"]
    assert abs(sg["mass_in_place"] - so["mass_in_place"]) <= 0.005 * so["mass_in_place"]
    d = ctx.diagnostics()
    assert abs(d.total_mass_outflux - So.total_mass_outflux) <= 0.005 * So.total_mass_outflux
    # global mass balance of the run: in place + produced = injected (2 I t for the parallel layout), in cell units
    injected = 2.0 * 1.0 * t / (Po.dx * Po.dy)
    assert abs(d.total_mass + d.total_mass_outflux - injected) <= 1e-6 * injected
    ctx.close()


# ------------------------------------------------------------------------------------------------ round 2
def test_generators_bitexact(): pc.check_generators(False, N=512)
def test_one_cell_wells(): pc.check_one_cell_wells(False, N=128)
def test_explicit_diffusive_dt_bitexact(): pc.check_explicit_diffusive_dt(False, N=300)
def test_diagnostic_values(): pc.check_diagnostic_values(False, N=256)


@pytest.mark.parametrize("N", [2048, 4096])
def test_impes_step_production_kernels_vs_oracle(N):
    """Oracle parity where the production kernels live: default kernel selection, cfg tolerance 1e-14; at 4096^2
    the fine level runs the register-march kernels the bench times (levels of >= 8M cells)."""
    pc.check_impes_step_production(False, N)


def test_dispersion_step_production_kernels_vs_oracle(): pc.check_dispersion_production(False, 2048, 1024)


def test_hyperbolic_test_config_at_4096():
    """BASELINE config 2 at its size: 3 transport steps, bit-exact against the oracle."""
    pc.check_hyperbolic_test_config(False, 4096, steps=3)
def test_integral_diagnostics_device_vs_host(): pc.check_integral_diagnostics(False, N=512, steps=3)
def test_fused_operator_build_bitexact(): pc.check_fused_build(False, 512, 777)
def test_fused_operator_build_bitexact_4096(): pc.check_fused_build(False, 4096, 4096, big=True)


def test_cpp_host_layer_on_the_gpu(tmp_path):
    """The C++ class surface (impec2d_b200/host: EllipticSolver/MGSolver, MUSCLSplitSolver, UpwindSplitSolver ...) linked against the PRODUCT
    library and run on the B200: tests/host_api_check.cpp (fluxes vs sweeps to the bit, two IMPES steps through the classes) and the
    `impec2d` driver on examples/parallel.cfg at 512^2."""
    import os, re, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    host = os.path.join(root, "impec2d_b200", "host"); lib_dir = os.path.join(root, "impec2d_b200")
    subprocess.check_call(["make", "-s", "-C", host, "all"])
    exe = str(tmp_path / "host_api_check")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-I", host, "-o", exe, os.path.join(root, "tests", "host_api_check.cpp"),
                           "-L", lib_dir, "-limpec2d_host", "-limpec2d_b200", "-Wl,-rpath," + lib_dir])
    text = open(os.path.join(root, "examples", "parallel.cfg")).read()
    text = re.sub(r"x_cells\s*=\s*\d+", "x_cells = 512", text); text = re.sub(r"y_cells\s*=\s*\d+", "y_cells = 512", text)
    text = text.replace('"/tmp', '"' + str(tmp_path))
    cfg = tmp_path / "parallel.cfg"; cfg.write_text(text)
    env = dict(os.environ)
    nccl = [os.path.join(p, "nvidia", "nccl", "lib") for p in __import__("sys").path if os.path.exists(os.path.join(p, "nvidia", "nccl", "lib", "libnccl.so.2"))]
    if nccl:
        env["LD_LIBRARY_PATH"] = nccl[0] + ":" + env.get("LD_LIBRARY_PATH", "")
    out = subprocess.run([exe, str(cfg)], capture_output=True, text=True, env=env)
    assert out.returncode == 0 and "host_api_check ok" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
    out = subprocess.run([os.path.join(lib_dir, "impec2d"), str(cfg), "--steps", "3"], capture_output=True, text=True, env=env)
    assert out.returncode == 0 and "[3]\tITS_e:" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


# source_type "horizontal" 
def test_horizontal_front(): pc.check_horizontal_front(False, 256, 192)
def test_horizontal_front_dispersion(): pc.check_horizontal_front(False, 256, 192, diffusion=1)
def test_horizontal_generators(): pc.check_horizontal_generators(False)
def test_horizontal_front_production_kernels(): pc.check_horizontal_front(False, 4096, 2048, diffusion=1, steps=1)      # fine level on the marching kernels


# dispersion with alpha_l != alpha_t (non-symmetric Crank-Nicolson operator)
def test_dispersion_anisotropic_tensor(): pc.check_dispersion(False, 320, 256, stype=O.DIAGONAL, steps=2, alpha_l=1.0, alpha_t=0.2)


# edge cases: odd and tiny grids (an odd extent leaves a single multigrid level; 4 x 4 is the smallest grid the ABI accepts)
@pytest.mark.parametrize("N1,N2,stype", [(30, 21, O.DIAGONAL), (21, 30, O.DIAGONAL), (6, 6, O.PARALLEL), (4, 4, O.DIAGONAL), (50, 36, O.DIAGONAL)])
def test_impes_step_ragged_and_tiny_grids(N1, N2, stype): pc.check_impes_step(False, N1, N2, stype=stype, steps=1)
def test_horizontal_front_tiny_grid(): pc.check_horizontal_front(False, 12, 8, steps=1)


def test_deferred_solution_update_bitexact(): pc.check_deferred_solution_update(False, 1024, 1536)


# an exactly one-dimensional flow with dispersion: faces with zero velocity (0/0 in the reference's dispersion tensor) get no dispersion
def test_one_dimensional_flow_with_dispersion(): pc.check_one_dimensional_flow_with_dispersion(False, 256, 192)
