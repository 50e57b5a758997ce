"""Parity checks of the C-ABI library against the CPU oracle.  Each function takes `emulated`:
tests/test_emu_parity.py runs them on the CPU emulator (no GPU needed), tests/test_gpu_parity.py runs the
very same checks on the B200 through the real library (-m gpu)."""
import os
import ctypes as C
import numpy as np
import oracle as O
import impec2d_b200 as I
from helpers import production_case, advection_case, aos, interior, rel_l2

# Tolerances (stated here, used below):
#   explicit transport path (fills, source, sweeps, CFL): BIT-EXACT (same operation order, no FMA contraction)
#   pressure / velocity: converged-solution parity, ||p_gpu - p_oracle||_2 / ||p_oracle||_2 <= 1e-9
TOL_P = 1e-9
TOL_VCYCLE = 1e-10


def check_ghost_fills(emulated):
    lib = O.lib()
    for (N1, N2, test) in ((8, 8, 0), (12, 20, 0), (16, 8, 1)):
        Po = O.default_params(N1=N1, N2=N2, is_test=test, testno=1)
        Pi = I.default_params(emulated=emulated, N1=N1, N2=N2, is_test=test, testno=1)
        rng = np.random.default_rng(N1 * 100 + N2)
        c = rng.random((N1 + 4, N2 + 4))
        ctx = I.Context(Pi, emulated=emulated)
        ctx.upload_c(c); ctx.fill_boundaries()
        cg = ctx.download_c()
        lib.orc_fill_boundaries(C.byref(Po), c)
        assert np.array_equal(cg, c), "scalar ghost fill differs (N=%dx%d test=%d)" % (N1, N2, test)
        ctx.close()


def check_velocity_roundtrip(emulated):
    Pi = I.default_params(emulated=emulated, N1=12, N2=8)
    ctx = I.Context(Pi, emulated=emulated)
    v = np.random.default_rng(3).random(ctx.velocity_shape())
    ctx.upload_velocity(v)
    assert np.array_equal(ctx.download_velocity(), v)
    ctx.close()


def check_advection_bitexact(emulated, N=40, steps=4, method=O.MUSCL, split=O.SYMMETRIC_PARALLEL, porosity=1.0, perturb=True):
    lib = O.lib()
    base, Po, c, u, v = advection_case(N, hyper_method=method, dim_splitting=split, porosity=porosity)
    if perturb:   # non-constant velocity and rough data exercise every limiter branch
        rng = np.random.default_rng(7)
        u += 0.4 * rng.random(u.shape); v -= 0.3 * rng.random(v.shape)
        interior(Po, c)[...] += rng.random((N, N)) * (rng.random((N, N)) > 0.5)
    Pi = I.default_params(emulated=emulated, **base)
    ctx = I.Context(Pi, emulated=emulated)
    q = np.zeros(N * N)
    ctx.upload_c(c); ctx.upload_q(q); ctx.upload_velocity(aos(u, v))
    dt_o = lib.orc_max_stable_dt(C.byref(Po), u, v)
    dt_g = ctx.max_stable_dt()
    assert dt_o == dt_g, (dt_o, dt_g)
    So = O.new_state()
    co = c.copy()
    for _ in range(steps):
        lib.orc_transport_step(C.byref(Po), C.byref(So), co, u, v, q, dt_o)
        ctx.transport_step(dt_o)
    cg = ctx.download_c()
    assert np.array_equal(cg, co), "max diff %g" % np.abs(cg - co).max()
    ctx.close()


def check_single_sweeps_bitexact(emulated, N=24):
    lib = O.lib()
    base, Po, c, u, v = advection_case(N)
    rng = np.random.default_rng(11)
    u += 0.4 * rng.random(u.shape); v -= 0.3 * rng.random(v.shape)
    lib.orc_fill_boundaries(C.byref(Po), c)
    Pi = I.default_params(emulated=emulated, **base)
    ctx = I.Context(Pi, emulated=emulated)
    ctx.upload_velocity(aos(u, v)); ctx.upload_q(np.zeros(N * N))
    for fn_o, fn_g, vel in ((lib.orc_advect_x, ctx.advect_x, u), (lib.orc_advect_y, ctx.advect_y, v)):
        co = c.copy()
        ctx.upload_c(c)
        fn_o(C.byref(Po), co, vel, 0.01)
        fn_g(0.01)
        assert np.array_equal(ctx.download_c(), co)
    ctx.close()


def check_intercell_fluxes(emulated, N=24):
    """CalculateIntercellFlux_F / _G as fields (src/HyperbolicSplitSolver.h:19-22): bit-exact against the oracle for MUSCL and upwind."""
    lib = O.lib()
    for method in (O.MUSCL, O.UPWIND):
        base, Po, c, u, v = advection_case(N, hyper_method=method)
        rng = np.random.default_rng(13)
        u += 0.4 * rng.random(u.shape); v -= 0.3 * rng.random(v.shape)
        lib.orc_fill_boundaries(C.byref(Po), c)
        Pi = I.default_params(emulated=emulated, **base)
        ctx = I.Context(Pi, emulated=emulated)
        ctx.upload_velocity(aos(u, v)); ctx.upload_q(np.zeros(N * N)); ctx.upload_c(c)
        g = Po.g
        F = np.zeros((N + 1, N + 2 * g)); G = np.zeros((N + 2 * g, N + 1))
        lib.orc_flux_F(C.byref(Po), method, c, u, 0.01, F)
        lib.orc_flux_G(C.byref(Po), method, c, v, 0.01, G)
        assert np.array_equal(ctx.intercell_flux(0, 0.01), F)
        assert np.array_equal(ctx.intercell_flux(1, 0.01), G)
        assert np.abs(F).max() > 0 and np.abs(G).max() > 0
        ctx.close()


def check_sweep_variants(emulated, N=150):
    """k_muscl_x / k_muscl_y (one reconstruction per cell) against k_sweep (three per cell): bit-identical, any size
    (no oracle needed, so this also runs at sizes the oracle would be slow on)."""
    base, Po, c, u, v = advection_case(N)
    rng = np.random.default_rng(13)
    u += 0.4 * rng.random(u.shape); v -= 0.3 * rng.random(v.shape)
    interior(Po, c)[...] += rng.random((N, N)) * (rng.random((N, N)) > 0.5)
    Pi = I.default_params(emulated=emulated, **base)
    ctx = I.Context(Pi, emulated=emulated)
    ctx.upload_velocity(aos(u, v)); ctx.upload_q(np.zeros(N * N))
    out = {}
    for shared in (1, 0):
        ctx.set_option("sweep_shared", shared)
        ctx.upload_c(c)
        ctx.transport_step(0.4 / N)
        ctx.advect_x(0.2 / N); ctx.advect_y(0.2 / N)
        out[shared] = ctx.download_c()
    ctx.set_option("sweep_shared", 1)
    assert np.array_equal(out[0], out[1]), "max diff %g" % np.abs(out[0] - out[1]).max()
    ctx.close()


def check_projection(emulated, N=48):
    """Impec2DSolver::Project / ParallelProjector -- the parallel field on the diagonal grid of
    round(N / sqrt 2) cells.  Same operation order as the oracle's restatement: BIT-EXACT, for c and for p."""
    lib = O.lib()
    base, Po, c, q = production_case(N, N, stype=O.PARALLEL)
    Pi = I.default_params(emulated=emulated, **base)
    ctx = I.Context(Pi, emulated=emulated)
    n1 = C.c_int(); n2 = C.c_int()
    lib.orc_project_dims(C.byref(Po), C.byref(n1), C.byref(n2))
    assert n1.value == int(round(N / np.sqrt(2))) and n2.value == n1.value
    rng = np.random.default_rng(21)
    for field, up in ((0, ctx.upload_c), (1, ctx.upload_p)):
        a = O.scalar_field(Po)
        interior(Po, a)[...] = rng.random((N, N))
        lib.orc_fill_boundaries(C.byref(Po), a)               # Neumann mirror, as both fields carry at output time
        up(a)
        ref = np.zeros((n1.value, n2.value))
        lib.orc_project_parallel(C.byref(Po), a, n1.value, n2.value, ref)
        got = ctx.project_to_diagonal(field)
        assert got.shape == ref.shape
        assert np.array_equal(got, ref), "max diff %g" % np.abs(got - ref).max()
    # a linear field x + y is reproduced exactly by bilinear interpolation: p_x + |p_y| = sqrt(2) * max(x_d, y_d) on the diagonal grid
    a = O.scalar_field(Po)
    dx = (Po.x_ub - Po.x_lb) / N
    xs = (np.arange(-Po.g, N + Po.g) + 0.5) * dx
    a[...] = xs[:, None] + xs[None, :]
    ctx.upload_c(a)
    got = ctx.project_to_diagonal(0)
    d_dx = (Po.x_ub - Po.x_lb) / (np.sqrt(2.) * n1.value)
    xd = (np.arange(n1.value) + 0.5) * d_dx
    assert np.abs(got - np.sqrt(2.) * np.maximum(xd[:, None], xd[None, :])).max() < 1e-12
    ctx.close()


def check_checkpoint_restart(emulated, tmpdir, N=32, split=O.ALTERNATING):
    """save after 2 IMPES steps, continue 2 more; a fresh context that loads the checkpoint and runs the same
    2 steps must agree BIT FOR BIT (c, p, produced mass; alternating sweeps exercise the restored sweep parity)."""
    base, Po, c, q = production_case(N, N, stype=O.PARALLEL, dim_splitting=split)
    Pi = I.default_params(emulated=emulated, **base)
    path = os.path.join(str(tmpdir), "state.ckpt")
    ctx = I.Context(Pi, emulated=emulated)
    ctx.upload_c(c); ctx.upload_q(q)
    t = 0.0
    for _ in range(2):
        dt, _i = ctx.impes_step(t, 0.0); t += dt
    ctx.checkpoint_save(path, t, 2)
    for _ in range(2):
        dt, _i = ctx.impes_step(t, 0.0); t += dt
    c_ref = ctx.download_c(); p_ref = ctx.download_p(); m_ref = ctx.diagnostics().total_mass_outflux
    ctx.close()
    ctx = I.Context(Pi, emulated=emulated)
    ctx.upload_q(q)
    t2, step = ctx.checkpoint_load(path)
    assert step == 2 and t2 > 0.
    for _ in range(2):
        dt, _i = ctx.impes_step(t2, 0.0); t2 += dt
    assert t2 == t
    assert np.array_equal(ctx.download_c(), c_ref) and np.array_equal(ctx.download_p(), p_ref)
    assert ctx.diagnostics().total_mass_outflux == m_ref
    ctx.close()
    # a checkpoint of another grid is refused loudly
    P2 = I.default_params(emulated=emulated, **dict(base, N1=N // 2, N2=N // 2))
    ctx = I.Context(P2, emulated=emulated)
    try:
        ctx.checkpoint_load(path)
        accepted = True
    except I.ImpecError:
        accepted = False
    assert not accepted, "mismatched checkpoint accepted"
    ctx.close()


def check_source_step(emulated, N=32):
    lib = O.lib()
    base, Po, c, q = production_case(N)
    Pi = I.default_params(emulated=emulated, **base)
    ctx = I.Context(Pi, emulated=emulated)
    interior(Po, c)[...] = np.random.default_rng(5).random((N, N))
    ctx.upload_c(c); ctx.upload_q(q)
    co = c.copy()
    md = 0.0
    for dt in (1e-3, 2.5e-3):
        md += abs(lib.orc_tr_integrate(C.byref(Po), co, q, dt))
        ctx.source_split(dt)
    assert np.array_equal(ctx.download_c(), co)
    d = ctx.diagnostics()
    assert abs(d.total_mass_outflux - md) <= 1e-13 * abs(md)
    ctx.close()


def check_scattered_sources(emulated, N1=70, N2=45):
    """impec_upload_q: the dense host array is scanned in 16-row chunks / 32-value blocks and only its non-zero cells are copied.
    Sources scattered over chunk and block edges, a ragged row length, +-0.0 entries, negative rates, and three uploads of
    different sizes in a row (staging buffers regrown and reused): every source step must equal the oracle's bit for bit."""
    lib = O.lib()
    base, Po, c, q0 = production_case(N1, N2)
    Pi = I.default_params(emulated=emulated, **base)
    ctx = I.Context(Pi, emulated=emulated)
    rng = np.random.default_rng(11)
    interior(Po, c)[...] = rng.random((N1, N2))
    ctx.upload_c(c)
    co = c.copy()
    md = 0.0
    for density in (0.002, 0.05, 0.0005):
        q = np.zeros((N1, N2))
        mask = rng.random((N1, N2)) < density
        mask[[0, 15, 16, 31, 32, N1 - 1], :] |= rng.random((6, N2)) < 0.2          # chunk edges
        mask[:, [0, 31, 32, N2 - 1]] |= rng.random((N1, 4)) < 0.2                  # block edges and the ragged tail
        q[mask] = rng.standard_normal(int(mask.sum()))
        q[~mask & (rng.random((N1, N2)) < 0.3)] = -0.0                             # negative zeros are not sources
        qf = np.ascontiguousarray(q.reshape(-1))
        ctx.upload_q(qf)
        md += abs(lib.orc_tr_integrate(C.byref(Po), co, qf, 1e-3))
        ctx.source_split(1e-3)
        assert np.array_equal(ctx.download_c(), co), density
    assert abs(ctx.diagnostics().total_mass_outflux - md) <= 1e-13 * abs(md)
    ctx.upload_q(np.zeros(N1 * N2))                                                # no source at all
    ctx.source_split(1e-3)
    assert np.array_equal(ctx.download_c(), co)
    ctx.close()


def check_vcycle(emulated, N1=32, N2=32, stype=O.PARALLEL, nu=2.0 / 3.0):
    lib = O.lib()
    base, Po, c, q = production_case(N1, N2, stype=stype, nu=nu)
    Pi = I.default_params(emulated=emulated, **base)
    ctx = I.Context(Pi, emulated=emulated)
    lib.orc_fill_boundaries(C.byref(Po), c)
    A = np.zeros(N1 * N2 * 9); b = np.zeros(N1 * N2)
    lib.orc_assemble_elliptic(C.byref(Po), c, q, A, b)
    r = np.random.default_rng(1).standard_normal(N1 * N2)
    zo = np.zeros(N1 * N2)
    assert lib.orc_mg_apply(C.byref(Po), N1, N2, A, -1, r, zo) == 0
    ctx.upload_c(c); ctx.upload_q(q)
    # the cycle kernels normally read FP32 copies of the coefficients (a 6e-8 relative perturbation of the
    # preconditioner, none of the solution): strict parity with FP64 coefficients, loose parity with the default
    ctx.set_option("mg_links_fp32", 0); ctx.build_operator()
    zg = ctx.vcycle(r)
    assert np.abs(zg - zo).max() <= TOL_VCYCLE * np.abs(zo).max()
    ctx.set_option("mg_links_fp32", 1); ctx.build_operator()
    zg = ctx.vcycle(r)
    assert np.abs(zg - zo).max() <= 1e-5 * np.abs(zo).max()
    ctx.close()


def check_impes_step(emulated, N1=48, N2=48, stype=O.PARALLEL, nu=2.0 / 3.0, avg=O.HARMONIC, eps=1e-12, steps=2, its_tol=1):
    """One and two full IMPES steps (a20): pressure to TOL_P, velocities/dt/c to the tolerance the pressure
    tolerance implies, same iteration count +-1 (its_tol = None: not compared -- with nu < 1/2 on cells of aspect ratio > 2 the
    rotated-frame operator is far from an M-matrix, the point-smoother cycle is a weak preconditioner (60-80 iterations on both sides)
    and the count moves by several iterations with the summation order of the dot products or the FP32 coefficient copies)."""
    lib = O.lib()
    base, Po, c, q = production_case(N1, N2, stype=stype, nu=nu, avgproc=avg, eps=eps)
    Pi = I.default_params(emulated=emulated, **base)
    ctx = I.Context(Pi, emulated=emulated)
    ctx.upload_c(c); ctx.upload_q(q)
    So = O.new_state()
    p = O.scalar_field(Po); u = O.velocity_field(Po); v = O.velocity_field(Po)
    co = c.copy()
    t = 0.0
    for _ in range(steps):
        dto = lib.orc_impes_step(C.byref(Po), C.byref(So), p, co, u, v, q, t, 0.0)
        dtg, info = ctx.impes_step(t, 0.0)
        assert info.converged == 1 and info.relres < eps
        assert info.true_relres < 50 * eps
        assert its_tol is None or abs(info.iterations - So.last_its) <= its_tol
        assert abs(dtg - dto) <= 1e-9 * dto
        t += dto
    pg = ctx.download_p(); cg = ctx.download_c(); vg = ctx.download_velocity()
    assert rel_l2(interior(Po, pg), interior(Po, p)) <= TOL_P
    assert np.abs(vg[..., 0] - u).max() <= 1e-8 * np.abs(u).max()
    assert np.abs(vg[..., 1] - v).max() <= 1e-8 * np.abs(v).max()
    assert np.abs(cg - co).max() <= 1e-8
    d = ctx.diagnostics()
    assert abs(d.total_mass - lib.orc_total_mass(C.byref(Po), co)) <= 1e-9 * d.total_mass
    assert abs(d.total_mass_outflux - So.total_mass_outflux) <= 1e-8 * max(So.total_mass_outflux, 1e-30)
    ctx.close()


def check_horizontal_front(emulated, N1=48, N2=32, diffusion=0, steps=2, links_fp32=1):
    """source_type "horizontal" (SURVEY 8f rank 4): the Dirichlet ROW i = 1 of the pressure problem (src/MGSolver.cpp:119-140) and, with
    dispersion, the Dirichlet row i = N1-2 of the Crank-Nicolson problem (:130-139, 1292-1315, 1357-1364), against the oracle."""
    lib = O.lib()
    base, Po, c, q = production_case(N1, N2, stype=O.HORIZONTAL, diffusion=diffusion, eps=1e-12)
    Pi = I.default_params(emulated=emulated, **base)
    ctx = I.Context(Pi, emulated=emulated)
    ctx.set_option("mg_links_fp32", links_fp32)
    ctx.upload_c(c); ctx.upload_q(q)
    So = O.new_state(); p = O.scalar_field(Po); u = O.velocity_field(Po); v = O.velocity_field(Po); co = c.copy()
    t = 0.0
    for _ in range(steps):
        dto = lib.orc_impes_step(C.byref(Po), C.byref(So), p, co, u, v, q, t, 0.0)
        dtg, info = ctx.impes_step(t, 0.0)
        # With FP64 coefficients in the cycle the iteration counts (and residual histories) equal the oracle's; the default FP32 copies of
        # the coefficients perturb the preconditioner by 6e-8, which costs 0-2 iterations here (the coarse cells next to the Dirichlet row
        # carry a large diagonal shift) and nothing in the converged pressure.
        assert info.converged == 1 and abs(info.iterations - So.last_its) <= (2 if links_fp32 else 0), (info.iterations, So.last_its)
        assert abs(dtg - dto) <= 1e-9 * dto
        if diffusion:
            assert abs(ctx.last_parabolic_info().iterations - So.last_parabolic_its) <= 1
        t += dto
    g = Po.g
    pg = ctx.download_p(); cg = ctx.download_c()
    assert rel_l2(interior(Po, pg), interior(Po, p)) <= TOL_P
    assert np.abs(cg - co).max() <= 1e-8
    # the pinned row holds its Dirichlet values exactly: p(1, j) = -q(1, j) dA, and (with dispersion) c(N1-2, j) = 0 after the last CN half step
    dA = Po.dx * Po.dy
    assert np.array_equal(pg[g + 1, g:-g], -q.reshape(N1, N2)[1] * dA)
    if diffusion:
        assert np.all(cg[g + N1 - 2, g:-g] == 0.0)
    ctx.close()


def check_one_dimensional_flow_with_dispersion(emulated, N1=48, N2=32, steps=3):
    """c = 0, horizontal front with constant injection, dispersion on: an exactly one-dimensional flow.  The transverse velocity is zero up
    to the round-off of the pressure solve; wherever both velocity components of a face are EXACTLY zero the reference's dispersion tensor is
    0/0 (src/MGSolver.cpp:868-869, 904-905) and its Crank-Nicolson matrix NaN.  The product gives such a face no dispersion (k_disp_links):
    every field stays finite, the solution stays one-dimensional to round-off and mass is injected."""
    base, Po, c, q = production_case(N1, N2, stype=O.HORIZONTAL, diffusion=1, eps=1e-12)
    c[...] = 0.0
    Pi = I.default_params(emulated=emulated, **base)
    ctx = I.Context(Pi, emulated=emulated)
    ctx.upload_c(c); ctx.upload_q(q)
    t = 0.0
    for _ in range(steps):
        dt, info = ctx.impes_step(t, 0.0)
        assert info.converged == 1 and np.isfinite(dt) and dt > 0
        assert ctx.last_parabolic_info().converged == 1
        t += dt
    g = Po.g
    cg = ctx.download_c()[g:-g, g:-g]; pg = ctx.download_p()[g:-g, g:-g]
    assert np.isfinite(cg).all() and np.isfinite(pg).all()
    assert cg.max() > 0.05 and cg.min() >= -1e-12
    assert np.abs(cg - cg[:, :1]).max() <= 1e-9                     # still a function of x alone
    assert np.all(cg[N1 - 2] == 0.0)                                # the Dirichlet row of the Crank-Nicolson problem
    ctx.close()


def check_deferred_solution_update(emulated, N1=48, N2=80, stype=O.DIAGONAL, diffusion=1):
    """PCG with the solution update of two iterations merged (k_update_xr MODE 1 / 2, option "defer_x") against the plain form: the
    same operations on x in the same order, so pressure and concentration are BIT-IDENTICAL, for even and odd iteration counts."""
    base, Po, c, q = production_case(N1, N2, stype=stype, diffusion=diffusion)
    out = {}
    for eps in (1e-6, 1e-9, 1e-12):                      # different tolerances: the solve ends on either parity of the iteration count
        for defer in (1, 0):
            Pi = I.default_params(emulated=emulated, **dict(base, eps=eps))
            ctx = I.Context(Pi, emulated=emulated)
            ctx.set_option("defer_x", defer)
            ctx.upload_c(c); ctx.upload_q(q)
            dt, info = ctx.impes_step(0.0, 0.0)
            out[(eps, defer)] = (ctx.download_p(), ctx.download_c(), info.iterations)
            ctx.close()
        a, b = out[(eps, 1)], out[(eps, 0)]
        assert a[2] == b[2] and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (eps, a[2], b[2], np.abs(a[0] - b[0]).max())


def check_horizontal_generators(emulated):
    """InitHorizontalFront with constant, pulsing and linearly ramped injection (src/Util.cpp:607-639): library == oracle bit for bit."""
    lib = O.lib()
    base = dict(N1=40, N2=24, x_ub=2.0, y_ub=1.0, stype=O.HORIZONTAL)
    Po = O.default_params(**base); Pi = I.default_params(emulated=emulated, **base)
    for mode in (0, 1, 2):
        for t in (0.0, 0.13, 0.26, 0.41, 0.9):
            qo = np.zeros(40 * 24)
            assert lib.orc_init_horizontal_front(C.byref(Po), 1.5, 0.1, mode, 2.0, 3.0, 0.0, 1.0, t, qo) == 4
            qg = I.init_horizontal_front(Pi, 1.5, 0.1, mode, 2.0, 3.0, 0.0, 1.0, t, emulated=emulated)
            assert np.array_equal(qo, qg) and abs(qo.sum()) < 1e-12 and qo[0] > 0 and qo[-1] < 0
    try:
        I.init_horizontal_front(Pi, 1.0, 0.001, 2, emulated=emulated)         # layer narrower than one cell: the reference exits
        assert False
    except I.ImpecError:
        pass


def check_elliptic_manufactured(emulated, N1=20, N2=30, testno=4, nu=1.0):
    """elliptic_test.cfg: p = cos x cos y, Neumann, K(x,y) test fields; GPU pressure vs oracle and vs exact."""
    lib = O.lib()
    kw = dict(N1=N1, N2=N2, x_ub=2 * np.pi, y_ub=2 * np.pi, mu_res=1.0, mu_inj=1.0, avgproc=O.MEANCONC, nu=nu,
              is_test=1, testno=testno, stype=O.DIAGONAL, eps=1e-12)
    Po = O.default_params(**kw); Pi = I.default_params(emulated=emulated, **kw)
    c = O.scalar_field(Po); p = O.scalar_field(Po); pex = O.scalar_field(Po)
    u = O.velocity_field(Po); v = O.velocity_field(Po); uex = O.velocity_field(Po); vex = O.velocity_field(Po)
    q = np.zeros(N1 * N2)
    lib.orc_init_elliptic_test(C.byref(Po), pex, q, uex, vex)
    So = O.new_state()
    lib.orc_solve_elliptic(C.byref(Po), C.byref(So), c, q, p)
    lib.orc_calc_velocities(C.byref(Po), p, c, u, v)
    ctx = I.Context(Pi, emulated=emulated)
    ctx.upload_c(c); ctx.upload_q(q)
    # Neumann ghosts although is_test: TestNeumann fills c by hand (src/EllipticSolver.cpp:231)
    info = ctx.pressure_solve()
    ctx.calc_velocities()
    pg = ctx.download_p(); vg = ctx.download_velocity()
    assert info.converged == 1
    assert rel_l2(interior(Po, pg), interior(Po, p)) <= TOL_P
    g = Po.g
    assert np.abs(vg[g:g + N1 + 1, g:g + N2, 0] - u[g:g + N1 + 1, g:g + N2]).max() <= 1e-8
    assert np.abs(vg[g:g + N1, g:g + N2 + 1, 1] - v[g:g + N1, g:g + N2 + 1]).max() <= 1e-8
    err_g = np.abs(interior(Po, pg) - interior(Po, pex)).sum() * Po.dx * Po.dy
    err_o = np.abs(interior(Po, p) - interior(Po, pex)).sum() * Po.dx * Po.dy
    assert abs(err_g - err_o) <= 1e-6 * err_o
    ctx.close()
    return Po.dx, err_g


def check_fused_vs_unfused(emulated, N1=256, N2=384, stype=O.DIAGONAL):
    """The fused tile kernels (kernels_fused.inl) against the one-kernel-per-colour path on the same operator:
    equal up to the rounding terms the fused residual drops (<= 1e-12 relative).  Grid large enough to contain
    check-free interior tiles as well as edge tiles."""
    base, Po, c, q = production_case(N1, N2, stype=stype)
    Pi = I.default_params(emulated=emulated, **base)
    r = np.random.default_rng(1).standard_normal(N1 * N2)
    out = []
    # tile heights of the fused kernels: all levels 64 rows / all 32 / all 16 / the default per-level choice
    policies = ((0, None, None), (1, "0", "0"), (1, "1e30", "0"), (1, "1e30", "1e30"), (1, None, None))
    try:
        for fused, th32, th16 in policies:
            for key, val in (("IMPEC_TH32_BELOW", th32), ("IMPEC_TH16_BELOW", th16)):
                if val is None: os.environ.pop(key, None)
                else: os.environ[key] = val
            ctx = I.Context(Pi, emulated=emulated)
            ctx.set_option("mg_fused", fused); ctx.set_option("mg_links_fp32", 0); ctx.set_option("mg_small", 0)
            ctx.upload_c(c); ctx.upload_q(q); ctx.fill_boundaries(); ctx.build_operator()
            z1 = ctx.vcycle(r); z2 = ctx.vcycle(r)
            assert np.array_equal(z1, z2), "cycle is not repeatable (race?)"
            out.append(z1)
            ctx.close()
    finally:
        os.environ.pop("IMPEC_TH32_BELOW", None); os.environ.pop("IMPEC_TH16_BELOW", None)
    assert np.abs(out[0] - out[1]).max() <= 1e-12 * np.abs(out[0]).max()
    # identical arithmetic per cell: the result does not depend on how the level is cut into tiles
    for z in out[2:]:
        assert np.array_equal(z, out[1]), "tile height changes the result"


def check_matvec_variants(emulated, N1=48, N2=272, stype=O.DIAGONAL, eps=1e-5):
    """k_pcg_matvec_march (register march, warp shuffles) against k_pcg_matvec (shared-memory tiles) inside a full
    pressure solve: identical arithmetic per cell, only the order of the dot-product partial sums differs, so the
    iteration count is the same and the pressures agree to rounding (<= 1e-12 relative).  Several row chunks and
    column strips, a partial last warp; callers add a small odd-N2 case (single level) for the unpaired last column."""
    base, Po, c, q = production_case(N1, N2, stype=stype, eps=eps)
    Pi = I.default_params(emulated=emulated, **base)
    res = []
    for march in (1, 0):
        ctx = I.Context(Pi, emulated=emulated)
        ctx.set_option("matvec_march", march)
        ctx.upload_c(c); ctx.upload_q(q); ctx.fill_boundaries()
        info = ctx.pressure_solve()
        assert info.converged == 1
        res.append((info.iterations, ctx.download_p()))
        ctx.close()
    assert res[0][0] == res[1][0], (res[0][0], res[1][0])
    assert rel_l2(interior(Po, res[0][1]), interior(Po, res[1][1])) <= 1e-12


def check_march_vs_tiles(emulated, N1=256, N2=384, stype=O.DIAGONAL):
    """The register-march multigrid kernels (kernels_march.inl) against the shared-memory tile kernels: the same
    arithmetic per cell in the same order, so one cycle must agree BIT FOR BIT -- with FP64 and FP32 coefficients, with
    the march forced onto every level that can run it (IMPEC_MARCH_MIN_CELLS = 0; chunks of 32 rows, several warps
    and CTAs per row, a partial last warp)."""
    base, Po, c, q = production_case(N1, N2, stype=stype)
    Pi = I.default_params(emulated=emulated, **base)
    r = np.random.default_rng(4).standard_normal(N1 * N2)
    os.environ["IMPEC_MARCH_MIN_CELLS"] = "0"
    try:
        for f32 in (0, 1):
            out = []
            for march in (0, 1):
                ctx = I.Context(Pi, emulated=emulated)
                ctx.set_option("mg_march", march); ctx.set_option("mg_links_fp32", f32)
                ctx.upload_c(c); ctx.upload_q(q); ctx.fill_boundaries(); ctx.build_operator()
                z1 = ctx.vcycle(r); z2 = ctx.vcycle(r)
                assert np.array_equal(z1, z2), "cycle is not repeatable (race?)"
                out.append(z1)
                ctx.close()
            assert np.array_equal(out[0], out[1]), "march vs tiles (fp32 links %d): max diff %g of %g" % (f32, np.abs(out[0] - out[1]).max(), np.abs(out[0]).max())
    finally:
        os.environ.pop("IMPEC_MARCH_MIN_CELLS", None)


def check_sgs_march(emulated, N1=64, N2=96, stype=O.DIAGONAL):
    """The Crank-Nicolson preconditioner (one symmetric red-black sweep from zero) as ONE march (k_sgs_march) against the
    same sweep issued as memset + forward PLAIN + backward PLAIN: identical arithmetic per cell, so the dispersion half
    step gives the same concentration and the same iteration count (the fused <r,z> is summed in another order: 1e-13)."""
    lib = O.lib()
    base, Po, c, q = production_case(N1, N2, stype=stype, diffusion=1)
    Pi = I.default_params(emulated=emulated, **base)
    os.environ["IMPEC_MARCH_MIN_CELLS"] = "0"
    try:
        out = []
        for sgs in (1, 0):
            ctx = I.Context(Pi, emulated=emulated)
            ctx.set_option("sgs_march", sgs)
            ctx.upload_c(c); ctx.upload_q(q)
            dt, _ = ctx.impes_step(0.0, 0.0)          # velocities for the dispersion tensor
            ctx.upload_c(c); ctx.fill_boundaries()
            info = ctx.diffusion_split(0.5 * dt)
            assert info.converged == 1
            out.append((info.iterations, ctx.download_c()))
            ctx.close()
    finally:
        os.environ.pop("IMPEC_MARCH_MIN_CELLS", None)
    assert out[0][0] == out[1][0], (out[0][0], out[1][0])
    assert np.abs(out[0][1] - out[1][1]).max() <= 1e-13, np.abs(out[0][1] - out[1][1]).max()


def check_fast_initial_residual(emulated, N1=48, N2=80, stype=O.DIAGONAL, eps=1e-10):
    """x0 = 0 (the reference's start): the initial PCG residual as a copy of the right-hand side plus a 3 x 3 fix-up around the
    pinned cell must give the very same solve as the full residual pass -- bit for bit."""
    base, Po, c, q = production_case(N1, N2, stype=stype, eps=eps)
    Pi = I.default_params(emulated=emulated, **base)
    res = []
    for fast in (1, 0):
        ctx = I.Context(Pi, emulated=emulated)
        ctx.set_option("fast_r0", fast)
        ctx.upload_c(c); ctx.upload_q(q); ctx.fill_boundaries()
        info = ctx.pressure_solve()
        res.append((info.iterations, info.relres, ctx.download_p()))
        ctx.close()
    assert res[0][0] == res[1][0] and res[0][1] == res[1][1]
    assert np.array_equal(res[0][2], res[1][2])


def check_small_vs_tiles(emulated, N1=256, N2=384, stype=O.DIAGONAL):
    """The single-CTA bottom-of-hierarchy kernel (kernels_small.inl, levels of <= 4096 cells) against the tile kernels
    on the same levels: same arithmetic per cell, so one cycle agrees to rounding (<= 1e-13 relative), with FP64 and
    with FP32 coefficient copies."""
    base, Po, c, q = production_case(N1, N2, stype=stype)
    Pi = I.default_params(emulated=emulated, **base)
    r = np.random.default_rng(2).standard_normal(N1 * N2)
    for f32 in (0, 1):
        out = []
        for small in (0, 1):
            ctx = I.Context(Pi, emulated=emulated)
            ctx.set_option("mg_small", small); ctx.set_option("mg_links_fp32", f32)
            ctx.upload_c(c); ctx.upload_q(q); ctx.fill_boundaries(); ctx.build_operator()
            n0 = ctx.kernel_launches()
            z1 = ctx.vcycle(r); n1 = ctx.kernel_launches(); z2 = ctx.vcycle(r)
            assert np.array_equal(z1, z2), "cycle is not repeatable (race?)"
            out.append((z1, n1 - n0))
            ctx.close()
        assert out[1][1] < out[0][1], "the small-level kernel was not used"
        assert np.abs(out[0][0] - out[1][0]).max() <= 1e-13 * np.abs(out[0][0]).max()


def check_dispersion(emulated, N1=32, N2=32, stype=O.PARALLEL, steps=1, alpha_l=1.0, alpha_t=1.0):
    """a19: Crank-Nicolson dispersion half step on its own and inside a full IMPES step of a *_dispersion.cfg
    operator (two CN solves per step).  Tolerance: max|dc| <= 1e-9 (both sides iterate to para_eps = 1e-14)."""
    lib = O.lib()
    # alpha_l != alpha_t: the full, non-symmetric 9-point Crank-Nicolson operator (src/MGSolver.cpp:923-948), BiCGStab on both sides
    base, Po, c, q = production_case(N1, N2, stype=stype, diffusion=1, eps=1e-12, alpha_l=alpha_l, alpha_t=alpha_t)
    Pi = I.default_params(emulated=emulated, **base)
    ctx = I.Context(Pi, emulated=emulated)
    ctx.upload_c(c); ctx.upload_q(q)
    So = O.new_state()
    p = O.scalar_field(Po); u = O.velocity_field(Po); v = O.velocity_field(Po)
    co = c.copy()
    # velocities first (pressure solve), then one stand-alone CN half step on both sides
    lib.orc_fill_boundaries(C.byref(Po), co)
    lib.orc_solve_elliptic(C.byref(Po), C.byref(So), co, q, p)
    lib.orc_calc_velocities(C.byref(Po), p, co, u, v)
    dt = lib.orc_max_stable_dt(C.byref(Po), u, v)
    ctx.fill_boundaries(); ctx.pressure_solve(); ctx.calc_velocities()
    assert abs(ctx.max_stable_dt() - dt) <= 1e-9 * dt
    lib.orc_solve_parabolic(C.byref(Po), C.byref(So), co, u, v, 0.5 * dt)
    info = ctx.diffusion_split(0.5 * dt)
    assert info.converged == 1 and abs(info.iterations - So.last_parabolic_its) <= 2
    cg = ctx.download_c()
    assert np.abs(interior(Po, cg) - interior(Po, co)).max() <= 1e-9
    # mass is conserved by the CN step with closed walls
    assert abs(interior(Po, cg).sum() - interior(Po, c).sum()) <= 1e-10 * interior(Po, c).sum()
    # full steps
    ctx.upload_c(c); co = c.copy(); t = 0.0
    for _ in range(steps):
        dto = lib.orc_impes_step(C.byref(Po), C.byref(So), p, co, u, v, q, t, 0.0)
        dtg, _ = ctx.impes_step(t, 0.0)
        assert abs(dtg - dto) <= 1e-9 * dto
        t += dto
    cg = ctx.download_c()
    assert np.abs(cg - co).max() <= 1e-8
    ctx.close()


def check_generators(emulated, N=96):
    """impec_synth_state / impec_init_source (the product's input generators used by bench.py) against the oracle's, which are pinned
    to the reference's well model (tests/test_oracle_vs_ref.py): bit for bit, circle and one-cell wells, both grid orientations."""
    lib = O.lib()
    for stype in (O.PARALLEL, O.DIAGONAL):
        L = 1.41421356237 if stype == O.PARALLEL else 1.0
        kw = dict(N1=N, N2=N + 16, x_ub=L, y_ub=L, stype=stype)
        if stype == O.PARALLEL:
            kw["N2"] = N
        Po = O.default_params(**kw); Pi = I.default_params(emulated=emulated, **kw)
        co = O.scalar_field(Po); lib.orc_synth_state(C.byref(Po), 20260917, co)
        assert np.array_equal(I.synth_state(Pi, emulated=emulated), co)
        for profile in (0, 2):
            qo = np.zeros(Po.N1 * Po.N2); lib.orc_init_source(C.byref(Po), 1.0, 0.05, profile, qo)
            assert np.array_equal(I.init_source(Pi, 1.0, 0.05, profile, emulated=emulated), qo) and np.abs(qo).max() > 0
        # slab-wise generation (multi-GPU bench) equals the rows of the global arrays
        r0, nr = Po.N1 // 4, Po.N1 // 2
        cs = I.synth_state_rows(Pi, r0, nr, emulated=emulated)
        assert np.array_equal(cs[Po.g:-Po.g], co[Po.g + r0:Po.g + r0 + nr])
        qs = I.init_source_rows(Pi, r0, nr, 1.0, 0.05, 0, emulated=emulated)
        qo = np.zeros(Po.N1 * Po.N2); lib.orc_init_source(C.byref(Po), 1.0, 0.05, 0, qo)
        assert np.array_equal(qs, qo.reshape(Po.N1, Po.N2)[r0:r0 + nr].ravel())


def check_one_cell_wells(emulated, N=64):
    """well_profile = one_cell (src/Util.cpp:565-586): two IMPES steps against the oracle, parallel and diagonal layout."""
    lib = O.lib()
    for stype in (O.PARALLEL, O.DIAGONAL):
        base, Po, c, _q = production_case(N, N, stype=stype, eps=1e-12)
        Pi = I.default_params(emulated=emulated, **base)
        q = np.zeros(N * N); lib.orc_init_source(C.byref(Po), 1.0, 0.05, 2, q)
        assert np.count_nonzero(q) in (2, 4)
        ctx = I.Context(Pi, emulated=emulated)
        ctx.upload_c(c); ctx.upload_q(q)
        So = O.new_state(); p = O.scalar_field(Po); u = O.velocity_field(Po); v = O.velocity_field(Po); co = c.copy()
        t = 0.0
        for _ in range(2):
            dto = lib.orc_impes_step(C.byref(Po), C.byref(So), p, co, u, v, q, t, 0.0)
            dtg, info = ctx.impes_step(t, 0.0)
            assert info.converged == 1 and abs(dtg - dto) <= 1e-9 * dto
            t += dto
        assert rel_l2(interior(Po, ctx.download_p()), interior(Po, p)) <= TOL_P
        assert np.abs(ctx.download_c() - co).max() <= 1e-8
        ctx.close()


def check_explicit_diffusive_dt(emulated, N=72):
    """para_method = explicit with dispersion: dt = min(dt_A, lambda min dx^2 / (2 (d11 + d22))) (src/HyperbolicSolver.cpp:157-196).  The
    device reduction evaluates the same expression per cell in the same order: BIT-EXACT."""
    lib = O.lib()
    for al, at, lim in ((1.0, 1.0, 0.9), (0.4, 0.4, 0.6)):
        base, Po, c, q = production_case(N, N + 8, stype=O.DIAGONAL, diffusion=1, ptype=O.EXPLICIT, alpha_l=al, alpha_t=at, dt_diff_limit=lim)
        Pi = I.default_params(emulated=emulated, **base)
        rng = np.random.default_rng(31)
        u = O.velocity_field(Po); v = O.velocity_field(Po)
        u[...] = 6.0 * (rng.random(u.shape) - 0.5); v[...] = 6.0 * (rng.random(v.shape) - 0.5)
        ctx = I.Context(Pi, emulated=emulated)
        ctx.upload_velocity(aos(u, v))
        dto = lib.orc_max_stable_dt(C.byref(Po), u, v)
        dta = lib.orc_max_stable_advection_dt(C.byref(Po), u, v)
        assert ctx.max_stable_dt() == dto, (ctx.max_stable_dt(), dto)
        assert dto < dta                                          # the diffusive limit is the binding one here
        ctx.close()


def check_diagnostic_values(emulated, N=80):
    """CalculateDivergenceError, CalculateVelocityBoundaryError, GetTotalMass (src/Impec2DSolver.cpp:542-620) as device reductions
    against the oracle's restatement on the velocity field of a converged step; also with a perturbed (non-solenoidal, leaking) field."""
    lib = O.lib()
    base, Po, c, q = production_case(N, N, eps=1e-12)
    Pi = I.default_params(emulated=emulated, **base)
    ctx = I.Context(Pi, emulated=emulated)
    ctx.upload_c(c); ctx.upload_q(q)
    ctx.fill_boundaries(); ctx.pressure_solve(); ctx.calc_velocities()
    vel = ctx.download_velocity()
    for perturb in (False, True):
        if perturb:
            vel = vel + 0.05 * np.random.default_rng(8).standard_normal(vel.shape)
            ctx.upload_velocity(np.ascontiguousarray(vel))
        u = np.ascontiguousarray(vel[..., 0]); v = np.ascontiguousarray(vel[..., 1])
        d = ctx.diagnostics()
        dv = lib.orc_divergence_error(C.byref(Po), u, v, q)
        bv = lib.orc_boundary_velocity_error(C.byref(Po), u, v)
        assert abs(d.divergence_error - dv) <= 1e-9 * max(dv, 1e-12) + 1e-13, (d.divergence_error, dv)
        assert abs(d.boundary_velocity_error - bv) <= 1e-10 * max(bv, 1e-12) + 1e-13, (d.boundary_velocity_error, bv)
    cg = ctx.download_c()
    assert abs(d.total_mass - lib.orc_total_mass(C.byref(Po), cg)) <= 1e-11 * d.total_mass
    assert d.c_max == interior(Po, cg).max() and d.c_min == interior(Po, cg).min()
    ctx.close()


def check_impes_step_production(emulated, N, steps=1, tol_p=TOL_P, tol_c=1e-8):
    """One IMPES step of the parallel.cfg operator with the cfg's own tolerance (eps = 1e-14) and the DEFAULT kernel selection (no
    option or environment overrides) at a size where the production kernels run: >= 2896^2 puts the fine level on the register-march
    kernels, everything below on the tile / small-level kernels.  Oracle: the same step on the CPU (seconds)."""
    lib = O.lib()
    base, Po, c, q = production_case(N, N, eps=1e-14)
    Pi = I.default_params(emulated=emulated, **base)
    ctx = I.Context(Pi, emulated=emulated)
    ctx.upload_c(c); ctx.upload_q(q)
    So = O.new_state(); p = O.scalar_field(Po); u = O.velocity_field(Po); v = O.velocity_field(Po); co = c.copy()
    t = 0.0
    for _ in range(steps):
        dto = lib.orc_impes_step(C.byref(Po), C.byref(So), p, co, u, v, q, t, 0.0)
        dtg, info = ctx.impes_step(t, 0.0)
        assert info.converged == 1 and abs(info.iterations - So.last_its) <= 1, (info.iterations, So.last_its)
        assert abs(dtg - dto) <= 1e-9 * dto
        t += dto
    ep = rel_l2(interior(Po, ctx.download_p()), interior(Po, p)); ec = np.abs(ctx.download_c() - co).max()
    print("impes step %d^2: its gpu/oracle %d/%d  true_relres gpu/oracle %.2e/%.2e  |dp|/|p| %.3e  max|dc| %.3e" % (N, info.iterations, So.last_its, info.true_relres, So.last_true_relres, ep, ec))
    assert ep <= tol_p and ec <= tol_c
    ctx.close()
    return ep, ec


def check_dispersion_production(emulated, N1, N2):
    """diagonal_dispersion.cfg step (two CN half steps) against the oracle at a size where the marching CN kernels run."""
    lib = O.lib()
    base, Po, c, q = production_case(N1, N2, stype=O.DIAGONAL, diffusion=1, eps=1e-14)
    Pi = I.default_params(emulated=emulated, **base)
    ctx = I.Context(Pi, emulated=emulated)
    ctx.upload_c(c); ctx.upload_q(q)
    So = O.new_state(); p = O.scalar_field(Po); u = O.velocity_field(Po); v = O.velocity_field(Po); co = c.copy()
    dto = lib.orc_impes_step(C.byref(Po), C.byref(So), p, co, u, v, q, 0.0, 0.0)
    dtg, info = ctx.impes_step(0.0, 0.0)
    assert abs(dtg - dto) <= 1e-9 * dto and abs(info.iterations - So.last_its) <= 1
    ep = rel_l2(interior(Po, ctx.download_p()), interior(Po, p)); ec = np.abs(ctx.download_c() - co).max()
    print("dispersion step %dx%d: |dp|/|p| %.3e max|dc| %.3e" % (N1, N2, ep, ec))
    assert ep <= TOL_P and ec <= 1e-8
    ctx.close()


def check_hyperbolic_test_config(emulated, N, steps=3):
    """BASELINE config 2: hyperbolic_test.cfg transport steps (sine cell averages, velocity (-1,+1), periodic, CFL 0.95) at the
    config's size: BIT-EXACT against the oracle."""
    lib = O.lib()
    base = dict(N1=N, N2=N, x_ub=1.0, y_ub=1.0, is_test=1, testno=1, stype=O.DIAGONAL, cfl=0.95, diffusion=0)
    Po = O.default_params(**base); Pi = I.default_params(emulated=emulated, **base)
    dx = 1.0 / N
    x = (np.arange(N) + 0.5) * dx
    f = np.cos(2.0 * np.pi * (x + dx / 2.0)) - np.cos(2.0 * np.pi * (x - dx / 2.0))
    c = O.scalar_field(Po); interior(Po, c)[...] = 1.0 + 1.0 / (dx * dx * 4.0 * np.pi * np.pi) * np.outer(f, f)
    u = O.velocity_field(Po, -1.0); v = O.velocity_field(Po, 1.0); q = np.zeros(N * N)
    ctx = I.Context(Pi, emulated=emulated)
    ctx.upload_c(c); ctx.upload_q(q); ctx.upload_velocity(aos(u, v))
    dt = ctx.max_stable_dt()
    assert dt == lib.orc_max_stable_dt(C.byref(Po), u, v) == 0.95 * Po.dx
    So = O.new_state(); co = c.copy()
    for _ in range(steps):
        lib.orc_transport_step(C.byref(Po), C.byref(So), co, u, v, q, dt)
        ctx.transport_step(dt)
    assert np.array_equal(ctx.download_c(), co)
    ctx.close()


def check_integral_diagnostics(emulated, N=96, steps=3):
    """impec_integral_diagnostics (device reductions) against the host definitions of impec2d_b200/diagnostics.py evaluated on the
    downloaded field: the mixing-zone cell count is exact, the weighted sums agree to rounding."""
    from impec2d_b200 import diagnostics as D
    base, Po, c, q = production_case(N, N, eps=1e-10)
    Pi = I.default_params(emulated=emulated, **base)
    ctx = I.Context(Pi, emulated=emulated)
    ctx.upload_c(c); ctx.upload_q(q)
    L = Po.x_ub - Po.x_lb
    rd_host = D.RunDiagnostics(q, Po.dx, Po.dy, L, L); rd_dev = D.RunDiagnostics(q, Po.dx, Po.dy, L, L)
    t = 0.0
    for _ in range(steps):
        dt, _i = ctx.impes_step(t, 0.0); t += dt
        ci = interior(Po, ctx.download_c())
        d = ctx.integral_diagnostics()
        assert d.mixing_zone_cells == float(((ci >= 0.1) & (ci <= 0.9)).sum())
        assert abs(d.mixing_zone_length - D.mixing_zone_length(ci, Po.dx, Po.dy, L, L)) <= 1e-14
        assert abs(d.mass_in_place - D.mass_in_place(ci, Po.dx, Po.dy)) <= 1e-12 * d.mass_in_place
        pf = D.produced_fraction(ci, q)
        assert pf > 0 and abs(d.produced_fraction - pf) <= 1e-12 * pf
        rd_host.record(t, ci); rd_dev.record_device(t, ctx)
    sh, sd = rd_host.summary(), rd_dev.summary()
    for k in sh:
        assert sh[k] == sd[k] or abs(sh[k] - sd[k]) <= 1e-12 * abs(sh[k]), k
    ctx.close()


def check_fused_build(emulated, N1=96, N2=141, big=False):
    """The one-march operator build (kernels_build.inl: c -> links, inverse diagonal, FP32 copies) and the one-march velocity + CFL
    kernel (c, p -> u, v, max speed, diffusive limit) against the first-generation passes (k_mobility, k_vertex, k_links, k_diag_inv,
    k_links_to_f32; k_velocity, k_dt_reduce): every array must be BIT-IDENTICAL, observed through one multigrid cycle on a random
    right-hand side (reads the FP32 or FP64 coefficient sets of every level), a full pressure solve (FP64 links in the PCG mat-vec),
    the velocity field and the time step -- 9-point and 5-point operators, both averaging procedures, odd and even column counts,
    dx != dy, with and without dispersion."""
    cases = [(N1, N2, O.DIAGONAL, 2.0 / 3.0, O.HARMONIC, 1), (N1, N2 + 1, O.DIAGONAL, 1.0, O.MEANCONC, 0), (N1, N1, O.PARALLEL, 0.5, O.MEANCONC, 0)]
    if big:
        cases = [(N1, N2, O.DIAGONAL, 2.0 / 3.0, O.HARMONIC, 1)]
    for (n1, n2, stype, nu, avg, diff) in cases:
        base, Po, c, q = production_case(n1, n2, stype=stype, nu=nu, avgproc=avg, eps=1e-9, diffusion=diff, ptype=O.EXPLICIT)
        Pi = I.default_params(emulated=emulated, **base)
        r = np.random.default_rng(12).standard_normal(n1 * n2)
        for f32 in ((1, 0) if not big else (1,)):
            out = []
            for fused in (1, 0):
                ctx = I.Context(Pi, emulated=emulated)
                ctx.set_option("build_fused", fused); ctx.set_option("vel_march", fused); ctx.set_option("mg_links_fp32", f32)
                ctx.upload_c(c); ctx.upload_q(q); ctx.fill_boundaries(); ctx.build_operator()
                z = ctx.vcycle(r)
                info = ctx.pressure_solve()
                ctx.calc_velocities()
                out.append((z, info.iterations, ctx.download_p(), ctx.download_velocity(), ctx.max_stable_dt()))
                ctx.close()
            assert np.array_equal(out[0][0], out[1][0]), "cycle differs (fp32 %d): %g" % (f32, np.abs(out[0][0] - out[1][0]).max())
            assert out[0][1] == out[1][1] and np.array_equal(out[0][2], out[1][2]), "pressure differs (fp32 %d)" % f32
            assert np.array_equal(out[0][3], out[1][3]), "velocity differs: %g" % np.abs(out[0][3] - out[1][3]).max()
            assert out[0][4] == out[1][4], ("time step differs", out[0][4], out[1][4])
