# first-look timings on the B200 (not the bench): per-phase wall times with stream syncs
import sys, time, ctypes as C
sys.path.insert(0, '.')
import numpy as np
import impec2d_b200 as I
import oracle as O
lib = O.lib()
for N in (2048, 8192, 16384):
    L = 1.41421356237
    kw = dict(N1=N, N2=N, x_ub=L, y_ub=L, stype=0, nu=2/3, eps=1e-8)
    Po = O.default_params(**kw); Pi = I.default_params(**kw)
    t0 = time.time()
    c = O.scalar_field(Po); lib.orc_synth_state(C.byref(Po), 20260917, c)
    q = np.zeros(N*N); lib.orc_init_source(C.byref(Po), 1.0, 0.05, 0, q)
    print('N', N, 'host setup %.1fs' % (time.time()-t0), flush=True)
    ctx = I.Context(Pi)
    t0 = time.time(); ctx.upload_c(c); ctx.upload_q(q); ctx.synchronize(); print(' upload %.3fs' % (time.time()-t0))
    def T(f, *a):
        ctx.synchronize(); t = time.perf_counter(); r = f(*a); ctx.synchronize(); return r, (time.perf_counter()-t)*1e3
    _, t = T(ctx.fill_boundaries); print(' fill %.3f ms' % t)
    _, t = T(ctx.build_operator); print(' build_operator %.3f ms' % t)
    _, t = T(ctx.build_operator); print(' build_operator(2) %.3f ms' % t)
    for k in range(3):
        _, t = T(ctx.vcycle); print(' vcycle %.3f ms (roofline 208B/cell: %.3f)' % (t, N*N*208/6.45e12*1e3))
    for eps in (1e-8, 1e-14):
        info, t = T(ctx.pressure_solve, eps); print(' solve eps=%g: its %d relres %.2e true %.2e  %.1f ms  (%.2f ms/it)' % (eps, info.iterations, info.relres, info.true_relres, t, t/max(info.iterations,1)))
    _, t = T(ctx.calc_velocities); print(' velocities %.3f ms' % t)
    dt, t = T(ctx.max_stable_dt); print(' max_stable_dt %.3f ms dt=%g' % (t, dt))
    for k in range(2):
        _, t = T(ctx.transport_step, dt); print(' transport_step %.3f ms (4-pass roofline 128B/cell: %.3f)' % (t, N*N*128/6.45e12*1e3))
    (dtu, info), t = T(ctx.impes_step); print(' impes_step (eps=%g) %.1f ms its %d -> %.3f Gcell/s' % (Pi.eps, t, info.iterations, N*N/t/1e6))
    d = ctx.diagnostics(); print(' diag mass %.6g div %.3e bnd %.3e cmin %.3g cmax %.6g launches %d' % (d.total_mass, d.divergence_error, d.boundary_velocity_error, d.c_min, d.c_max, ctx.kernel_launches()))
    ctx.close()
