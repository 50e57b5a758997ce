timeout 400 python - <<PY
import sys, os, time; sys.path.insert(0,'.')
import impec2d_b200 as I
N=16384
# diagonal_dispersion.cfg operator (BASELINE config 4): unit square, diagonal wells, alpha_l = alpha_t = 1, Crank-Nicolson
P = I.default_params(N1=N, N2=N, x_ub=1.0, y_ub=1.0, stype=1, nu=2/3, eps=1e-14, diffusion=1)
c = I.synth_state(P); q = I.init_source(P)
ctx = I.Context(P); ctx.upload_c(c); ctx.upload_q(q)
t = 0.0
for s in range(3):
    ctx.synchronize(); t0 = time.perf_counter()
    dt, info = ctx.impes_step(t, 0.0); ctx.synchronize(); el = time.perf_counter() - t0
    t += dt
    d = ctx.diagnostics()
    print("step %d: %.1f ms  pressure its %d relres %.2e  dt %.3e  mass %.6e" % (s, el * 1e3, info.iterations, info.relres, dt, d.total_mass), flush=True)
ctx.synchronize(); t0 = time.perf_counter(); info = ctx.diffusion_split(0.5 * dt); ctx.synchronize()
print("one CN half step: %.1f ms, its %d relres %.2e" % ((time.perf_counter() - t0) * 1e3, info.iterations, info.relres))
ctx.close()
PY
