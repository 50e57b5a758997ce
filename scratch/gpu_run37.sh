timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "sgs_march or dispersion" 2>&1 | tail -2
timeout 400 python - <<PY
import sys, os, time; sys.path.insert(0,'.')
import impec2d_b200 as I
N=16384
P = I.default_params(N1=N, N2=N, x_ub=1.0, y_ub=1.0, stype=1, nu=2/3, eps=1e-14, diffusion=1)
c = I.synth_state(P); q = I.init_source(P)
ctx = I.Context(P); ctx.upload_c(c); ctx.upload_q(q)
dt, info = ctx.impes_step(0.0, 0.0)
for sgs in (0, 1, 0, 1):
    ctx.set_option("sgs_march", sgs)
    ctx.upload_c(c); ctx.fill_boundaries(); ctx.synchronize(); t0 = time.perf_counter(); info = ctx.diffusion_split(0.5 * dt); ctx.synchronize()
    print("sgs_march %d: CN half step %.1f ms, its %d relres %.2e" % (sgs, (time.perf_counter() - t0) * 1e3, info.iterations, info.relres), flush=True)
ctx.set_option("sgs_march", 1)
t = 0.0
for s in range(2):
    ctx.synchronize(); t0 = time.perf_counter(); dt, info = ctx.impes_step(t, 0.0); ctx.synchronize(); t += dt
    print("full step with dispersion: %.1f ms (pressure its %d)" % ((time.perf_counter() - t0) * 1e3, info.iterations))
ctx.close()
PY
