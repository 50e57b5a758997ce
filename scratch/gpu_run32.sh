timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "persistent_tail or small_vs or fused_vs" 2>&1 | tail -3
timeout 300 python - <<PY
import sys, os; sys.path.insert(0,'.')
import impec2d_b200 as I
N=16384
P = I.default_params(N1=N, N2=N, x_ub=1.41421356237, y_ub=1.41421356237, stype=0, nu=2/3, eps=1e-14)
c = I.synth_state(P); q = I.init_source(P)
ctx = I.Context(P); ctx.upload_c(c); ctx.upload_q(q)
import time
for persist in (0, 1, 0, 1):
    ctx.set_option("mg_persist_tail", persist)
    ctx.upload_c(c); dt, info = ctx.impes_step(0.0)
    ctx.upload_c(c); ctx.synchronize(); t0 = time.perf_counter(); dt, info = ctx.impes_step(0.0); ctx.synchronize(); el = time.perf_counter() - t0
    print("persist", persist, "its", info.iterations, "relres %.3e" % info.relres, "step %.1f ms" % (el * 1e3), " ".join("%s %.4f" % (k, ctx.profile_kernel(k, 20)) for k in ("tail", "cycle@5", "cycle@3", "cycle@2", "vcycle")))
ctx.close()
PY
