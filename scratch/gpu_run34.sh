timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "march_vs_tile" 2>&1 | tail -2
timeout 300 python - <<PY
import sys, os; sys.path.insert(0,'.')
import impec2d_b200 as I
N=16384
P = I.default_params(N1=N, N2=N, x_ub=1.41421356237, y_ub=1.41421356237, stype=0, nu=2/3, eps=1e-14)
c = I.synth_state(P); q = I.init_source(P)
ctx = I.Context(P); ctx.upload_c(c); ctx.upload_q(q); dt, info = ctx.impes_step(0.0)
for v in (1, 0, 1):
    ctx.set_option("mg_march_variant", v)
    print("variant", v, "its", info.iterations, " ".join("%s %.4f" % (k, ctx.profile_kernel(k, 20)) for k in ("fused_pre", "fused_pre@1", "fused_pre@2", "fused_post", "vcycle")), flush=True)
ctx.close()
PY
