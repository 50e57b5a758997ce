mkdir -p gpurun_out
python - <<PY
import sys, os; sys.path.insert(0,'.')
import impec2d_b200 as I
N=16384
P = I.default_params(N1=N, N2=N, x_ub=1.41421356237, y_ub=1.41421356237, stype=0, nu=2/3, eps=1e-14)
c = I.synth_state(P); q = I.init_source(P)
ctx = I.Context(P); ctx.upload_c(c); ctx.upload_q(q); dt, info = ctx.impes_step(0.0)
for v in (0, 1, 2, 3, 0):
    ctx.set_option("mg_march_variant", v)
    print("  variant", v, " ".join("%s %.3f" % (k, ctx.profile_kernel(k, 10)) for k in ("fused_pre", "fused_post", "fused_plain", "fused_pre@1", "fused_post@1", "cycle@1", "vcycle")))
ctx.close()
PY
