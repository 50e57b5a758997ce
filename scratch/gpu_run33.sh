timeout 500 python - <<PY
import sys, os; sys.path.insert(0,'.')
import impec2d_b200 as I
N=16384
P = I.default_params(N1=N, N2=N, x_ub=1.41421356237, y_ub=1.41421356237, stype=0, nu=2/3, eps=1e-14)
c = I.synth_state(P); q = I.init_source(P)
for name, env in (("default", {}), ("L3 TH64", {"IMPEC_TH32_BELOW": "3e6"}), ("L4 TH32", {"IMPEC_TH16_BELOW": "5e5"}), ("L3 TH16", {"IMPEC_TH16_BELOW": "6e6", "IMPEC_TH32_BELOW": "6e6"}),
                  ("march>=4e6", {"IMPEC_MARCH_MIN_CELLS": "4e6"}), ("march>=1e6", {"IMPEC_MARCH_MIN_CELLS": "1e6"})):
    for k in ("IMPEC_TH32_BELOW", "IMPEC_TH16_BELOW", "IMPEC_MARCH_MIN_CELLS"): os.environ.pop(k, None)
    os.environ.update(env)
    ctx = I.Context(P); ctx.upload_c(c); ctx.upload_q(q); dt, info = ctx.impes_step(0.0)
    print("%-12s its %d " % (name, info.iterations) + " ".join("%s %.4f" % (k, ctx.profile_kernel(k, 20)) for k in ("tail", "cycle@3", "cycle@2", "cycle@1", "vcycle")), flush=True)
    ctx.close()
PY
