# helper for ncu captures: launches each hot kernel on the fine level of the bench workload (16384^2 by default)
import sys, ctypes as C
sys.path.insert(0, '.')
import numpy as np
import impec2d_b200 as I
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
names = sys.argv[2].split(',') if len(sys.argv) > 2 else ["fused_pre", "fused_post", "fused_plain"]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
P = I.default_params(N1=N, N2=N, x_ub=1.41421356237, y_ub=1.41421356237, stype=0, nu=2/3, eps=1e-14)
c = I.synth_state(P); q = I.init_source(P)
ctx = I.Context(P)
ctx.upload_c(c); ctx.upload_q(q); ctx.fill_boundaries(); ctx.build_operator()
r = np.random.default_rng(0).standard_normal(N * N)
for n in names:
    print(n, ctx.profile_kernel(n, reps))
