# side measurement: warm per-level cost of the multigrid tile kernels and of the sub-cycles
import sys, json
sys.path.insert(0, '.')
import impec2d_b200 as I
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
P = I.default_params(N1=N, N2=N, x_ub=1.41421356237, y_ub=1.41421356237, stype=0, nu=2/3, eps=1e-14)
c = I.synth_state(P); q = I.init_source(P)
ctx = I.Context(P); ctx.upload_c(c); ctx.upload_q(q); ctx.impes_step(0.0)
out = {}
nl = ctx.mg_levels()
for l in range(0, nl - 1):
    n = N >> l
    row = {}
    for k in ("fused_pre", "fused_post", "fused_plain", "cycle"):
        row[k] = ctx.profile_kernel("%s@%d" % (k, l), 20 if l > 1 else 5) * 1e3
    out[l] = row
    print("level %2d (%5d^2): pre %9.1f us  post %9.1f us  plain %9.1f us  cycle-from-here %9.1f us" % (l, n, row["fused_pre"], row["fused_post"], row["fused_plain"], row["cycle"]), flush=True)
json.dump(out, open('gpurun_out/levels.json', 'w'), indent=1)
