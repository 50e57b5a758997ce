# side measurement: chunk height of the marching multigrid kernels (wave quantisation: 3 CTAs/SM x 148 SMs = 444 resident CTAs)
import sys, json
sys.path.insert(0, '.')
import impec2d_b200 as I
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
P = I.default_params(N1=N, N2=N, x_ub=1.41421356237, y_ub=1.41421356237, stype=0, nu=2/3, eps=1e-14)
c = I.synth_state(P); q = I.init_source(P)
ctx = I.Context(P); ctx.upload_c(c); ctx.upload_q(q)
dt, info = ctx.impes_step(0.0)
print("N", N, "its", info.iterations, flush=True)
out = {}
lv = 0
while True:
    n = N >> lv
    if n * n < 8.0e6: break
    xs_pre = -(-(-(-n // 56)) // 4); xs_post = -(-(-(-n // 60)) // 4)
    for rows in [int(a) for a in sys.argv[2].split(",")]:
        try:
            ctx.set_option("mg_march_rows@%d" % lv, rows)
        except Exception as e:
            print("level", lv, "rows", rows, "rejected", e); continue
        ch = -(-n // rows)
        t = [ctx.profile_kernel("%s@%d" % (k, lv), 20) for k in ("fused_pre", "fused_post", "fused_plain")]
        out["%d:%d" % (lv, rows)] = t
        print("level %d (%d^2) rows %3d chunks %3d  pre %.4f ms (%.2f waves)  post %.4f ms (%.2f waves)  plain %.4f ms  sum %.4f" % (lv, n, rows, ch, t[0], xs_pre * ch / 444., t[1], xs_post * ch / 444., t[2], sum(t)), flush=True)
    lv += 1
json.dump(out, open('gpurun_out/r2y_march_rows_%d.json' % N, 'w'), indent=1)
