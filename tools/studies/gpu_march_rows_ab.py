# side measurement: chunk heights of the marching kernels per level, whole V-cycle and whole step, configurations interleaved (noise control)
import sys, time, json
sys.path.insert(0, '.')
import impec2d_b200 as I
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
P = I.default_params(N1=N, N2=N, x_ub=1.41421356237, y_ub=1.41421356237, stype=0, nu=2/3, eps=1e-14)
c = I.synth_state(P); q = I.init_source(P)
ctx = I.Context(P); ctx.upload_c(c); ctx.upload_q(q)
t = 0.0
dt, info = ctx.impes_step(t); t += dt
configs = [tuple(int(a) for a in s.split(",")) for s in sys.argv[2].split(":")]
res = {cf: {"vcycle": [], "step": []} for cf in configs}
for rnd in range(3):
    for cf in configs:
        for l, r in enumerate(cf): ctx.set_option("mg_march_rows@%d" % l, r)
        v = ctx.profile_kernel("vcycle", 30)
        ctx.synchronize(); t0 = time.perf_counter()
        for s in range(2):
            dt, info = ctx.impes_step(t); t += dt
        ctx.synchronize(); ms = (time.perf_counter() - t0) / 2 * 1e3
        res[cf]["vcycle"].append(v); res[cf]["step"].append(ms)
        print("round %d rows %-14s vcycle %.3f ms  step %.1f ms  its %d" % (rnd, cf, v, ms, info.iterations), flush=True)
print()
for cf in configs:
    v = res[cf]["vcycle"]; s = res[cf]["step"]
    print("rows %-14s vcycle min %.3f mean %.3f   step min %.1f mean %.1f" % (cf, min(v), sum(v) / len(v), min(s), sum(s) / len(s)))
json.dump({",".join(map(str, k)): v for k, v in res.items()}, open('gpurun_out/r2z_march_rows_ab_%d.json' % N, 'w'), indent=1)
