# side measurement: PCG iterations and step time for different multigrid cycle shapes at the bench workload
import sys, time, json
sys.path.insert(0, '.')
import impec2d_b200 as I
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
P = I.default_params(N1=N, N2=N, x_ub=1.41421356237, y_ub=1.41421356237, stype=0, nu=2/3, eps=1e-14)
c = I.synth_state(P); q = I.init_source(P)
out = {}
ctx = I.Context(P); ctx.upload_c(c); ctx.upload_q(q)
for (wf, wt) in ((2, 4), (2, 3), (3, 4), (3, 3), (2, 2), (4, 4), (1, 3), (5, 4)):
    ctx.set_option("mg_w_from", wf); ctx.set_option("mg_w_to", wt)
    ctx.reset_state(); ctx.upload_c(c)
    t = 0.0; its = []
    dt, info = ctx.impes_step(t); t += dt
    ctx.synchronize(); t0 = time.perf_counter()
    for s in range(2):
        dt, info = ctx.impes_step(t); t += dt; its.append(info.iterations)
    ctx.synchronize(); ms = (time.perf_counter() - t0) / 2 * 1e3
    vc = ctx.profile_kernel("vcycle", 5); tl = ctx.profile_kernel("tail", 20)
    out["w%d-%d" % (wf, wt)] = {"ms_per_step": ms, "iterations": its, "vcycle_ms": vc, "tail_visit_ms": tl, "true_relres": info.true_relres}
    print("W levels %d..%d: %.1f ms/step its %s vcycle %.2f ms tail visit %.3f ms true %.2e" % (wf, wt, ms, its, vc, tl, info.true_relres), flush=True)
json.dump(out, open('gpurun_out/r2x_cycle_shapes.json', 'w'), indent=1)
