# side measurements (not the bench line): effect of the optional solver switches on the 16384^2 IMPES step
import sys, time, json
sys.path.insert(0, '.')
import numpy as np
import impec2d_b200 as I
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
P = I.default_params(N1=N, N2=N, x_ub=1.41421356237, y_ub=1.41421356237, stype=0, nu=2/3, eps=1e-14)
c = I.synth_state(P); q = I.init_source(P)
out = {}
for name, opts in (("default", {}), ("stagnation_guard", {"stagnation_guard": 1}), ("warm_start", {"warm_start": 1}),
                   ("warm_start+guard", {"warm_start": 1, "stagnation_guard": 1}), ("fp64_links", {"mg_links_fp32": 0}), ("staged", {"mg_staged": 1}), ("no_graph", {"mg_graph": 0})):
    ctx = I.Context(P)
    for k, v in opts.items(): ctx.set_option(k, v)
    ctx.upload_c(c); ctx.upload_q(q)
    t = 0.0; its = []; rr = []
    for s in range(2): dt, info = ctx.impes_step(t); t += dt
    ctx.synchronize(); t0 = time.perf_counter()
    for s in range(3):
        dt, info = ctx.impes_step(t); t += dt; its.append(info.iterations); rr.append(info.true_relres)
    ctx.synchronize(); ms = (time.perf_counter() - t0) / 3 * 1e3
    out[name] = {"ms_per_step": ms, "iterations": its, "true_relres": rr, "converged_flag": info.converged}
    print(name, "%.1f ms/step" % ms, its, ["%.2e" % x for x in rr], flush=True)
    ctx.close()
json.dump(out, open('gpurun_out/options_r1.json', 'w'), indent=1)
