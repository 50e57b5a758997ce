mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
python - <<PY
import sys, os; sys.path.insert(0,'.')
import impec2d_b200 as I
N=16384
P = I.default_params(N1=N, N2=N, x_ub=1.41421356237, y_ub=1.41421356237, stype=0, nu=2/3, eps=1e-14)
c = I.synth_state(P); q = I.init_source(P)
ctx = I.Context(P); ctx.upload_c(c); ctx.upload_q(q); dt, info = ctx.impes_step(0.0)
for pf in (2, 4):
    ctx.set_option("mg_march_pf", pf)
    print("  march_pf", pf, " ".join("%s %.3f" % (k, ctx.profile_kernel(k, 10)) for k in ("fused_pre", "fused_post", "fused_plain", "fused_pre@1", "fused_post@1", "cycle@1", "cycle@2", "cycle@3", "cycle@4", "cycle@5", "vcycle")))
ctx.set_option("mg_march_pf", 2)
import time
for (wf, wt) in ((2, 4), (2, 3), (1, 3), (3, 4), (2, 5)):
    ctx.set_option("mg_w_from", wf); ctx.set_option("mg_w_to", wt)
    ctx.upload_c(c)
    dt, info = ctx.impes_step(0.0)
    ctx.upload_c(c); ctx.synchronize(); t0 = time.perf_counter()
    dt, info = ctx.impes_step(0.0); ctx.synchronize(); el = time.perf_counter() - t0
    print("W levels %d-%d: its %d relres %.2e true %.2e step %.1f ms vcycle %.3f ms" % (wf, wt, info.iterations, info.relres, info.true_relres, el * 1e3, ctx.profile_kernel("vcycle", 5)))
ctx.close()
PY
for pf in 2 4; do
IMPEC_MARCH_PF=$pf timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1o_pf$pf.json 2> gpurun_out/bench_r1o_pf$pf.err; tail -c 300 gpurun_out/bench_r1o_pf$pf.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r1o_pf$pf.json').read().strip().splitlines()[-1])
print('pf $pf value',d['value'],'ms/step',d['ms_per_step'],'its',d['config']['pcg_iterations'],'e2e',d['e2e']['value'],'vcycle',d['vcycle_ms'],'launches',d['gpu_launches'])
print({k:round(v,3) for k,v in d['roofline']['kernel_ms'].items()})
PY
done
