mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python - <<PY
import sys, os; sys.path.insert(0,'.')
import impec2d_b200 as I
N=16384
P = I.default_params(N1=N, N2=N, x_ub=1.41421356237, y_ub=1.41421356237, stype=0, nu=2/3, eps=1e-14)
c = I.synth_state(P); q = I.init_source(P)
for mc in ("8e6", "2e6", "3e7"):
    os.environ["IMPEC_MARCH_MIN_CELLS"] = mc
    ctx = I.Context(P); ctx.upload_c(c); ctx.upload_q(q); dt, info = ctx.impes_step(0.0)
    print("MARCH_MIN_CELLS", mc, "its", info.iterations, "relres %.2e" % info.relres)
    for march in ((1, 0) if mc == "8e6" else (1,)):
        ctx.set_option("mg_march", march)
        print("  mg_march", march, " ".join("%s %.3f" % (k, ctx.profile_kernel(k, 10)) for k in ("fused_pre", "fused_post", "fused_plain", "fused_pre@1", "fused_post@1", "fused_pre@2", "fused_post@2", "cycle@1", "cycle@2", "cycle@3", "vcycle")))
    ctx.close()
PY
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1n_n1.json 2> gpurun_out/bench_r1n_n1.err; tail -c 300 gpurun_out/bench_r1n_n1.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r1n_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'its',d['config']['pcg_iterations'],'e2e',d['e2e']['value'],'vcycle',d['vcycle_ms'],'launches',d['gpu_launches'])
print({k:round(v,3) for k,v in d['roofline']['kernel_ms'].items()})
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_mg_march" -c 4 -o gpurun_out/prof_r1n python scratch/ncu_kernels.py 16384 fused_pre,fused_post > gpurun_out/ncu_full_r1n.log 2>&1
tail -2 gpurun_out/ncu_full_r1n.log
