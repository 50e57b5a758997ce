mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python - <<PY
import sys; sys.path.insert(0,'.')
import impec2d_b200 as I
N=16384
P = I.default_params(N1=N, N2=N, x_ub=1.41421356237, y_ub=1.41421356237, stype=0, nu=2/3, eps=1e-14)
c = I.synth_state(P); q = I.init_source(P)
ctx = I.Context(P); ctx.upload_c(c); ctx.upload_q(q); ctx.impes_step(0.0)
for march in (0, 1):
    ctx.set_option("matvec_march", march)
    print("matvec_march", march, "pcg_matvec %.3f ms" % ctx.profile_kernel("pcg_matvec", 10))
for sh in (0, 1):
    ctx.set_option("sweep_shared", sh)
    print("sweep_shared", sh, "sweep_x %.3f ms sweep_y %.3f ms" % (ctx.profile_kernel("sweep_x", 10), ctx.profile_kernel("sweep_y", 10)))
print("update_xr %.3f ms" % ctx.profile_kernel("update_xr", 10))
PY
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1l_n1.json 2> gpurun_out/bench_r1l_n1.err; tail -c 300 gpurun_out/bench_r1l_n1.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r1l_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'its',d['config']['pcg_iterations'],'e2e',d['e2e']['value'],'vcycle',d['vcycle_ms'],'launches',d['gpu_launches'])
print({k:round(v,3) for k,v in d['roofline']['kernel_ms'].items()})
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_pcg_matvec_march|k_update_xr|k_muscl" -c 8 -o gpurun_out/prof_r1l python scratch/ncu_kernels.py 16384 pcg_matvec,update_xr,sweep_x,sweep_y > gpurun_out/ncu_full_r1l.log 2>&1
tail -2 gpurun_out/ncu_full_r1l.log
