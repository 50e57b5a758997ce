mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1h_n1.json 2> gpurun_out/bench_r1h_n1.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r1h_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'its',d['config']['pcg_iterations'],'e2e',d['e2e']['value'],'vcycle',d['vcycle_ms'],'launches',d['gpu_launches'])
print({k:round(v,3) for k,v in d['roofline']['kernel_ms'].items()})
PY
python - <<PY
import sys; sys.path.insert(0,'.')
import impec2d_b200 as I
N=16384
P = I.default_params(N1=N, N2=N, x_ub=1.41421356237, y_ub=1.41421356237, stype=0, nu=2/3, eps=1e-14)
c = I.synth_state(P); q = I.init_source(P)
ctx = I.Context(P); ctx.upload_c(c); ctx.upload_q(q); ctx.impes_step(0.0)
for small in (1, 0):
    ctx.set_option("mg_small", small)
    print("mg_small", small, "tail visit %.4f ms  vcycle %.3f ms" % (ctx.profile_kernel("tail", 50), ctx.profile_kernel("vcycle", 10)))
PY
