mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "fast_initial or impes_step or dispersion" 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1u_n1.json 2> gpurun_out/bench_r1u_n1.err; tail -c 300 gpurun_out/bench_r1u_n1.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r1u_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'its',d['config']['pcg_iterations'],'true',d['config']['true_relres'][:2],'e2e',d['e2e']['value'],'vcycle',d['vcycle_ms'],'launches',d['gpu_launches'])
PY
