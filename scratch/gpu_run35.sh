mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1r_n1.json 2> gpurun_out/bench_r1r_n1.err; tail -c 300 gpurun_out/bench_r1r_n1.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r1r_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'its',d['config']['pcg_iterations'],'e2e',d['e2e']['value'],'vcycle',d['vcycle_ms'],'launches',d['gpu_launches'],'cpu',d['cpu_baseline']['value'],d['cpu_baseline']['cores'], d['cpu_baseline']['ms_per_step'])
r=d['roofline']; print('roofline', r['achieved'], r['frac'], 'traffic', r['traffic'], 'vcycle', r['vcycle']['frac'], r['vcycle']['frac_moved'], 'step', r['step'])
print(d['clocks'])
PY
