mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; tail -c 300 gpurun_out/bench_r1f.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r1f.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'its',d['config']['pcg_iterations'],'e2e',d['e2e']['value'],'vcycle',d['vcycle_ms'],'launches',d['gpu_launches'])
print({k:round(v,3) for k,v in d['roofline']['kernel_ms'].items()})
PY
