set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; tail -c 400 gpurun_out/bench_r1b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1b.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],'its',d['config']['pcg_iterations'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'])
for k,v in d['roofline']['kernel_ms'].items(): print('  %-16s %8.3f ms'%(k,v))
for k,v in d['roofline']['kernels'].items(): print('  %-16s %6.0f GB/s  frac %.3f'%(k,v['achieved_gbs'],v['frac']))
PY
ncu --set full --clock-control none --import-source on -k regex:k_mg_fused -s 30 -c 3 -o gpurun_out/prof_fused_r1b python bench.py --steps 1 --warmup 0 --profile-mode > gpurun_out/ncu_full3.log 2>&1
tail -3 gpurun_out/ncu_full3.log
