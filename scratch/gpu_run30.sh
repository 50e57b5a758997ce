mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1q_n1.json 2> gpurun_out/bench_r1q_n1.err; tail -c 300 gpurun_out/bench_r1q_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1q_ref.json 2>/dev/null; cat gpurun_out/bench_r1q_ref.json | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r1q_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'its',d['config']['pcg_iterations'],'e2e',d['e2e']['value'],'vcycle',d['vcycle_ms'],'launches',d['gpu_launches'],'cpu',d['cpu_baseline']['value'],d['cpu_baseline']['cores'])
print({k:round(v,3) for k,v in d['roofline']['kernel_ms'].items()})
print(d['roofline']['achieved'], d['roofline']['frac'], d['clocks'])
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_mg_march|k_pcg_matvec_march|k_update_xr|k_muscl" -c 7 -o gpurun_out/prof_r1q python scratch/ncu_kernels.py 16384 fused_pre,fused_post,fused_plain,pcg_matvec,update_xr,sweep_x,sweep_y > gpurun_out/ncu_full_r1q.log 2>&1
tail -2 gpurun_out/ncu_full_r1q.log
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 2100 --csv --log-file gpurun_out/launches_r1q.csv python bench.py --steps 1 --warmup 1 --profile-mode > gpurun_out/ncu_launch_r1q.log 2>&1
tail -1 gpurun_out/ncu_launch_r1q.log
