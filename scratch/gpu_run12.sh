mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scratch/gpu_options.py 16384 2>&1 | tail -8
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1_final_n1.json 2> gpurun_out/bench_r1_final_n1.err; tail -c 300 gpurun_out/bench_r1_final_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_final_ref.json 2>/dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r1_final_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'its',d['config']['pcg_iterations'],'e2e',d['e2e']['value'],'vcycle',d['vcycle_ms'],'launches',d['gpu_launches'],'cpu',d['cpu_baseline']['value'],d['cpu_baseline']['cores'])
print({k:round(v,3) for k,v in d['roofline']['kernel_ms'].items()})
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 1 --warmup 1 --profile-mode > gpurun_out/ncu_launch2.log 2>&1
tail -1 gpurun_out/ncu_launch2.log
ncu --set full --clock-control none --import-source on -k regex:"k_mg_fused|k_pcg_matvec|k_update_xr" -c 10 -o gpurun_out/prof_final_r1 python scratch/ncu_kernels.py 16384 fused_pre,fused_post,pcg_matvec,update_xr > gpurun_out/ncu_full7.log 2>&1
tail -2 gpurun_out/ncu_full7.log
