#!/bin/bash
# round 2, call Q (1 GPU): final single-GPU evidence -- GPU tests, bench lines of the three workloads + reference arm, launch list, ncu full
# captures of the hot kernels of the final build (rotating PRE)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q -rA --durations=8 2>&1 | tail -110 ) > gpurun_out/r2q_pytest.txt 2>&1
tail -4 gpurun_out/r2q_pytest.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2q_bench_parallel.json 2> gpurun_out/r2q_bench_parallel.err
timeout 600 python bench.py --workload hyperbolic_test --steps 20 --warmup 5 > gpurun_out/r2q_bench_hyperbolic.json 2> gpurun_out/r2q_bench_hyperbolic.err
timeout 900 python bench.py --workload diagonal_dispersion --steps 3 --warmup 3 > gpurun_out/r2q_bench_dispersion.json 2> gpurun_out/r2q_bench_dispersion.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2q_bench_reference_parallel.json 2> gpurun_out/r2q_bench_reference_parallel.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2q_launches.csv python bench.py --profile-mode --steps 1 --warmup 1 > gpurun_out/r2q_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_mg_march_pre|k_mg_march_post|k_pcg_matvec_march|k_update_xr|k_muscl_x|k_muscl_y2" -c 6 -f -o /tmp/r2q_hot python tools/ncu_kernels.py 16384 fused_pre,fused_post,pcg_matvec,update_xr,sweep_x,sweep_y 0 > gpurun_out/r2q_ncu.log 2>&1
python tools/ncu_summary.py /tmp/r2q_hot.ncu-rep > gpurun_out/r2q_ncu_summary.txt 2>&1
for f in gpurun_out/r2q_bench_*.json; do echo $f; cut -c1-200 $f; done; tail -n 3 gpurun_out/r2q_*.err; grep -n "^==\|gpu__time_dur" gpurun_out/r2q_ncu_summary.txt | cut -c1-150
