#!/bin/bash
# round 2, call J (1 GPU): GPU test suite, bench lines of the three single-GPU workloads, march-threshold A/B, ncu launch list of the bench
# command and full captures of the hot kernels (the report stays on the box; summaries and CSV pages come home)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -rA 2>&1 | tail -90 ) > gpurun_out/r2j_pytest.txt 2>&1
tail -4 gpurun_out/r2j_pytest.txt
for m in 8000000 4000000 1000000; do
  echo "== IMPEC_MARCH_MIN_CELLS=$m" >> gpurun_out/r2j_march_threshold.txt
  IMPEC_MARCH_MIN_CELLS=$m timeout 300 python tools/ncu_kernels.py 16384 vcycle 10 >> gpurun_out/r2j_march_threshold.txt 2>&1
done
cat gpurun_out/r2j_march_threshold.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2j_bench_parallel.json 2> gpurun_out/r2j_bench_parallel.err
timeout 600 python bench.py --workload hyperbolic_test --steps 20 --warmup 5 > gpurun_out/r2j_bench_hyperbolic.json 2> gpurun_out/r2j_bench_hyperbolic.err
timeout 900 python bench.py --workload diagonal_dispersion --steps 3 --warmup 3 > gpurun_out/r2j_bench_dispersion.json 2> gpurun_out/r2j_bench_dispersion.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2j_bench_reference_parallel.json 2> gpurun_out/r2j_bench_reference_parallel.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2j_launches.csv python bench.py --profile-mode --steps 1 --warmup 1 > gpurun_out/r2j_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_build_links_march|k_velocity_march|k_mg_march_pre|k_mg_march_post|k_pcg_matvec_march|k_update_xr|k_muscl_x|k_muscl_y2" -c 10 -f -o /tmp/r2j_hot python tools/ncu_kernels.py 16384 build_operator,calc_velocities,fused_pre,fused_post,pcg_matvec,update_xr,sweep_x,sweep_y > gpurun_out/r2j_ncu.log 2>&1
python tools/ncu_summary.py /tmp/r2j_hot.ncu-rep > gpurun_out/r2j_ncu_summary.txt 2>&1
ncu -i /tmp/r2j_hot.ncu-rep --page details --csv > gpurun_out/r2j_ncu_details.csv 2>/dev/null
for f in gpurun_out/r2j_bench_*.json; do echo $f; cut -c1-220 $f; done; tail -n 3 gpurun_out/r2j_*.err; ls -la gpurun_out | tail -15
