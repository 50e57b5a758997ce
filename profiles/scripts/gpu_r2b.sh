#!/bin/bash
# round 2, call B: GPU tests with the fused build / velocity march, bench lines, ncu launch list + full captures of the hot kernels
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | tail -30 ) > gpurun_out/r2b_pytest.txt 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench_parallel.json 2> gpurun_out/r2b_bench_parallel.err
timeout 600 python bench.py --workload hyperbolic_test --steps 20 --warmup 5 > gpurun_out/r2b_bench_hyperbolic.json 2> gpurun_out/r2b_bench_hyperbolic.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --profile-mode --steps 1 --warmup 1 > gpurun_out/r2b_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_build_links_march|k_velocity_march|k_mg_march_pre|k_mg_march_post|k_pcg_matvec_march|k_update_xr|k_muscl_x|k_muscl_y" -c 18 -f -o gpurun_out/r2b_hot python tools/ncu_kernels.py 16384 build_operator,calc_velocities,fused_pre,fused_post,pcg_matvec,update_xr,sweep_x,sweep_y > gpurun_out/r2b_ncu.log 2>&1
tail -4 gpurun_out/r2b_pytest.txt; cut -c1-300 gpurun_out/r2b_bench_parallel.json; cut -c1-300 gpurun_out/r2b_bench_hyperbolic.json; tail -n 3 gpurun_out/r2b_bench_hyperbolic.err gpurun_out/r2b_ncu.log; ls -la gpurun_out
