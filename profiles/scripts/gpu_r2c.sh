#!/bin/bash
# round 2, call C: A/B of the rotating-slot march kernels against the shifting-pipeline build, GPU tests, bench lines of three workloads,
# ncu launch list + full captures of the hot kernels (report kept on the box, CSV pages come home)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
K=vcycle,fused_pre,fused_post,fused_plain,pcg_matvec,build_operator,calc_velocities
timeout 300 python tools/ncu_kernels.py 16384 $K 10 > gpurun_out/r2c_ab_new.txt 2>&1
cp impec2d_b200/libimpec2d_b200.so /tmp/new.so; cp tools/studies/ab/libold.so impec2d_b200/libimpec2d_b200.so
timeout 300 python tools/ncu_kernels.py 16384 $K 10 > gpurun_out/r2c_ab_old.txt 2>&1
cp /tmp/new.so impec2d_b200/libimpec2d_b200.so
paste gpurun_out/r2c_ab_new.txt gpurun_out/r2c_ab_old.txt
( time timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | tail -30 ) > gpurun_out/r2c_pytest.txt 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c_bench_parallel.json 2> gpurun_out/r2c_bench_parallel.err
timeout 600 python bench.py --workload hyperbolic_test --steps 20 --warmup 5 > gpurun_out/r2c_bench_hyperbolic.json 2> gpurun_out/r2c_bench_hyperbolic.err
timeout 900 python bench.py --workload diagonal_dispersion --steps 3 --warmup 3 > gpurun_out/r2c_bench_dispersion.json 2> gpurun_out/r2c_bench_dispersion.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --profile-mode --steps 1 --warmup 1 > gpurun_out/r2c_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_build_links_march|k_velocity_march|k_mg_march_pre|k_mg_march_post|k_pcg_matvec_march|k_update_xr|k_muscl_x|k_muscl_y" -c 10 -f -o /tmp/r2c_hot python tools/ncu_kernels.py 16384 build_operator,calc_velocities,fused_pre,fused_post,pcg_matvec,update_xr,sweep_x,sweep_y > gpurun_out/r2c_ncu.log 2>&1
python tools/ncu_summary.py /tmp/r2c_hot.ncu-rep > gpurun_out/r2c_ncu_summary.txt 2>&1
ncu -i /tmp/r2c_hot.ncu-rep --page details --csv > gpurun_out/r2c_ncu_details.csv 2>/dev/null
ncu -i /tmp/r2c_hot.ncu-rep --page source --csv -k regex:k_mg_march_pre 2>/dev/null | gzip > gpurun_out/r2c_ncu_source_pre.csv.gz
tail -4 gpurun_out/r2c_pytest.txt; cut -c1-200 gpurun_out/r2c_bench_parallel.json; cut -c1-200 gpurun_out/r2c_bench_hyperbolic.json; cut -c1-200 gpurun_out/r2c_bench_dispersion.json; tail -n 3 gpurun_out/*.err; ls -la gpurun_out
