#!/bin/bash
# round 2, call A: full GPU test suite with the new parity cases, bench lines of three workloads, reference-arm sanity
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nproc > gpurun_out/r2a_host.txt; free -g >> gpurun_out/r2a_host.txt; nvidia-smi -L >> gpurun_out/r2a_host.txt
( time timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | tail -60 ) > gpurun_out/r2a_pytest.txt 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench_parallel.json 2> gpurun_out/r2a_bench_parallel.err
timeout 300 python bench.py --workload hyperbolic_test --steps 20 --warmup 5 > gpurun_out/r2a_bench_hyperbolic.json 2> gpurun_out/r2a_bench_hyperbolic.err
timeout 600 python bench.py --workload diagonal_dispersion --steps 3 --warmup 3 > gpurun_out/r2a_bench_dispersion.json 2> gpurun_out/r2a_bench_dispersion.err
tail -5 gpurun_out/r2a_pytest.txt; cut -c1-600 gpurun_out/r2a_bench_parallel.json; cut -c1-400 gpurun_out/r2a_bench_hyperbolic.json; cut -c1-400 gpurun_out/r2a_bench_dispersion.json; tail -3 gpurun_out/*.err
