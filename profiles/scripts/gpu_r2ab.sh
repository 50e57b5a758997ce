#!/bin/bash
# round 2, call AB (1 GPU): final confirmation of the final build -- GPU tests, smoke(), default bench lines of the three workloads
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > gpurun_out/r2ab_pytest.txt 2>&1; tail -5 gpurun_out/r2ab_pytest.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r2ab_bench_parallel.json 2> gpurun_out/r2ab_bench_parallel.err
timeout 600 python bench.py --workload hyperbolic_test --steps 20 --warmup 5 > gpurun_out/r2ab_bench_hyperbolic.json 2> gpurun_out/r2ab_bench_hyperbolic.err
timeout 900 python bench.py --workload diagonal_dispersion --steps 3 --warmup 3 > gpurun_out/r2ab_bench_dispersion.json 2> gpurun_out/r2ab_bench_dispersion.err
for f in gpurun_out/r2ab_bench_*.json; do echo $f; cut -c1-200 $f; done; tail -n 3 gpurun_out/r2ab_*.err
