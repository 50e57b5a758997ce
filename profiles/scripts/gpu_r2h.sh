#!/bin/bash
# round 2, call H (1 GPU): y sweep variants and sweep prefetch distance; GPU test suite with the new tests
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
K=sweep_x,sweep_y,transport_step
for cfg in "1 0" "1 4" "1 8" "1 16" "0 8"; do
  set -- $cfg
  echo "== IMPEC_SWEEP_Y_PAIRS=$1 IMPEC_SWEEP_PREFETCH=$2" >> gpurun_out/r2h_sweeps.txt
  IMPEC_SWEEP_Y_PAIRS=$1 IMPEC_SWEEP_PREFETCH=$2 timeout 300 python tools/ncu_kernels.py 16384 $K 10 >> gpurun_out/r2h_sweeps.txt 2>&1
done
cat gpurun_out/r2h_sweeps.txt
( time timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 ) > gpurun_out/r2h_pytest.txt 2>&1
cat gpurun_out/r2h_pytest.txt
timeout 600 python bench.py --workload hyperbolic_test --steps 20 --warmup 5 > gpurun_out/r2h_bench_hyperbolic.json 2> gpurun_out/r2h_bench_hyperbolic.err
cut -c1-250 gpurun_out/r2h_bench_hyperbolic.json
