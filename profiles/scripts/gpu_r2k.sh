#!/bin/bash
# round 2, call K (4 GPUs): parallel and diagonal_dispersion bench lines on 4 B200
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29721 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2k_bench_parallel_n4.json 2> gpurun_out/r2k_bench_parallel_n4.err
timeout 600 $TR --master-port 29722 bench.py --gpus 4 --workload diagonal_dispersion --steps 3 --warmup 3 > gpurun_out/r2k_bench_dispersion_n4.json 2> gpurun_out/r2k_bench_dispersion_n4.err
for f in gpurun_out/r2k_*.json; do echo $f; cut -c1-230 $f | head -2; done; tail -n 3 gpurun_out/r2k_*.err | grep -v "^\*\|OMP_NUM\|^$" | head
