#!/bin/bash
# round 2, call R (N GPUs, N = $1): final parallel + dispersion bench lines of the final build on N B200
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29731 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2r_bench_parallel_n$N.json 2> gpurun_out/r2r_bench_parallel_n$N.err
timeout 600 $TR --master-port 29732 bench.py --gpus $N --workload diagonal_dispersion --steps 3 --warmup 3 > gpurun_out/r2r_bench_dispersion_n$N.json 2> gpurun_out/r2r_bench_dispersion_n$N.err
for f in gpurun_out/r2r_*_n$N.json; do echo $f; cut -c1-230 $f | head -2; done; tail -n 3 gpurun_out/r2r_*_n$N.err | grep -v "^\*\|OMP_NUM\|^$" | head
