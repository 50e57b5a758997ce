#!/bin/bash
# round 2, call AC (2 GPUs): final build on two B200 -- multirank GPU tests (slabs vs one rank, C++ driver as two processes), parallel bench line
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -rA -x 2>&1 | tail -14 ) > gpurun_out/r2ac_pytest_multirank_n2.txt 2>&1
cut -c1-200 gpurun_out/r2ac_pytest_multirank_n2.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29741 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2ac_bench_parallel_n2.json 2> gpurun_out/r2ac_bench_parallel_n2.err
cut -c1-230 gpurun_out/r2ac_bench_parallel_n2.json; tail -n 3 gpurun_out/r2ac_bench_parallel_n2.err | grep -v "^\*\|OMP_NUM\|^$" | head
