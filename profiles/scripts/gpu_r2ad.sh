#!/bin/bash
# round 2, call AD (8 GPUs): final build on eight B200 -- default parallel bench line (with e2e after the host-scan change) and the
# W-cycle range 2..3 variant (the replicated tail is entered 4 x 1 instead of 4 x 2 times per cycle; 23 instead of 20 iterations on one GPU)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29751 bench.py --gpus 8 --steps 5 --warmup 3 2> gpurun_out/r2ad_bench_parallel_n8.err | grep '^{' > gpurun_out/r2ad_bench_parallel_n8.json
timeout 200 $TR --master-port 29752 bench.py --gpus 8 --steps 5 --warmup 3 --no-e2e --opt mg_w_to=3 2> gpurun_out/r2ad_bench_parallel_n8_wto3.err | grep '^{' > gpurun_out/r2ad_bench_parallel_n8_wto3.json
for f in gpurun_out/r2ad_*.json; do echo $f; cut -c1-200 $f; done
tail -n 3 gpurun_out/r2ad_*.err | grep -v "^\*\|OMP_NUM\|^$" | head
