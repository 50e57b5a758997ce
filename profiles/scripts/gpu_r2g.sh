#!/bin/bash
# round 2, call G (8 GPUs): scaling of the parallel workload, overlap / replication-threshold A/B, diagonal_dispersion at 16384^2 and
# parallel_dispersion at 32768^2 on 8 B200, multi-GPU parity tests on 8 ranks
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29711 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2g_bench_parallel_n8.json 2> gpurun_out/r2g_bench_parallel_n8.err
IMPEC_HALO_OVERLAP=0 timeout 300 $TR --master-port 29712 bench.py --gpus 8 --steps 3 --warmup 3 --no-e2e > gpurun_out/r2g_ab_overlap0_n8.json 2> gpurun_out/r2g_ab_overlap0_n8.err
IMPEC_DIST_MIN_CELLS=5000000 timeout 300 $TR --master-port 29713 bench.py --gpus 8 --steps 3 --warmup 3 --no-e2e > gpurun_out/r2g_ab_distmin5e6_n8.json 2> gpurun_out/r2g_ab_distmin5e6_n8.err
IMPEC_DIST_MIN_CELLS=5000000 IMPEC_HALO_OVERLAP=0 timeout 300 $TR --master-port 29714 bench.py --gpus 8 --steps 3 --warmup 3 --no-e2e > gpurun_out/r2g_ab_distmin5e6_overlap0_n8.json 2> gpurun_out/r2g_ab_distmin5e6_overlap0_n8.err
timeout 600 $TR --master-port 29715 bench.py --gpus 8 --workload diagonal_dispersion --steps 3 --warmup 3 > gpurun_out/r2g_bench_dispersion_n8.json 2> gpurun_out/r2g_bench_dispersion_n8.err
timeout 900 $TR --master-port 29716 bench.py --gpus 8 --workload parallel_dispersion --steps 3 --warmup 3 > gpurun_out/r2g_bench_parallel_dispersion_32768_n8.json 2> gpurun_out/r2g_bench_parallel_dispersion_32768_n8.err
timeout 300 $TR --master-port 29717 tools/gpu_comm_diag.py 16384 > gpurun_out/r2g_diag_n8.txt 2>&1; cp gpurun_out/comm_diag_n8.json gpurun_out/r2g_comm_diag_n8.json
( time timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q -rA 2>&1 | tail -12 ) > gpurun_out/r2g_pytest_multirank_n8.txt 2>&1
for f in gpurun_out/r2g_*.json; do echo $f; cut -c1-230 $f | head -2; done; tail -4 gpurun_out/r2g_pytest_multirank_n8.txt; grep -h "step\|tail\|cycle@" gpurun_out/r2g_diag_n8.txt; tail -n 3 gpurun_out/r2g_*.err | grep -v "^\*\|OMP_NUM\|^$" | head -30
