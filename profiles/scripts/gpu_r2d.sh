#!/bin/bash
# round 2, call D (2 GPUs): multi-GPU parity tests with the overlapped halo exchange and the peer-memory all-gather, overlap on/off,
# bench lines of the parallel and diagonal_dispersion workloads on 2 GPUs
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( time timeout 1200 python -m pytest tests/test_gpu_multirank.py -m gpu -q -s -rA 2>&1 | tail -40 ) > gpurun_out/r2d_pytest_multirank.txt 2>&1
IMPEC_VERBOSE=1 timeout 600 $TR --master-port 29701 tools/gpu_comm_diag.py 16384 > gpurun_out/r2d_diag_overlap1.txt 2>&1; cp gpurun_out/comm_diag_n2.json gpurun_out/r2d_comm_diag_n2_overlap1.json
IMPEC_VERBOSE=1 IMPEC_HALO_OVERLAP=0 timeout 600 $TR --master-port 29702 tools/gpu_comm_diag.py 16384 > gpurun_out/r2d_diag_overlap0.txt 2>&1; cp gpurun_out/comm_diag_n2.json gpurun_out/r2d_comm_diag_n2_overlap0.json
IMPEC_P2P_GATHER=0 timeout 600 $TR --master-port 29703 tools/gpu_comm_diag.py 16384 > gpurun_out/r2d_diag_ncclgather.txt 2>&1
timeout 600 $TR --master-port 29704 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2d_bench_parallel_n2.json 2> gpurun_out/r2d_bench_parallel_n2.err
timeout 600 $TR --master-port 29705 bench.py --gpus 2 --workload diagonal_dispersion --steps 3 --warmup 3 > gpurun_out/r2d_bench_dispersion_n2.json 2> gpurun_out/r2d_bench_dispersion_n2.err
tail -6 gpurun_out/r2d_pytest_multirank.txt; grep -h "step\|gather_rhs\|halo\|cycle@1 \|vcycle\|overlapped" gpurun_out/r2d_diag_*.txt | head -40; cut -c1-200 gpurun_out/r2d_bench_*_n2.json; tail -n 5 gpurun_out/r2d_*.err
