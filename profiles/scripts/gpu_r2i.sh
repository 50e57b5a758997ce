#!/bin/bash
# round 2, call I (2 GPUs): multi-GPU parity tests including the C++ driver on two GPUs
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_multirank.py -m gpu -q -rA -x 2>&1 | tail -40 ) > gpurun_out/r2i_pytest_multirank.txt 2>&1
cat gpurun_out/r2i_pytest_multirank.txt | cut -c1-300
