#!/bin/bash
# round 2, call Z (1 GPU): chunk heights of the marching kernels, whole cycle / whole step, interleaved configurations
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/studies/gpu_march_rows_ab.py 16384 256,128,32:256,128,176:256,176,176:176,176,176:64,176,176:240,140,176:512,176,176:256,132,64 > gpurun_out/r2z_march_rows_ab.txt 2>&1
tail -12 gpurun_out/r2z_march_rows_ab.txt
