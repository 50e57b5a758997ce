#!/bin/bash
# round 2, call Y (1 GPU): programmatic dependent launch in the multigrid tail (on/off, bit-identity of p), chunk heights of the marching kernels
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python tools/studies/gpu_pdl.py 16384 > gpurun_out/r2y_pdl.txt 2>&1
tail -6 gpurun_out/r2y_pdl.txt
timeout 400 python tools/studies/gpu_march_rows.py 16384 64,96,112,128,132,140,148,160,176,192,208,224,240,256,288,320,384,512 > gpurun_out/r2y_march_rows.txt 2>&1
tail -60 gpurun_out/r2y_march_rows.txt
