#!/bin/bash
# round 2, call X (1 GPU): PCG iterations, cycle time and tail-visit time for different W-cycle level ranges at 16384^2
# (input to the multi-GPU choice of the W range: the replicated tail costs the same on any number of GPUs)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/studies/gpu_cycle_shapes.py 16384 > gpurun_out/r2x_cycle_shapes.txt 2>&1
cat gpurun_out/r2x_cycle_shapes.txt | tail -20
