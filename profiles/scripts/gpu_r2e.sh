#!/bin/bash
# round 2, call E (1 GPU): L2 software prefetch distance of the marching multigrid kernels (0 = off)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
K=vcycle,fused_pre,fused_post,fused_plain
for d in 0 4 8 16 32; do
  echo "== IMPEC_L2_PREFETCH=$d" >> gpurun_out/r2e_prefetch.txt
  IMPEC_L2_PREFETCH=$d timeout 300 python tools/ncu_kernels.py 16384 $K 10 >> gpurun_out/r2e_prefetch.txt 2>&1
done
cat gpurun_out/r2e_prefetch.txt
