#!/bin/bash
# round 2, call F (1 GPU): fine tuning of the L2 prefetch distance per kernel family
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
K=vcycle,fused_pre,fused_post,fused_plain,pcg_matvec
for d in 0 2 3 4 5 6; do
  echo "== IMPEC_L2_PREFETCH=$d" >> gpurun_out/r2f_prefetch.txt
  IMPEC_L2_PREFETCH=$d timeout 300 python tools/ncu_kernels.py 16384 $K 10 >> gpurun_out/r2f_prefetch.txt 2>&1
done
echo "== IMPEC_L2_PREFETCH=4 IMPEC_MARCH_VARIANT=0" >> gpurun_out/r2f_prefetch.txt
IMPEC_MARCH_VARIANT=0 IMPEC_L2_PREFETCH=4 timeout 300 python tools/ncu_kernels.py 16384 fused_pre 10 >> gpurun_out/r2f_prefetch.txt 2>&1
echo "== IMPEC_L2_PREFETCH=6 IMPEC_MARCH_VARIANT=0" >> gpurun_out/r2f_prefetch.txt
IMPEC_MARCH_VARIANT=0 IMPEC_L2_PREFETCH=6 timeout 300 python tools/ncu_kernels.py 16384 fused_pre 10 >> gpurun_out/r2f_prefetch.txt 2>&1
cat gpurun_out/r2f_prefetch.txt
