set -x
ncu --set full --clock-control none --import-source on -k regex:k_mg_fused -c 6 -o gpurun_out/prof_fused_fine_r1b python scratch/ncu_kernels.py 16384 fused_pre,fused_post,fused_plain > gpurun_out/ncu_full4.log 2>&1
tail -4 gpurun_out/ncu_full4.log
