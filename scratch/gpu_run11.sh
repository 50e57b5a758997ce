mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_mg_staged" -c 6 -o gpurun_out/prof_staged_r1g python scratch/ncu_kernels.py 16384 fused_pre,fused_post,fused_plain > gpurun_out/ncu_full6.log 2>&1
tail -2 gpurun_out/ncu_full6.log
